cd $GRAFT_REPO_ROOT
export NO_NCU=1
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default ua4 ub4 uab4" --ecn syndrome
