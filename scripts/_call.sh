cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export NO_NCU=1
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default nola park0"
unset NO_NCU
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default"
