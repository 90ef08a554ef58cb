cd $GRAFT_REPO_ROOT
bash scripts/gpu_ncu.sh AD_64800_R12_GF256 296
