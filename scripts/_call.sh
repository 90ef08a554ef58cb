cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
export NO_NCU=1
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default"
bash scripts/gpu_variants.sh MatDeclercq_R12_GF64 4096 "default"
bash scripts/gpu_variants.sh Ahmed_64800_R34_GF16 4096 "default"
bash scripts/gpu_variants.sh Mat24_N480_M240 65536 "default"
