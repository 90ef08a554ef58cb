cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export NO_NCU=1
bash scripts/gpu_variants.sh AD_64800_R12_GF256 2368 "default"
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default" --ecn syndrome
bash scripts/gpu_variants.sh MatDeclercq_R12_GF64 4096 "default"
bash scripts/gpu_variants.sh Mat24_N480_M240 65536 "default"
bash scripts/gpu_variants.sh KN_64800_R34_GF256 2368 "default"
