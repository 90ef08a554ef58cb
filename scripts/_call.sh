cd $GRAFT_REPO_ROOT
export NO_NCU=1
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default nopf" 2>&1 | grep -E "VARIANT"
bash scripts/gpu_variants.sh KN_64800_R34_GF256 2368 "default nopf" 2>&1 | grep -E "VARIANT"
bash scripts/gpu_variants.sh MatDeclercq_R12_GF64 4096 "default nopf" 2>&1 | grep -E "VARIANT"
bash scripts/gpu_variants.sh Ahmed_64800_R34_GF16 4096 "default nopf" 2>&1 | grep -E "VARIANT"
bash scripts/gpu_variants.sh Mat24_N480_M240 65536 "default nopf" 2>&1 | grep -E "VARIANT"
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default nopf" --ecn syndrome 2>&1 | grep -E "VARIANT"
