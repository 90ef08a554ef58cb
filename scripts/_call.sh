cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash scripts/gpu_variants.sh AD_64800_R12_GF256 592 "default p3rev p3rev_park0" 2>&1 | grep -v "stalled\|pipe_\|bank\|warps_active\|issue_active\|wavefronts"
