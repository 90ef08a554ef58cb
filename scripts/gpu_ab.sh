#!/bin/bash
# A/B of two builds of the library on a few workloads: NBLDPC_B200_LIB selects the variant
for WL in "$@"; do
  for V in current variant_old current variant_old; do
    if [ $V = current ]; then unset NBLDPC_B200_LIB; else export NBLDPC_B200_LIB=$GRAFT_REPO_ROOT/ems-decoder-of-nb-ldpc-codes_b200/$V.so; fi
    python bench.py --steps 3 --warmup 3 --no-cpu --no-also --workload $WL 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$WL', '$V', round(j['value'],2), 'e2e', round(j['e2e']['value'],2), j['clocks']['sm_mhz'])"
  done
done
