#!/bin/bash
# Developer helper (run under gpurun): GPU parity tests, then one bench line of the given workload.
#   gpurun -- 'bash scripts/gpu_quick.sh [workload] [extra bench args]'
WL=${1:-AD_64800_R12_GF256}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-also --workload $WL "$@" > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err
python - <<PY
import json
j = json.load(open('gpurun_out/bench_$WL.json'))
print('VALUE', '$WL', round(j['value'], 2), 'Mbit/s', round(j['frames_per_s'], 1), 'frames/s  frac', round(j['roofline']['frac'], 4), j['geometry'])
PY
tail -3 gpurun_out/bench_$WL.err
