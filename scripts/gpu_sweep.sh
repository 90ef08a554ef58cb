#!/bin/bash
# Developer helper (run under gpurun): bench one workload under several tuning environments.
#   gpurun -- 'bash scripts/gpu_sweep.sh WORKLOAD "ENV1" "ENV2" ...'   e.g. "NBGPU_WARPS=12 NBGPU_CPW=8"
WL=$1; shift
mkdir -p gpurun_out
for E in "$@"; do
  env $E timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-also --workload $WL > gpurun_out/sweep.json 2> gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
  python - "$E" <<PY
import json, sys
j = json.load(open('gpurun_out/sweep.json'))
print('SWEEP', sys.argv[1], '|', round(j['value'], 2), 'Mbit/s', round(j['frames_per_s'], 1), 'f/s', j['geometry'])
PY
done
