#!/bin/bash
# Developer helper (run under gpurun): one bench line per workload (no CPU baseline leg).
mkdir -p gpurun_out
for WL in AD_64800_R12_GF256 Ahmed_64800_R34_GF16 MatDeclercq_R12_GF64 Mat24_N480_M240 N96_K48_GF64 KN_64800_R34_GF256; do
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --workload $WL "$@" > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err || tail -3 gpurun_out/bench_$WL.err
  python - <<PY
import json
j = json.load(open('gpurun_out/bench_$WL.json'))
print('VALUE', '$WL', round(j['value'], 2), 'Mbit/s', round(j['frames_per_s'], 1), 'frames/s  frac', round(j['roofline']['frac'], 4), 'e2e', round(j['e2e']['value'], 2), 'sim', round(j['simulation']['value'], 2), 'also', round((j.get('also') or {}).get('value', 0), 2), j['geometry'])
PY
done
