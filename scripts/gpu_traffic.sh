#!/bin/bash
# DRAM bytes per launch of the decode kernel at the bench's launch size, with the commit and the kernel duration of the capture
# (bench.py drops the figure when the live kernel duration differs by more than 3 %):
#   gpurun -- 'bash scripts/gpu_traffic.sh WORKLOAD ECN [FRAMES]'   ->  gpurun_out/traffic_WORKLOAD_ECN.json (copy to profiles/)
WL=${1:-AD_64800_R12_GF256}; ECN=${2:-bubble}; FR=${3:-0}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-also --workload $WL --ecn $ECN --frames $FR"
$CMD > gpurun_out/traffic_pre_${WL}_$ECN.json 2> gpurun_out/traffic_pre_${WL}_$ECN.err && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:decode_kernel -s 3 -c 1 --csv \
    --log-file gpurun_out/traffic_${WL}_$ECN.csv $CMD > gpurun_out/traffic_ncu_${WL}_$ECN.log 2>&1
python - <<PY
import csv, json, subprocess
pre = json.load(open('gpurun_out/traffic_pre_${WL}_$ECN.json'))
rows = [r for r in csv.reader(open('gpurun_out/traffic_${WL}_$ECN.csv')) if len(r) > 10]
h = rows[0]; n, v, u = h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
def val(name):
    r = [x for x in rows[1:] if x[n] == name][0]
    x = float(r[v].replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3, 'inst': 1}.get(r[u], 1)
rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
out = {"workload": "$WL", "ecn": "$ECN", "frames": pre["roofline"]["frames_per_launch"], "dram_bytes_read": rd, "dram_bytes_write": wr,
       "dram_bytes_per_launch": rd + wr, "kernel_ms": pre["roofline"]["kernel_ms"], "kernel_ms_under_ncu": val('gpu__time_duration.sum'),
       "warp_instructions": val('smsp__inst_executed.sum'), "algorithmic_bytes_per_launch": pre["roofline"]["bytes_per_frame"] * pre["roofline"]["frames_per_launch"],
       "git": pre.get("git"), "source_id": pre.get("source_id"), "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one launch of the bench's resident-input step"}
out["traffic_over_algorithmic"] = out["dram_bytes_per_launch"] / out["algorithmic_bytes_per_launch"]
json.dump(out, open('gpurun_out/traffic_${WL}_$ECN.json', 'w'), indent=1)
print('TRAFFIC', json.dumps(out))
PY
