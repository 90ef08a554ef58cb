#!/bin/bash
# A/B of tuning builds on one workload, then a light ncu metric pass per build (run under gpurun):
#   gpurun -- 'bash scripts/gpu_variants.sh WORKLOAD FRAMES "default cap0 park0" [extra bench args]'
WL=${1:-AD_64800_R12_GF256}; FR=${2:-592}; VARS=${3:-default}; shift; shift; shift
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
for V in $VARS; do
  if [ $V = default ]; then unset NBLDPC_B200_LIB; else export NBLDPC_B200_LIB=$GRAFT_REPO_ROOT/ems-decoder-of-nb-ldpc-codes_b200/variants/$V.so; fi
  CMD="python bench.py --steps 3 --warmup 2 --no-cpu --no-also --workload $WL --frames $FR $*"
  $CMD > gpurun_out/var_${WL}_$V.json 2> gpurun_out/var_${WL}_$V.err
  RC=$?
  python - <<PY
import json
try:
    j = json.load(open('gpurun_out/var_${WL}_$V.json'))
    print('VARIANT', '$WL', '$V', 'value', round(j['value'], 2), 'Mbit/s kernel_ms', round(j['roofline']['kernel_ms'], 3), 'frac', round(j['roofline']['frac'], 4), 'clk', j['clocks']['sm_mhz'], 'slow', j['counters']['slow_path_selects'], j['geometry'])
except Exception as e:
    print('VARIANT', '$WL', '$V', 'FAILED rc=$RC', e)
PY
  if [ $RC = 0 ] && [ -z "$NO_NCU" ]; then
    ncu --metrics $M --clock-control none -k regex:decode_kernel -s 3 -c 1 --csv --log-file gpurun_out/ncum_${WL}_$V.csv python bench.py --steps 1 --warmup 2 --no-cpu --no-also --workload $WL --frames $FR $* > gpurun_out/ncum_${WL}_$V.log 2>&1
    python - <<PY
import csv
try:
    rows = [r for r in csv.reader(open('gpurun_out/ncum_${WL}_$V.csv')) if len(r) > 10]
    h = rows[0]; i_n = h.index('Metric Name'); i_v = h.index('Metric Value')
    d = {r[i_n]: r[i_v] for r in rows[1:]}
    for k in sorted(d): print('   NCU', '$V', k, d[k])
except Exception as e:
    print('   NCU', '$V', 'unreadable', e)
PY
  fi
done
