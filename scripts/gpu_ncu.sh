#!/bin/bash
# one `ncu --set full` capture of the decode kernel of a workload (after the same command exited 0 without ncu)
#   gpurun -- 'bash scripts/gpu_ncu.sh WORKLOAD FRAMES [extra bench args]'
WL=${1:-AD_64800_R12_GF256}; FR=${2:-592}; shift; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-also --workload $WL --frames $FR $*"
$CMD > gpurun_out/ncu_pre_$WL.json 2> gpurun_out/ncu_pre_$WL.err && \
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 3 -c 1 -f -o gpurun_out/prof_$WL $CMD > gpurun_out/ncu_$WL.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_$WL.log
