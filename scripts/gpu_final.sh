#!/bin/bash
# End-of-round evidence set (run under gpurun): parity tests, default bench line + reference arm, per-launch DRAM traffic of the benched
# kernels, ncu digests and a launch list.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -2 gpurun_out/final_ref.err
bash scripts/gpu_traffic.sh AD_64800_R12_GF256 bubble
bash scripts/gpu_traffic.sh AD_64800_R12_GF256 syndrome
bash scripts/gpu_traffic.sh Ahmed_64800_R34_GF16 bubble
for WL in MatDeclercq_R12_GF64 Ahmed_64800_R34_GF16 Mat24_N480_M240 N96_K48_GF64 KN_64800_R34_GF256; do
  python bench.py --steps 3 --warmup 3 --no-cpu --no-also --workload $WL > gpurun_out/final_bench_$WL.json 2> gpurun_out/final_bench_$WL.err
done
# launch list of a short default-shaped run (after the same command ran plainly above with other sizes: run it plainly first)
CMD="python bench.py --steps 2 --warmup 3 --frames 296 --no-cpu"
$CMD > gpurun_out/final_ll_pre.json 2> gpurun_out/final_ll_pre.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ll_ncu.log 2>&1
