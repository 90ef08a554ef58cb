cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -4
timeout 900 python -m pytest tests -m gpu -x -q -k "several_devices or apsk64 or reference_main or 40th" 2>&1 | tail -6
timeout 600 python bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu --no-also --frames 592 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err
python - <<PY
import json
j = json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('2GPU value', round(j['value'], 2), 'e2e', round(j['e2e']['value'], 2), 'n_gpus', j['n_gpus'], 'sharding_check', j.get('sharding_check'))
PY
