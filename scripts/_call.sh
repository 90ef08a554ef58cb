cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 3 --warmup 2 --no-cpu --no-also --ecn syndrome --frames 592 > gpurun_out/synd.json 2> gpurun_out/synd.err; tail -3 gpurun_out/synd.err
python - <<PY
import json
j = json.load(open('gpurun_out/synd.json'))
print('SYND value', round(j['value'], 2), 'kernel_ms', round(j['roofline']['kernel_ms'], 2), j['geometry'])
PY
