cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu.json')); print('8GPU value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sim', d['simulation'] and round(d['simulation']['value'],1), 'synd', round(d['config5_as_written']['value'],1), 'n_gpus', d['n_gpus'], 'sharding_check', d.get('sharding_check'))"
timeout 600 python -m pytest tests -m gpu -q -k "several_devices or multi_gpu or two_gpus" 2>&1 | tail -3
