cd $GRAFT_REPO_ROOT
timeout 120 python scripts/e2e_probe.py 2>&1 | tail -10
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value'],2),'e2e',round(d['e2e']['value'],2), d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']); o=d['config5_as_written']; print('synd',round(o['value'],2),'e2e',round(o['e2e']['value'],2))"
