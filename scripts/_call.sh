cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "synd or fuzz or random_conf or check_node" 2>&1 | tail -4
python bench.py --steps 2 --warmup 3 --no-cpu --no-also --ecn syndrome --frames 592 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('SYND value',d['value'],'kernel_ms',d['roofline']['kernel_ms'])"
