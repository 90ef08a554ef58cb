#!/bin/bash
# syndrome check node: parity tests then the syndrome bench line
timeout 900 python -m pytest tests -m gpu -x -q -k "synd or fuzz" 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu --ecn syndrome 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('SYND', round(j['value'],2), 'Mbit/s e2e', round(j['e2e']['value'],2), j['geometry'])"
