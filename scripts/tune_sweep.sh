#!/bin/bash
# development helper: parity tests, then the headline bench for several flag sets
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for f in ${FLAGS:-0 4}; do
python bench.py --steps 100 --warmup 5 --no-cpu --flags $f 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags', d['config']['flags'], '%.1f Gcoord/s'%(d['value']/1e9), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2f'%(d['e2e']['value']/1e9))"
done
