#!/bin/bash
# development helper: parity tests, then bench over kernel tuning variants (VBQ_TUNE=<U><threads/256>) and modes
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for t in ${TUNES:-22 23 42 41}; do for f in ${FLAGS:-0 4}; do
VBQ_TUNE=$t python bench.py --steps 20 --warmup 3 --no-cpu --flags $f 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tune $t flags', d['config']['flags'], '%.1f Gcoord/s'%(d['value']/1e9), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2f'%(d['e2e']['value']/1e9))"
done; done
