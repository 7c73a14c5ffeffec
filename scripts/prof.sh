#!/bin/bash
# development helper: parity tests, bench (exact + fast), one ncu --set full capture of the quantize kernel
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for f in ${FLAGS:-0 4}; do
python bench.py --steps 20 --warmup 3 --no-cpu --flags $f 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags', d['config']['flags'], '%.1f Gcoord/s'%(d['value']/1e9), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2f'%(d['e2e']['value']/1e9))"
done
ncu --set full --clock-control none --import-source on -k regex:vbq_bisect_kernel -s 3 -c 1 -o gpurun_out/${TAG:-prof} python bench.py --steps 3 --warmup 3 --no-cpu --flags ${PFLAGS:-0} > gpurun_out/${TAG:-prof}.log 2>&1
