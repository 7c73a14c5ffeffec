python -m pytest tests -m gpu -q 2>&1 | tail -5
for t in 0; do
VBQ_TUNE=$t python bench.py --steps 30 --warmup 3 --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tune $t flags', d['config']['flags'], '%.1f Gcoord/s'%(d['value']/1e9), 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done
python bench.py --steps 30 --warmup 3 --no-cpu --flags 130 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bracket walk', d['config']['flags'], '%.1f Gcoord/s'%(d['value']/1e9), 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
