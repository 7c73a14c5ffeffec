run() { python bench.py --steps 50 --warmup 5 --no-cpu $2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['config']['flags'], '%.1f Gcoord/s'%(d['value']/1e9), 'step_ms %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2f' % (d['e2e']['value']/1e9))"; }
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
VBQ_TUNE=0 run graph768
VBQ_TUNE=5 run graph640
python scripts/bench_configs.py sweep 2>&1 | cut -c1-220
