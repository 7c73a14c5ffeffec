python -m pytest tests -m gpu -q -x -k "host_pipeline or golden" 2>&1 | tail -2
for cr in 9216 12288 6144; do python bench.py --steps 20 --no-cpu --chunk-rows $cr | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $cr e2e %.2f G' % (d['e2e']['value']/1e9))"; done
