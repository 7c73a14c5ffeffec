python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/bench_configs.py emb sweep 2>&1 | cut -c1-200
python bench.py --steps 200 --no-cpu | cut -c1-120
python bench.py --steps 200 --no-cpu | cut -c1-120
