python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/bench_configs.py deep 2>&1 | cut -c1-220
python bench.py --steps 200 --no-cpu | cut -c1-120
