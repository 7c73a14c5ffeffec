python scripts/bench_configs.py sweep 2>&1 | cut -c1-200 | grep '"flags": 0'
python -m pytest tests -m gpu -q -x -k "sweep or index_parity or fullsize or golden" 2>&1 | tail -2
