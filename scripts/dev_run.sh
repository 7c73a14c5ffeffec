python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/profile_round.sh 2>&1 | tail -6
python bench.py --impl reference --steps 5 --warmup 3 | cut -c1-200
