#!/bin/bash
# development helper: sweep-kernel tuning (VBQ_SWEEP_TUNE=<threads/128><lambdas per group>)
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sweep or index_parity" 2>&1 | tail -2
for t in 22 21 24 41 42 32; do
VBQ_SWEEP_TUNE=$t python - <<PY
import json, sys
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import bench_configs as b
for L in (16, 64):
    r = b.sweep_case(L, 0, outputs=False)
    print("tune $t L", L, "%.1f G coord*lambda/s" % (r["coord_lambda_per_s"] / 1e9))
PY
done
