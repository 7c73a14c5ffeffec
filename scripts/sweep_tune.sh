#!/bin/bash
# development helper: sweep-kernel tuning (VBQ_SWEEP_TUNE=<threads/128><lambdas per group>)
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for t in ${TUNES:-41 51 61}; do
VBQ_SWEEP_TUNE=$t python - <<PY
import json, sys
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import bench_configs as b
from vbq_b200 import ops
for L in (16, 64):
    for f in (0, ops.FLAG_REFERENCE_WALK, ops.FLAG_FAST):
        r = b.sweep_case(L, f, outputs=False)
        print("tune $t L", L, "flags", f, "%.1f G coord*lambda/s" % (r["coord_lambda_per_s"] / 1e9))
r = b.sweep_case(16, 0, outputs=True)
print("tune $t L 16 full outputs %.1f G" % (r["coord_lambda_per_s"] / 1e9))
PY
done
