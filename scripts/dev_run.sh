python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import torch, time, numpy as np
from vbq_b200 import ops
n = 24*32*48*192
q = torch.randint(0, 2047, (n,), dtype=torch.int32, device="cuda")
for name, fn in (("pack", lambda: ops.pack_indices(q, 10)), ("hist", lambda: ops.symbol_histogram(q.view(-1, 192), 10))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print(name, "%.1f us" % (t * 1e6), "%.1f G symbols/s" % (n / t / 1e9), "%.0f GB/s" % (n * (4 + 11 / 8) / t / 1e9 if name == "pack" else n * 4 / t / 1e9))
w = ops.pack_indices(q, 10)
for _ in range(3): ops.unpack_indices(w, n, 10)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.unpack_indices(w, n, 10)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
print("unpack %.1f us %.1f G symbols/s" % (t * 1e6, n / t / 1e9))
PY
