"""Copy-only replica of the host pipeline's stream pattern (no kernel): what do the DMA engines give us?"""
import torch, time
R, C = 36864, 192
h_in = [torch.empty((R, C), dtype=torch.float32).pin_memory() for _ in range(2)]
h_out = [torch.empty((R, C), dtype=torch.float32).pin_memory() for _ in range(2)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(chunk, dep=True, n=20):
    nchunks = (R + chunk - 1) // chunk
    d_in = [[torch.empty((chunk, C), device="cuda") for _ in range(2)] for _ in range(3)]
    d_out = [[torch.empty((chunk, C), device="cuda") for _ in range(2)] for _ in range(3)]
    def once():
        ev_in = [torch.cuda.Event() for _ in range(nchunks)]
        ev_out = [torch.cuda.Event() for _ in range(nchunks)]
        for k in range(nchunks):
            a, b = k * chunk, min(R, (k + 1) * chunk)
            sl = k % 3
            with torch.cuda.stream(s_in):
                if k >= 3: s_in.wait_event(ev_out[k - 3])
                for j in range(2): d_in[sl][j][:b - a].copy_(h_in[j][a:b], non_blocking=True)
                ev_in[k].record(s_in)
            with torch.cuda.stream(s_out):
                if dep: s_out.wait_event(ev_in[k])
                for j in range(2): h_out[j][a:b].copy_(d_out[sl][j][:b - a], non_blocking=True)
                ev_out[k].record(s_out)
        s_in.synchronize(); s_out.synchronize()
    once(); once()
    t0 = time.perf_counter()
    for _ in range(n): once()
    return (time.perf_counter() - t0) / n * 1e3

for chunk in (2304, 4608, 9216, 18432):
    print("chunk %6d: with chunk dependency %.3f ms | independent streams %.3f ms" % (chunk, run(chunk, True), run(chunk, False)))
