"""Latency of the single-sequence GEMMs (M = 452 rows): per launch in a CUDA graph of 48 back-to-back launches, weights hot
(one weight buffer) or cold (48 different weight buffers, 170 MB > what stays in L2 next to the rest), plus a minimal GEMM
(M = 128, N = 64, K = 64) for the kernel's fixed cost."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops

def bench(M, N, K, act, resid, cold, n=48):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    ws = [(torch.randn(N, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16) for _ in range(n if cold else 1)]
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g) if resid else None
    out = r if resid else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    def run():
        for i in range(n):
            ops.gemm(a, ws[i % len(ws)], b, act, r, None, out=out)
    run(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
        with torch.cuda.graph(gr, stream=s):
            run()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(5):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return round(min(ts), 2)

for name, (M, N, K, act, resid) in {"tiny": (128, 64, 64, 0, False), "qkv": (452, 2304, 768, 0, False), "proj": (452, 768, 768, 0, True),
                                    "fc1": (452, 3072, 768, 1, False), "fc2": (452, 768, 3072, 0, True)}.items():
    print(name, {"hot_us": bench(M, N, K, act, resid, False), "cold_us": bench(M, N, K, act, resid, True)})
