"""Micro-benchmark of the LayerNorm kernel at the bench shape (28 928 rows x 768, fp32 in, bf16 out)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops
for rows in (28928, 57856):
    C = 768
    xs = [torch.randn(rows, C, device="cuda") for _ in range(4)]      # rotate buffers: > L2
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    out = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
    for x in xs:
        ops.layernorm(x, g, b, None, None, 0, 1e-6, out_bf16=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        ops.layernorm(xs[i % 4], g, b, None, None, 0, 1e-6, out_bf16=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    print(json.dumps({"rows": rows, "us": round(us, 1), "GBps": round(rows * C * 6 / us / 1e3, 1)}))
