"""Micro-benchmark of the ConvMAE stem's depthwise 5x5 kernel (back-to-back launches, CUDA events)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops
for (B, H, E) in ((64, 72, 256), (64, 36, 384), (64, 32, 256), (64, 16, 384)):
    x = torch.randn(B * H * H, E, device="cuda").to(torch.bfloat16)
    w = torch.randn(25, E, device="cuda")
    b = torch.randn(E, device="cuda")
    out = torch.empty_like(x)
    for _ in range(3):
        ops.dwconv5x5(x, w, b, B, H, H, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.dwconv5x5(x, w, b, B, H, H, out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(json.dumps({"B": B, "H": H, "E": E, "us": round(us, 1), "GBps": round(x.numel() * 4 / us / 1e3, 1)}))
