"""Micro-benchmark of the tcgen05 mixed-attention kernel at the bench shapes (64 sequences x 452 tokens, 12 heads)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_attention_gpu import _tiles, C, HEADS, HD

nseq, Lt, Ls = 64, 128, 324
cross = len(sys.argv) > 1 and sys.argv[1] == "cross"
iters = 3 if (len(sys.argv) > 2 and sys.argv[2] == "short") else 20
N = Lt + Ls
qkv = torch.randn(nseq * N, 3 * C, device="cuda").to(torch.bfloat16)
tiles, segs = _tiles(nseq, N, Lt, Ls, cross)
if os.environ.get("SORT_TILES", "1") != "0":
    tiles = torch.from_numpy(ops.order_tiles(tiles.numpy()))
tiles = tiles.cuda()
out = torch.empty(nseq * N, C, device="cuda", dtype=torch.bfloat16)
mk = max(sum(l for _, l in s[2]) for s in segs)
for _ in range(3):
    ops.mixattn(qkv, None, C, HEADS, tiles, mk, out, HD ** -0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.mixattn(qkv, None, C, HEADS, tiles, mk, out, HD ** -0.5)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / iters * 1e3
flops = sum(4.0 * qn * sum(l for _, l in sg) * HD for _, qn, sg in segs) * HEADS
print(json.dumps({"cross": cross, "us": round(us, 1), "tflops": round(flops / us / 1e6, 1), "tiles": int(tiles.shape[0])}))
