"""Developer check: the tcgen05 attention kernel against a torch fp32 reference at other model shapes
(default: MixViT-L - 16 heads, 288 template + 576 search tokens; also logits scaled up to stress the softmax)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops
import test_attention_gpu as T

for heads, Lt, Ls, gain in ((16, 288, 576, 1.0), (16, 288, 576, 4.0), (12, 128, 324, 4.0), (16, 288, 576, 8.0)):
    T.HEADS, T.HD = heads, 64
    T.C = C = heads * 64
    nseq, N = 2, Lt + Ls
    torch.manual_seed(0)
    qkv = torch.randn(nseq * N, 3 * C, device="cuda")
    qkv[:, :2 * C] *= gain ** 0.5            # logits x gain
    qkv = qkv.to(torch.bfloat16)
    tiles, segs = T._tiles(nseq, N, Lt, Ls, False)
    out = torch.empty(nseq * N, C, device="cuda", dtype=torch.bfloat16)
    mk = max(sum(l for _, l in s[2]) for s in segs)
    ops.mixattn(qkv, None, C, heads, tiles.cuda(), mk, out, 64 ** -0.5)
    ref = T._reference(qkv, segs, 64 ** -0.5)
    d = (out.float() - ref).abs()
    rows = d.max(dim=1).values
    worst = int(rows.argmax())
    print(f"heads {heads} Lt {Lt} Ls {Ls} logit gain {gain}: max abs diff {float(d.max()):.4e} (row {worst}, row-in-seq {worst % N}), mean {float(d.mean()):.3e}, ref max {float(ref.abs().max()):.2f}")
