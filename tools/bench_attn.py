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

if len(sys.argv) > 2 and sys.argv[2] == "timing":
    import ctypes
    from mmt_b200 import _lib
    f = _lib.fn("mmt_dev_attn_timing")
    n_tiles = int(tiles.shape[0])
    dbg = torch.zeros(HEADS * n_tiles * 8, device="cuda", dtype=torch.int64)
    st = f(ctypes.c_void_p(qkv.data_ptr()), ctypes.c_int(qkv.shape[0]), ctypes.c_int(qkv.stride(0)), ctypes.c_int(C),
           ctypes.c_int(HEADS), ctypes.c_void_p(tiles.data_ptr()), ctypes.c_int(n_tiles), ctypes.c_void_p(out.data_ptr()),
           ctypes.c_int(out.stride(0)), ctypes.c_float(HD ** -0.5), ctypes.c_void_p(dbg.data_ptr()),
           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0, st
    torch.cuda.synchronize()
    d = dbg.view(HEADS, n_tiles, 8).cpu().numpy().astype("float64")
    names = ["setup(alloc,barriers)", "first S ready (TMA Q+K, MMA)", "pass 1 rest", "pass 2", "wait O", "epilogue", "teardown"]
    qrows = tiles[:, 1].cpu().numpy()
    nkeys = tiles[:, 7:10].sum(1).cpu().numpy()
    for label, sel in (("template tiles (128 keys)", nkeys <= 128), ("search tiles 128 rows", (nkeys > 128) & (qrows == 128)),
                       ("search tiles tail rows", (nkeys > 128) & (qrows < 128))):
        dd = d[:, sel, :]
        seg = dd[:, :, 1:] - dd[:, :, :-1]
        print(label, "n =", int(sel.sum()) * HEADS, "total cycles median", float(np.median(dd[:, :, 7] - dd[:, :, 0])))
        for i, nme in enumerate(names):
            print(f"   {nme:32s} median {np.median(seg[:, :, i]):9.0f}  p90 {np.percentile(seg[:, :, i], 90):9.0f}")
