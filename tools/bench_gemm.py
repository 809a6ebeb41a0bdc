"""Micro-benchmark of the tcgen05 GEMM on the backbone shapes (CUDA events, L2-busting rotation).
The `timing` mode (MMA-thread stall counters) and the MMT_GEMM_DBG isolation switches exist only in the developer library:
build it with `python multi-modal-tracking_b200/build.py --dev` and run with MMT_B200_DEV_LIB=1."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops


def bench(M, N, K, act=0, resid=False, iters=20):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if resid else None
    out = r if resid else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, bias, act, r, None, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, bias, act, r, None, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS reference for context
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    return {"M": M, "N": N, "K": K, "act": act, "resid": resid, "ms": round(ms, 4), "tflops": round(tf, 1),
            "cublas_ms": round(ms_ref, 4), "cublas_tflops": round(2.0 * M * N * K / ms_ref / 1e9, 1)}


if __name__ == "__main__":
    M = 57856
    if len(sys.argv) > 2 and sys.argv[2] == "timing":     # MMA-thread stall breakdown (developer hook)
        import ctypes
        from mmt_b200 import _lib
        hook = _lib.fn("mmt_dev_gemm_timing")
        case = {"qkv": (M, 2304, 768, 0, False), "proj": (M, 768, 768, 0, True), "fc1": (M, 3072, 768, 1, False),
                "fc2": (M, 768, 3072, 0, True)}[sys.argv[1]]
        dbg = torch.zeros(148 * 4, device="cuda", dtype=torch.int64)
        Mm, N, K, act, resid = case
        a = torch.randn(Mm, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        r = torch.randn(Mm, N, device="cuda") if resid else None
        out = r if resid else torch.empty(Mm, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, bias, act, r, None, out=out)
        hook(ctypes.c_void_p(dbg.data_ptr()))
        ops.gemm(a, w, bias, act, r, None, out=out)
        hook(ctypes.c_void_p(0))
        torch.cuda.synchronize()
        d = dbg.view(148, 4).cpu().numpy().astype("float64")
        d = d[d[:, 0] > 0]
        print(sys.argv[1], "pair" if os.environ.get("MMT_GEMM_PAIR", "1") != "0" else "single", "MMA threads:", len(d),
              "total cyc %.0f  stalled on epilogue %.1f%%  stalled on TMA data %.1f%%  tiles/CTA %.1f" % (
                  d[:, 0].mean(), 100 * d[:, 1].sum() / d[:, 0].sum(), 100 * d[:, 2].sum() / d[:, 0].sum(), d[:, 3].mean()))
        sys.exit(0)
    if len(sys.argv) > 1:      # single case for ncu: qkv | proj | fc1 | fc2
        case = {"qkv": (M, 2304, 768, 0, False), "proj": (M, 768, 768, 0, True), "fc1": (M, 3072, 768, 1, False),
                "fc2": (M, 768, 3072, 0, True)}[sys.argv[1]]
        print(json.dumps(bench(*case, iters=3)), flush=True)
        sys.exit(0)
    for args in [(M, 2304, 768, 0, False), (M, 768, 768, 0, True), (M, 3072, 768, 1, False), (M, 768, 3072, 0, True),
                 (M // 2, 2304, 768, 0, False), (904, 2304, 768, 0, False)]:
        print(json.dumps(bench(*args)), flush=True)
