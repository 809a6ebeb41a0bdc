"""Which kernels cuBLAS picks for the four backbone GEMM shapes (names carry tile and cluster shape), and their
CUDA-event rate, back to back with ours in the same process (interleaved A/B, 30 launches each, 3 rounds)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops

M = 57856
CASES = {"qkv": (2304, 768, 0, False), "proj": (768, 768, 0, True), "fc1": (3072, 768, 1, False), "fc2": (768, 3072, 0, True)}


def main():
    from torch.profiler import profile, ProfilerActivity
    for name, (N, K, act, resid) in CASES.items():
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if resid else None
        out = r if resid else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            torch.matmul(a, w.t())
            torch.cuda.synchronize()
        names = sorted({e.name for e in prof.events() if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()})
        res = {"mine": [], "cublas": []}
        for rnd in range(3):
            for which in ("mine", "cublas"):
                fn = (lambda: ops.gemm(a, w, bias, act, r, None, out=out)) if which == "mine" else (lambda: torch.matmul(a, w.t()))
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(30):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[which].append(round(2.0 * M * N * K * 30 / e0.elapsed_time(e1) / 1e9, 1))
        print(name, "N", N, "K", K, "mine TF/s", res["mine"], "cuBLAS TF/s", res["cublas"], "cuBLAS kernels:", names, flush=True)


if __name__ == "__main__":
    main()
