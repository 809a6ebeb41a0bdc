"""(set MMT_B200_DEV_LIB=1 MMT_GEMM_DBG=<bits> for the isolation switches of the developer library: 64 = consumer without
the statistics loads, 128 = without the column-sum reads, 256 = producer without the bf16 shadow store, 512 = without the
statistics store)
Micro-benchmark of the folded-LayerNorm GEMM forms against the plain ones on the backbone shapes (CUDA events,
interleaved A/B in one process, 3 rounds x 30 launches): where does the fold cost GEMM time?"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops
if os.environ.get("MMT_CL4") == "0":
    ops.config_cluster4(False)

M = int(sys.argv[1]) if len(sys.argv) > 1 else 28928


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dim = 768
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, dim, device="cuda", generator=g)
    xb = x.to(torch.bfloat16)
    sums = torch.empty(dim // 128, M, 2, device="cuda")
    ops.rowstats_cast(x, xb, sums)
    h = torch.empty(M, dim, device="cuda", dtype=torch.bfloat16)
    gam, bet = torch.ones(dim, device="cuda"), torch.zeros(dim, device="cuda")
    for name, N, K, act in (("qkv", 2304, 768, 0), ("fc1", 3072, 768, 1)):
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
        b = torch.randn(N, device="cuda", generator=g)
        cs = w.float().sum(1).contiguous()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        res = {"plain": [], "plain_on_raw_rows": [], "ln_consumer": [], "layernorm_kernel": []}
        ops.layernorm(x, gam, bet, None, None, 0, 1e-6, out_bf16=h)
        for _ in range(3):
            res["plain"].append(timeit(lambda: ops.gemm(h, w, b, act, out=out)))
            res["plain_on_raw_rows"].append(timeit(lambda: ops.gemm(xb, w, b, act, out=out)))
            res["ln_consumer"].append(timeit(lambda: ops.gemm(xb, w, b, act, out=out, ln_stats=sums, ln_eps=1e-6, colsum=cs)))
            res["layernorm_kernel"].append(timeit(lambda: ops.layernorm(x, gam, bet, None, None, 0, 1e-6, out_bf16=h)))
        print(name, "M", M, {k: [round(v, 1) for v in vs] for k, vs in res.items()}, "us", flush=True)
    for name, N, K in (("proj", 768, 768), ("fc2", 768, 3072)):
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
        b = torch.randn(N, device="cuda", generator=g)
        r = torch.randn(M, N, device="cuda", generator=g)
        res = {"plain": [], "ln_producer": [], "rowstats_kernel": []}
        for _ in range(3):
            res["plain"].append(timeit(lambda: ops.gemm(a, w, b, 0, r, None, out=r)))
            res["ln_producer"].append(timeit(lambda: ops.gemm(a, w, b, 0, r, None, out=r, xb_out=xb, stats_out=sums)))
            res["rowstats_kernel"].append(timeit(lambda: ops.rowstats_cast(x, xb, sums)))
        print(name, "M", M, {k: [round(v, 1) for v in vs] for k, vs in res.items()}, "us", flush=True)


if __name__ == "__main__":
    main()
