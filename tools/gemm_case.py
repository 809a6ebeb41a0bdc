"""One GEMM case a few times, for `ncu --set full` captures:  python tools/gemm_case.py <qkv|fc1|proj|fc2> <plain|ln> [M]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmt_b200  # noqa
from mmt_b200 import ops
if os.environ.get("MMT_CL4") == "0":
    ops.config_cluster4(False)

name, mode = sys.argv[1], sys.argv[2]
M = int(sys.argv[3]) if len(sys.argv) > 3 else 28928
N, K, act, resid = {"qkv": (2304, 768, 0, False), "proj": (768, 768, 0, True), "fc1": (3072, 768, 1, False),
                    "fc2": (768, 3072, 0, True)}[name]
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
b = torch.randn(N, device="cuda", generator=g)
if resid:
    r = torch.randn(M, N, device="cuda", generator=g)
    xb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sums = torch.empty(N // 128, M, 2, device="cuda")
    fn = (lambda: ops.gemm(a, w, b, act, r, None, out=r)) if mode == "plain" else \
        (lambda: ops.gemm(a, w, b, act, r, None, out=r, xb_out=xb, stats_out=sums))
else:
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sums = torch.rand(K // 128, M, 2, device="cuda") + 1.0
    cs = w.float().sum(1).contiguous()
    fn = (lambda: ops.gemm(a, w, b, act, out=out)) if mode == "plain" else \
        (lambda: ops.gemm(a, w, b, act, out=out, ln_stats=sums, ln_eps=1e-6, colsum=cs))
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok", name, mode)
