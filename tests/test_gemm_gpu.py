"""GPU parity of the two GEMM kernels (tcgen05 bf16 and SIMT fp32) against a plain torch fp32
reference of the same op, through the C ABI.  Tolerances are written next to each case."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, act, resid, rowadd):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    if rowadd is not None:
        idx = torch.arange(y.shape[0], device=y.device) % rowadd.shape[0]
        y = y + rowadd[idx]
    if resid is not None:
        y = y + resid
    return y


CASES = [
    # M, N, K, bias, act, resid, rowadd_period, out_fp32
    (904, 2304, 768, True, 0, False, 0, False),      # qkv
    (904, 768, 768, True, 0, True, 0, True),         # proj + residual (in place)
    (904, 3072, 768, True, 1, False, 0, False),      # fc1 + GELU
    (904, 768, 3072, True, 0, True, 0, True),        # fc2 + residual
    (452, 768, 768, True, 0, False, 452, True),      # patch embed + pos table
    (1296, 512, 768, True, 0, False, 0, True),       # fusion 1x1 conv
    (648, 192, 1024, True, 0, False, 0, True),       # fusion offsets|weights
    (1296, 2048, 512, True, 2, False, 0, False),     # fusion FFN relu
    (648, 1344, 6912, True, 2, False, 0, False),     # head conv1|adjust1|adjust2 (BN=192)
    (2592, 96, 1728, True, 2, False, 0, False),      # head conv3
    (5184, 48, 864, True, 2, False, 0, False),       # head conv4 (K tail: 864 = 13.5 * 64)
    (324, 1, 432, True, 2, False, 0, True),          # 1-channel adjust conv (scalar tail path)
    (19000, 2304, 768, True, 0, False, 0, False),    # CTA-pair kernel (cta_group::2): qkv, ragged M (74.2 pair tiles)
    (18945, 768, 768, True, 0, True, 0, True),       # CTA-pair: proj + in-place residual, odd M
    (6400, 3072, 768, True, 1, False, 0, False),     # CTA-pair: fc1 + GELU
    (19200, 768, 3072, True, 0, True, 0, True),      # CTA-pair: fc2 + residual, K = 48 slices
    (130, 40, 72, False, 0, False, 0, False),        # ragged everything
    (100, 300, 200, True, 1, True, 7, True),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_gemm_matches_torch(built_lib, case, mode):
    from mmt_b200 import ops
    M, N, K, has_bias, act, has_resid, period, out_fp32 = case
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    a = torch.randn(M, K, device="cuda", generator=g).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(dt)
    bias = torch.randn(N, device="cuda", generator=g) if has_bias else None
    resid = torch.randn(M, N, device="cuda", generator=g) if has_resid else None
    rowadd = torch.randn(period, N, device="cuda", generator=g) if period else None
    ref = _ref(a, w, bias, act, resid, rowadd)
    out_dtype = torch.float32 if (out_fp32 or mode == "fp32") else torch.bfloat16
    if has_resid:
        out = resid.clone()   # in-place residual stream update, as the model does
        ops.gemm(a, w, bias, act, out, rowadd, out=out)
    else:
        out = ops.gemm(a, w, bias, act, None, rowadd, out_dtype=out_dtype)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    if mode == "fp32":
        tol = 2e-5 * max(scale, 1.0)            # fp32 FMA vs cuBLAS fp32: reduction-order noise only
    elif out_dtype == torch.float32:
        tol = 2e-3 * max(scale, 1.0)            # bf16 inputs are exact in both; fp32 accumulate
    else:
        tol = 1e-2 * max(scale, 1.0)            # + bf16 output rounding (2^-9 relative)
    assert err <= tol, f"max err {err} > {tol} (scale {scale})"


def test_gemm_strided_views(built_lib):
    """A and out as column slices of wider buffers (the head writes branch outputs side by side)."""
    from mmt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    big_a = torch.randn(700, 1024, device="cuda", generator=g).to(torch.bfloat16)
    a = big_a[:, 256:256 + 384]
    w = (torch.randn(192, 384, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    big_out = torch.zeros(700, 512, device="cuda", dtype=torch.bfloat16)
    out = big_out[:, 64:64 + 192]
    ops.gemm(a, w, None, 0, None, None, out=out)
    ref = a.float() @ w.float().t()
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert big_out[:, :64].abs().max().item() == 0 and big_out[:, 256:].abs().max().item() == 0


def test_gemm_rejects_bad_args(built_lib):
    from mmt_b200 import ops
    a = torch.zeros(8, 20, device="cuda", dtype=torch.bfloat16)[:, :12]   # lda=20 not a multiple of 8
    w = torch.zeros(8, 16, device="cuda", dtype=torch.bfloat16)[:, :12]
    with pytest.raises(RuntimeError):
        ops.gemm(a, w)
    with pytest.raises(NotImplementedError):
        ops.gemm(torch.zeros(8, 16), torch.zeros(8, 16))
