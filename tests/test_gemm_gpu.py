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


# ---------------------------------------------------------------------------- folded LayerNorm (mmt_gemm_bf16_ex)
def _fold(W, bias, gamma, beta):
    """Host-side fold of LayerNorm(gamma, beta) into Linear(W, bias): (W' bf16, bias', colsum) - engine._pack_backbone."""
    Wf = (W * gamma[None, :]).to(torch.bfloat16)
    return Wf, bias + W @ beta, Wf.float().sum(dim=1)


@pytest.mark.parametrize("M,dim,N,act", [(904, 768, 2304, 0), (19000, 768, 3072, 1), (700, 1024, 3072, 0), (133, 768, 64, 1)])
def test_folded_layernorm_consumer(built_lib, M, dim, N, act):
    """Linear(LayerNorm(x)) through the folded form: A = bf16(x), statistics from mmt_rowstats_cast, normalisation in the
    GEMM epilogue - against torch fp32 LayerNorm + Linear (+ GELU) of the same fp32 rows.  Rows carry a per-row offset of
    up to one standard deviation (the |mean|/std range the residual stream shows)."""
    from mmt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, dim, device="cuda", generator=g) * 1.3 + torch.randn(M, 1, device="cuda", generator=g)
    W = torch.randn(N, dim, device="cuda", generator=g) * dim ** -0.5
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = 1.0 + 0.1 * torch.randn(dim, device="cuda", generator=g)
    beta = 0.05 * torch.randn(dim, device="cuda", generator=g)
    Wf, bf, cs = _fold(W, bias, gamma, beta)
    xb = torch.empty(M, dim, device="cuda", dtype=torch.bfloat16)
    sums = torch.empty(dim // 128, M, 2, device="cuda")          # slot-major [slots, rows, 2]
    ops.rowstats_cast(x, xb, sums)
    out = ops.gemm(xb, Wf.contiguous(), bf.contiguous(), act, ln_stats=sums, ln_eps=1e-6, colsum=cs.contiguous(),
                   out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert (sums[:, :, 0].sum(0) - x.sum(1)).abs().max().item() <= 1e-3
    assert ((sums[:, :, 1].sum(0) - (x * x).sum(1)).abs() / (x * x).sum(1)).max().item() <= 1e-5
    if M > 200:      # row-range call (modality-specific norms: one consumer launch per row range of the same buffers)
        r0 = 96
        part = ops.gemm(xb[r0:], Wf.contiguous(), bf.contiguous(), act, ln_stats=sums[:, r0:], ln_eps=1e-6,
                        colsum=cs.contiguous(), out_dtype=torch.bfloat16)
        assert torch.equal(part, out[r0:])
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (dim,), gamma, beta, 1e-6), W, bias)
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    err = (out.float() - ref).abs().max().item()
    # bf16 rounding of x (2^-9 of |x| <= ~5 sigma), of W' and of the output: the same budget as bf16(LN(x)) @ bf16(W)
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), err
    # and the unfused bf16 pipeline (LayerNorm kernel -> bf16 -> plain GEMM) is no closer to fp32 than the folded one x 2
    h = torch.empty(M, dim, device="cuda", dtype=torch.bfloat16)
    ops.layernorm(x, gamma.contiguous(), beta.contiguous(), None, None, 0, 1e-6, out_bf16=h)
    unf = ops.gemm(h, W.to(torch.bfloat16).contiguous(), bias.contiguous(), act, out_dtype=torch.bfloat16)
    e_unf = (unf.float() - ref).abs().mean().item()
    e_fold = (out.float() - ref).abs().mean().item()
    print(f"folded LN consumer M={M} N={N}: mean abs err folded {e_fold:.3e}  unfused {e_unf:.3e}  max folded {err:.3e}")
    assert e_fold <= 2.0 * e_unf + 1e-4


@pytest.mark.parametrize("M,N,K", [(19000, 768, 768), (18945, 768, 3072), (904, 768, 768), (6500, 1024, 1024), (300, 256, 64)])
def test_folded_layernorm_producer(built_lib, M, N, K):
    """Residual GEMM that also leaves the bf16 copy and the per-row partial sums of its fp32 output (fused epilogue of
    the CTA-pair kernel for the big shapes, mmt_rowstats_cast behind the other kernels): the fp32 output is bit-identical
    to the plain call, xb is its bf16 rounding, the slot sums add up to the row sums."""
    from mmt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g)
    plain = x0.clone()
    ops.gemm(a, w, bias, 0, plain, None, out=plain)
    out = x0.clone()
    xb = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    sums = torch.full((N // 128, M, 2), float("nan"), device="cuda")
    ops.gemm(a, w, bias, 0, out, None, out=out, xb_out=xb, stats_out=sums)
    torch.cuda.synchronize()
    assert torch.equal(out, plain)
    assert torch.equal(xb, out.to(torch.bfloat16))
    assert torch.isfinite(sums).all()
    s1, s2 = sums[:, :, 0].sum(0), sums[:, :, 1].sum(0)
    assert (s1 - out.sum(1)).abs().max().item() <= 2e-3
    assert ((s2 - (out * out).sum(1)).abs() / (out * out).sum(1)).max().item() <= 1e-5
    # the stand-alone producer accumulates every slot in the fused epilogue's order: bit-identical partial sums
    xb2 = torch.empty_like(xb)
    sums2 = torch.full_like(sums, float("nan"))
    ops.rowstats_cast(out, xb2, sums2)
    torch.cuda.synchronize()
    assert torch.equal(xb2, xb) and torch.equal(sums2, sums)


def test_folded_layernorm_argument_checks(built_lib):
    from mmt_b200 import ops
    a = torch.zeros(64, 768, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(256, 768, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(256, device="cuda")
    sums = torch.zeros(6, 64, 2, device="cuda")
    with pytest.raises(RuntimeError):        # fp32 output with the consumer-side fold
        ops.gemm(a, w, b, 0, ln_stats=sums, ln_eps=1e-6, colsum=b, out_dtype=torch.float32)
    with pytest.raises(AssertionError):      # the fp32 parity kernel has no folded form
        ops.gemm(a.float(), w.float(), b, 0, ln_stats=sums, ln_eps=1e-6, colsum=b)


@pytest.mark.parametrize("case", [(19000, 2304, 768, 0, False), (18945, 768, 768, 0, True), (13000, 3072, 768, 1, False)])
def test_cluster_of_four_is_bit_identical(built_lib, case):
    """mmt_config_cluster4: two CTA pairs per cluster sharing the weight tile by TMA multicast (off by default) must
    reproduce the CTA-pair launch bit for bit - qkv-like, proj-like (in-place residual + shadow rows, odd row count: the
    last cluster tile has one real and one empty 256-row block) and fc1-like (GELU) shapes."""
    from mmt_b200 import ops
    M, N, K, act, resid = case
    g = torch.Generator(device="cuda").manual_seed(M)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g) if resid else None

    def run():
        if resid:
            out = x0.clone()
            xb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            sums = torch.empty(N // 128, M, 2, device="cuda")
            ops.gemm(a, w, bias, act, out, None, out=out, xb_out=xb, stats_out=sums)
            return out, xb, sums
        return (ops.gemm(a, w, bias, act, out_dtype=torch.bfloat16),)

    prev = ops.config_cluster4(False)
    try:
        ref = run()
        ops.config_cluster4(True)
        got = run()
        torch.cuda.synchronize()
    finally:
        ops.config_cluster4(prev)
    for r, g_ in zip(ref, got):
        assert torch.equal(r, g_)


def test_small_gemm_sm_budget_changes_no_result(built_lib):
    """mmt_config_small_gemm_sms: with half the SMs as budget the one-wave rule picks wider tiles for the single-sequence
    GEMMs (M = 452) - the K order per output element is the same, so every result is bit-identical."""
    from mmt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    M = 452
    for N, K, act, resid in ((2304, 768, ops.ACT_NONE, False), (3072, 768, ops.ACT_GELU, False), (768, 3072, ops.ACT_NONE, True),
                             (768, 768, ops.ACT_NONE, True)):
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        b = torch.randn(N, device="cuda", generator=g)
        r = torch.randn(M, N, device="cuda", generator=g) if resid else None
        outs = []
        for budget in (0, 74, 37):
            prev = ops.config_small_gemm_sms(budget)
            try:
                o = torch.empty(M, N, device="cuda", dtype=torch.float32 if resid else torch.bfloat16)
                ops.gemm(a, w, b, act, r, None, out=o)
                torch.cuda.synchronize()
                outs.append(o)
            finally:
                ops.config_small_gemm_sms(prev)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), (N, K)
