"""GPU parity of the implicit-GEMM 3x3 convolution (tcgen05 + 4-D TMA boxes) and of upsample_add against
torch.nn.functional on the same bf16 inputs, at the pyramid-head shapes (lib/models/mixformer_cvt/head.py:159-198)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# the torch references must be true fp32 (cuDNN / cuBLAS default to TF32 for convolutions on this GPU)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# B, H, W, C, N, channel-slice offset of the input inside a wider buffer (None = dense)
CASES = [(2, 18, 18, 768, 1344, None), (2, 18, 18, 384, 192, 384), (3, 36, 36, 192, 96, None),
         (2, 72, 72, 96, 48, None), (2, 18, 18, 192, 96, None), (2, 18, 18, 48, 1, None), (1, 36, 36, 96, 48, None),
         (2, 36, 36, 48, 1, None), (1, 24, 24, 1024, 384, None), (2, 7, 5, 64, 32, None)]


@pytest.mark.parametrize("case", CASES)
def test_conv3x3_matches_torch(built_lib, case):
    from mmt_b200 import ops
    B, H, W, C, N, off = case
    g = torch.Generator(device="cuda").manual_seed(H * 100 + C + N)
    wide = C if off is None else off + C + 64
    buf = torch.randn(B * H * W, wide, device="cuda", generator=g).to(torch.bfloat16)
    src = buf if off is None else buf[:, off:off + C]
    w4 = (torch.randn(N, C, 3, 3, device="cuda", generator=g) * (9 * C) ** -0.5)
    wp = w4.permute(0, 2, 3, 1).reshape(N, 9 * C).to(torch.bfloat16).contiguous()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((B * H * W, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv3x3(src, B, H, W, C, wp, bias, ops.ACT_RELU, out)
    torch.cuda.synchronize()
    x = src.float().view(B, H, W, C).permute(0, 3, 1, 2)
    wr = wp.float().view(N, 3, 3, C).permute(0, 3, 1, 2)
    ref = F.relu(F.conv2d(x, wr, bias, padding=1)).permute(0, 2, 3, 1).reshape(B * H * W, N)
    assert bool(torch.isfinite(out.float()).all()), "unwritten output rows"
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2 * max(1.0, ref.abs().max().item()), err     # bf16 output rounding (2^-9 relative)


def test_upsample_add_matches_torch(built_lib):
    from mmt_b200 import ops
    B, gs, C = 2, 18, 96
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(B * gs * gs, C, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(B * 4 * gs * gs, C, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(B * 16 * gs * gs, C, device="cuda", dtype=torch.bfloat16)
    ops.upsample_add(a, 4, B, 4 * gs, 4 * gs, C, out, b, 2)
    torch.cuda.synchronize()
    am = a.float().view(B, gs, gs, C).permute(0, 3, 1, 2)
    bm = b.float().view(B, 2 * gs, 2 * gs, C).permute(0, 3, 1, 2)
    ref = (F.interpolate(am, scale_factor=4) + F.interpolate(bm, scale_factor=2)).permute(0, 2, 3, 1).reshape(-1, C)
    assert (out.float() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 72, 72, 256), (3, 16, 16, 384), (1, 9, 7, 64)])
def test_dwconv5x5_and_patchify2x2_match_torch(built_lib, shape, dtype):
    """ConvMAE stem kernels against torch: depthwise Conv2d(E, E, 5, padding 2, groups E) and the 2x2/2 patch matrix."""
    from mmt_b200 import ops
    B, H, W, E = shape
    g = torch.Generator(device="cuda").manual_seed(H + E)
    x = torch.randn(B * H * W, E, device="cuda", generator=g).to(dtype)
    w4 = torch.randn(E, 1, 5, 5, device="cuda", generator=g) * 0.2
    bias = torch.randn(E, device="cuda", generator=g)
    out = torch.empty_like(x)
    ops.dwconv5x5(x, w4.reshape(E, 25).t().contiguous(), bias, B, H, W, out)
    ref = F.conv2d(x.float().view(B, H, W, E).permute(0, 3, 1, 2), w4, bias, padding=2, groups=E)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, E)
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert (out.float() - ref).abs().max().item() <= tol
    if H % 2 == 0 and W % 2 == 0:
        xf = x.float().contiguous()
        pm = torch.empty(B * (H // 2) * (W // 2), 4 * E, device="cuda")
        ops.patchify2x2(xf, B, H, W, pm)
        wq = torch.randn(8, E, 2, 2, device="cuda", generator=g)
        got = pm @ wq.permute(0, 2, 3, 1).reshape(8, -1).t()
        want = F.conv2d(xf.view(B, H, W, E).permute(0, 3, 1, 2), wq, stride=2).permute(0, 2, 3, 1).reshape(-1, 8)
        assert (got - want).abs().max().item() <= 1e-3
