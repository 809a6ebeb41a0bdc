"""GPU parity of the two native-op replacements through the C ABI: mmt_prroi_fwd against the PrRoIPool oracle
(incl. the reference's avg_pool2d known answer), mmt_msda_fwd / mmt_msda_bimodal_fwd against the MSDA oracle
(incl. the reference test's toy shapes, deformable_attention/ops/test.py:31-60)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_prroi_known_answer_and_layouts(built_lib):
    from mmt_b200 import ops
    from oracle import native_ops_oracle as NO
    torch.manual_seed(0)
    features = torch.rand(4, 16, 24, 32)
    rois = torch.tensor([[0, 0, 0, 14, 14], [1, 14, 14, 28, 28]], dtype=torch.float32)
    out = ops.prroi_pool(features.cuda(), rois.cuda(), 7, 7, 0.5).cpu()
    golden = F.avg_pool2d(features, kernel_size=2, stride=1)
    assert torch.allclose(out[0], golden[0, :, :7, :7], atol=1e-5)
    assert torch.allclose(out[1], golden[1, :, 7:14, 7:14], atol=1e-5)
    # token layout (NHWC in, [R, ph*pw, C] out) == the reference layout transposed
    out_cl = ops.prroi_pool(features.permute(0, 2, 3, 1).contiguous().cuda(), rois.cuda(), 7, 7, 0.5,
                            channels_last=True).cpu()
    assert torch.allclose(out_cl.view(2, 7, 7, 16).permute(0, 3, 1, 2), out, atol=1e-6)
    assert np.allclose(out.numpy(), NO.prroi_pool_forward(features.numpy(), rois.numpy(), 7, 7, 0.5), atol=1e-5)


def test_prroi_random_rois_match_oracle(built_lib):
    from mmt_b200 import ops
    from oracle import native_ops_oracle as NO
    rng = np.random.default_rng(5)
    feat = rng.standard_normal((2, 8, 18, 18)).astype(np.float32)
    rois = np.array([[0, 2.3, 4.1, 9.7, 12.2], [1, -2.0, -1.5, 6.0, 7.5], [1, 10.0, 11.0, 21.0, 19.5],
                     [0, 5.0, 5.0, 5.0, 9.0], [0, 0.0, 0.0, 17.0, 17.0]], dtype=np.float32)   # SPM: 4x4 bins, scale 1
    out = ops.prroi_pool(torch.from_numpy(feat).cuda(), torch.from_numpy(rois).cuda(), 4, 4, 1.0).cpu().numpy()
    ref = NO.prroi_pool_forward(feat, rois, 4, 4, 1.0)
    assert np.abs(out - ref).max() <= 1e-5
    with pytest.raises(NotImplementedError):
        ops.prroi_pool(torch.from_numpy(feat), torch.from_numpy(rois), 4, 4, 1.0)     # CPU tensors: no fallback


def _msda_inputs(N, M, D, Lq, shapes, P, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    L = len(shapes)
    value = (torch.rand(N, S, M, D, generator=g) * 0.01).to(dtype)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.4 - 0.2          # includes out-of-map samples
    attn = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, loc, attn


@pytest.mark.parametrize("cfg", [(1, 2, 2, 2, [(6, 4), (3, 2)], 2), (2, 8, 64, 648, [(18, 18), (18, 18)], 4),
                                 (3, 4, 30, 17, [(5, 7)], 3)])
def test_msda_generic_matches_oracle(built_lib, cfg):
    from mmt_b200 import ops
    from oracle import mixformer_oracle as O
    N, M, D, Lq, shapes, P = cfg
    value, loc, attn = _msda_inputs(N, M, D, Lq, shapes, P, 3)
    ref = O.msda_core(value, shapes, loc, attn)
    out = ops.msda(value.cuda(), shapes, loc.cuda(), attn.cuda()).cpu()
    assert (out - ref).abs().max().item() <= 1e-6 + 1e-5 * ref.abs().max().item()
    # bf16 value / output (the fast mode's storage type): rtol 1e-2, atol 1e-3 like the reference's fp32 check
    out16 = ops.msda(value.cuda().bfloat16(), shapes, loc.cuda(), attn.cuda()).float().cpu()
    assert torch.allclose(out16, ref, rtol=2e-2, atol=1e-3 * max(1.0, ref.abs().max().item()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_msda_bimodal_fused_matches_oracle(built_lib, dtype):
    """Fused form: raw offset|logit projection rows in, reference points + normalisation + softmax + sampling inside
    for fp32 and bf16 value / output storage."""
    from mmt_b200 import ops
    from oracle import mixformer_oracle as O
    B, H, W, M, D, P = 2, 18, 18, 8, 64, 4
    HW = H * W
    g = torch.Generator().manual_seed(9)
    value = torch.randn(B, 2 * HW, M * D, generator=g).to(dtype).float()      # bf16-representable in the bf16 case
    offw = torch.cat([torch.randn(B * HW, M * 2 * P * 2, generator=g) * 2.5, torch.randn(B * HW, M * 2 * P, generator=g)], 1)
    out = torch.empty(B * 2 * HW, M * D, device="cuda", dtype=dtype)
    ops.msda_bimodal(value.view(B * 2 * HW, M * D).to(dtype).cuda().contiguous(), offw.cuda().contiguous(), out, B, H, W, M, D, P)
    # reference formulation (ms_deform_attn_bimodal.py:108-118) on the same numbers
    off = offw[:, :M * 2 * P * 2].view(B, HW, M, 2, P, 2)
    aw = F.softmax(offw[:, M * 2 * P * 2:].view(B, HW, M, 2 * P), -1).view(B, HW, M, 2, P)
    ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
    ref_pts = torch.stack((rx.reshape(-1) / W, ry.reshape(-1) / H), -1).view(1, HW, 1, 1, 1, 2)
    loc = ref_pts + off / torch.tensor([W, H], dtype=torch.float32)
    loc2, aw2 = torch.cat([loc, loc], 1), torch.cat([aw, aw], 1)      # both modalities' queries share them
    ref = O.msda_core(value.view(B, 2 * HW, M, D), [(H, W), (H, W)], loc2, aw2)
    tol = 2e-5 if dtype == torch.float32 else 1e-2          # bf16: output rounding only (2^-9 relative)
    assert (out.float().cpu().view(B, 2 * HW, M * D) - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
