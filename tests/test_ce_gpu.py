"""GPU parity of the candidate-elimination score kernels (fp32 SIMT kernel and the tcgen05 bf16 kernel) against a
plain torch fp32 evaluation of the reference formula (asymmetric_shared_ce.py:202-205, :91-92):
softmax over all 2*Ls search keys of [template rows of both modalities] x [search keys of both modalities], mean over
the 2*Lt rows, mean over heads."""
import pytest
import torch

pytestmark = pytest.mark.gpu

HEADS, HD = 12, 64
C = HEADS * HD


@pytest.mark.parametrize("Ls", [324, 227, 159])
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_ce_scores_match_torch(built_lib, mode, Ls):
    from mmt_b200 import ops
    B, Lt = 3, 128
    N = Lt + Ls
    g = torch.Generator(device="cuda").manual_seed(Ls)
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    qkv = (torch.randn(2 * B * N, 3 * C, device="cuda", generator=g) * 1.2).to(dt)
    partial = torch.empty(B * HEADS * 8 * 2 * Ls, device="cuda")
    scores = torch.empty(B, 2 * Ls, device="cuda")
    ops.ce_scores(qkv, C, HEADS, B, N, Lt, Ls, HD ** -0.5, partial, scores)
    torch.cuda.synchronize()
    x = qkv.float().view(2, B, N, 3, HEADS, HD)
    q = torch.cat([x[0, :, :Lt, 0], x[1, :, :Lt, 0]], dim=1).permute(0, 2, 1, 3)          # [B, H, 2Lt, hd]
    k = torch.cat([x[0, :, Lt:, 1], x[1, :, Lt:, 1]], dim=1).permute(0, 2, 1, 3)          # [B, H, 2Ls, hd]
    ref = ((q @ k.transpose(-2, -1)) * HD ** -0.5).softmax(dim=-1).mean(dim=2).mean(dim=1)
    err = (scores - ref).abs().max().item()
    tol = 2e-8 if mode == "fp32" else 2e-6          # scores are ~1/(2 Ls) = 1.5e-3; bf16 P: 2^-9 per term, averaged
    assert err <= tol, (err, ref.abs().max().item())
    assert abs(scores.sum(dim=1) - 1.0).max().item() <= 1e-3      # rows of attn sum to 1 -> column means sum to 1
