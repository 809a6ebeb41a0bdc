"""GPU parity of the mixed-attention kernels (tcgen05 bf16 kernel and the fp32 parity kernel) against a plain
torch fp32 evaluation of the reference formula softmax(q k^T * scale) v over each query tile's key segments
(lib/models/mixformer_vit/mixformer.py:51-77; cross-modal asymmetric_shared.py:55-104), through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HEADS, HD = 12, 64
C = HEADS * HD


def _tiles(nseq, N, Lt, Ls, cross):
    """Same per-tile key-segment table the engine builds (engine.ForwardEngine._attn_tiles)."""
    recs, segs_of_q = [], []

    def add(q0, qn, segs):
        for o in range(0, qn, 128):
            r = [q0 + o, min(128, qn - o), q0 + o, len(segs)]
            rows = [s[0] for s in segs] + [0] * (3 - len(segs))
            lens = [s[1] for s in segs] + [0] * (3 - len(segs))
            recs.append(r + rows + lens + [0, 0, 0] + [0, 0, 0])
            segs_of_q.append((q0 + o, min(128, qn - o), segs))

    if not cross:
        for s in range(nseq):
            base = s * N
            add(base, Lt, [(base, Lt)])
            add(base + Lt, Ls, [(base, Lt + Ls)])
    else:
        B = nseq // 2
        for m in range(2):
            for b in range(B):
                base = (m * B + b) * N
                add(base, Lt, [(base, Lt)])
                add(base + Lt, Ls, [(b * N, Lt), ((B + b) * N, Lt), (base + Lt, Ls)])
    return torch.tensor(recs, dtype=torch.int32), segs_of_q


def _reference(qkv, segs_of_q, scale):
    x = qkv.float()
    out = torch.zeros(x.shape[0], C, device=x.device)
    for q0, qn, segs in segs_of_q:
        krows = torch.cat([torch.arange(r0, r0 + ln, device=x.device) for r0, ln in segs])
        for h in range(HEADS):
            q = x[q0:q0 + qn, h * HD:(h + 1) * HD]
            k = x[krows, C + h * HD:C + (h + 1) * HD]
            v = x[krows, 2 * C + h * HD:2 * C + (h + 1) * HD]
            a = ((q @ k.t()) * scale).softmax(dim=-1)
            out[q0:q0 + qn, h * HD:(h + 1) * HD] = a @ v
    return out


# (nseq, Lt, Ls, cross): full MixViT-B shapes, the candidate-elimination lengths (227/159/112), tiny ragged ones
CASES = [(2, 128, 324, False), (2, 128, 324, True), (4, 128, 227, True), (2, 128, 159, True), (2, 128, 112, True),
         (3, 128, 324, False), (2, 64, 37, False), (2, 32, 70, True)]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_mixattn_matches_torch(built_lib, case, mode):
    from mmt_b200 import ops
    nseq, Lt, Ls, cross = case
    N = Lt + Ls
    g = torch.Generator(device="cuda").manual_seed(nseq * 1000 + Ls)
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    qkv = (torch.randn(nseq * N, 3 * C, device="cuda", generator=g) * 1.5).to(dt)
    tiles, segs = _tiles(nseq, N, Lt, Ls, cross)
    max_keys = max(sum(l for _, l in s[2]) for s in segs)
    out = torch.full((nseq * N, C), float("nan"), device="cuda", dtype=dt)
    ops.mixattn(qkv, None, C, HEADS, tiles.cuda(), max_keys, out, HD ** -0.5)
    torch.cuda.synchronize()
    ref = _reference(qkv, segs, HD ** -0.5)
    assert bool(torch.isfinite(out.float()).all()), "unwritten or non-finite output rows"
    err = (out.float() - ref).abs().max().item()
    # fp32: reduction-order noise only.  bf16: P and the output are rounded to bf16 (2^-9 relative) -> 1e-2 of |v| ~ 4
    tol = 2e-5 if mode == "fp32" else 2.5e-2
    assert err <= tol, f"max err {err} > {tol}"


def test_mixattn_peaked_rows(built_lib):
    """Large logits (one dominant key per row): the two-pass softmax must not overflow or lose the peak."""
    from mmt_b200 import ops
    nseq, Lt, Ls = 2, 128, 324
    N = Lt + Ls
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(nseq * N, 3 * C, device="cuda", generator=g)
    qkv[:, :2 * C] *= 6.0                        # |q.k| * scale reaches ~300
    qkv = qkv.to(torch.bfloat16)
    tiles, segs = _tiles(nseq, N, Lt, Ls, False)
    out = torch.empty((nseq * N, C), device="cuda", dtype=torch.bfloat16)
    ops.mixattn(qkv, None, C, HEADS, tiles.cuda(), N, out, HD ** -0.5)
    torch.cuda.synchronize()
    ref = _reference(qkv, segs, HD ** -0.5)
    assert bool(torch.isfinite(out.float()).all())
    assert (out.float() - ref).abs().max().item() <= 4e-2


def test_mixattn_growing_maxima(built_lib):
    """Scores that keep growing along the key axis force the lazy running maximum of the single-pass softmax to move
    (and the O accumulator in TMEM to be rescaled) in every key block."""
    from mmt_b200 import ops
    nseq, Lt, Ls = 2, 128, 324
    N = Lt + Ls
    g = torch.Generator(device="cuda").manual_seed(21)
    qkv = torch.randn(nseq * N, 3 * C, device="cuda", generator=g)
    ramp = (torch.arange(nseq * N, device="cuda") % N).float() / N            # 0..1 along each sequence's tokens
    qkv[:, :C] = qkv[:, :C].abs() * 1.5                                       # positive queries
    qkv[:, C:2 * C] = qkv[:, C:2 * C].abs() * (0.2 + 4.0 * ramp[:, None])     # keys grow with the position
    qkv = qkv.to(torch.bfloat16)
    tiles, segs = _tiles(nseq, N, Lt, Ls, False)
    out = torch.empty((nseq * N, C), device="cuda", dtype=torch.bfloat16)
    ops.mixattn(qkv, None, C, HEADS, tiles.cuda(), N, out, HD ** -0.5)
    torch.cuda.synchronize()
    ref = _reference(qkv, segs, HD ** -0.5)
    assert bool(torch.isfinite(out.float()).all())
    assert (out.float() - ref).abs().max().item() <= 4e-2


@pytest.mark.parametrize("cross", [False, True])
def test_mixattn_large_model_shapes_and_guard_rows(built_lib, monkeypatch, cross):
    """MixViT-L / ConvMAE-L geometry (16 heads, 288 template + 576 search tokens: query tiles of 32 and 64 rows at the tails,
    a 32-key block in the MIDDLE of the key list) - the tail warps take the direct-store path, full 32-row groups the TMA
    store, warps without a valid row only keep the barrier protocol.  Guard rows after the last (partial) tile stay untouched."""
    import sys
    from mmt_b200 import ops
    me = sys.modules[__name__]
    monkeypatch.setattr(me, "HEADS", 16)
    monkeypatch.setattr(me, "C", 16 * HD)
    heads, c = 16, 16 * HD
    nseq, Lt, Ls = 2, 288, 576
    N = Lt + Ls
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = (torch.randn(nseq * N, 3 * c, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    tiles, segs = _tiles(nseq, N, Lt, Ls, cross)
    max_keys = max(sum(l for _, l in s[2]) for s in segs)
    guard = 96
    buf = torch.full((nseq * N + guard, c), float("nan"), device="cuda", dtype=torch.bfloat16)
    out = buf[: nseq * N]
    ops.mixattn(qkv, None, c, heads, tiles.cuda(), max_keys, out, HD ** -0.5)
    torch.cuda.synchronize()
    ref = _reference(qkv, segs, HD ** -0.5)
    assert bool(torch.isfinite(out.float()).all()), "unwritten or non-finite output rows"
    assert bool(torch.isnan(buf[nseq * N:].float()).all()), "rows past the last tile were written"
    assert (out.float() - ref).abs().max().item() <= 2.5e-2


def test_mixattn_partial_tiles_do_not_touch_neighbour_rows(built_lib):
    """A tile table that covers only SOME query tiles (every second one): rows of the tiles that are not in the table must
    keep their previous contents - a 32 x 32 TMA store box or a straddling warp may not spill into them."""
    from mmt_b200 import ops
    nseq, Lt, Ls = 2, 128, 324
    N = Lt + Ls
    g = torch.Generator(device="cuda").manual_seed(9)
    qkv = (torch.randn(nseq * N, 3 * C, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    tiles, segs = _tiles(nseq, N, Lt, Ls, False)
    keep = list(range(0, tiles.shape[0], 2)) + [tiles.shape[0] - 1]       # incl. the 68-row tail tile of the last sequence
    keep = sorted(set(keep))
    sub = tiles[keep].contiguous()
    out = torch.full((nseq * N, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.mixattn(qkv, None, C, HEADS, sub.cuda(), N, out, HD ** -0.5)
    torch.cuda.synchronize()
    ref = _reference(qkv, [segs[i] for i in keep], HD ** -0.5)
    written = torch.zeros(nseq * N, dtype=torch.bool, device="cuda")
    for i in keep:
        q0, qn, _ = segs[i]
        written[q0:q0 + qn] = True
    assert bool(torch.isnan(out[~written].float()).all()), "rows of tiles that were not launched were written"
    assert (out[written].float() - ref[written]).abs().max().item() <= 2.5e-2
