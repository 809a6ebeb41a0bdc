"""TEST INFRASTRUCTURE - CPU oracle of the tracker forward.  NOT part of the product.

A functional, fp32, torch-CPU restatement of the reference's per-frame network forward, operating directly on
a reference-layout state_dict.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  oracle/gen_golden.py runs the UNMODIFIED reference modules (imported from
/root/reference through oracle/ref_shims.py, in the build container) and this restatement on the same seeded
weights and inputs, asserts agreement, and writes the reference's outputs to tests/golden/*.npz; the
`-m "not gpu"` tests re-check this module against those committed vectors.  The two native ops are
additionally pinned by the reference's own known-answer tests (MSDA vs grid_sample,
deformable_attention/ops/test.py:31-60; PrRoIPool vs avg_pool2d, external/PreciseRoIPooling/pytorch/tests/
test_prroi_pooling2d.py:21-35), restated in tests/test_oracle_native_ops.py.

All paths cited below are relative to the reference root.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------- a1
def sincos_pos_embed_2d(embed_dim: int, grid_size: int) -> torch.Tensor:
    """lib/models/mixformer_vit/pos_utils.py:20-67 - fixed 2-D sin-cos table, w-major meshgrid."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])

    def one_d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one_d(embed_dim // 2, grid[0]), one_d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float()


def patch_embed(x, w, b):
    """PatchEmbed.forward lib/models/mixformer_vit/mixformer.py:28-33."""
    p = w.shape[-1]
    return F.conv2d(x, w, b, stride=p).flatten(2).transpose(1, 2).contiguous()


# ------------------------------------------------------------------------------------------------- a3/a4
def _heads(t, B, N, H):
    return t.reshape(B, N, 3, H, -1).permute(2, 0, 3, 1, 4)


def mixed_attention(x, p, heads, n_t, n_s):
    """Attention.forward lib/models/mixformer_vit/mixformer.py:51-77 (p: qkv.weight/bias, proj.weight/bias)."""
    B, N, C = x.shape
    scale = (C // heads) ** -0.5
    q, k, v = _heads(F.linear(x, p["qkv.weight"], p["qkv.bias"]), B, N, heads).unbind(0)
    q_mt, q_s = torch.split(q, [n_t, n_s], dim=2)
    k_mt, _ = torch.split(k, [n_t, n_s], dim=2)
    v_mt, _ = torch.split(v, [n_t, n_s], dim=2)
    a = ((q_mt @ k_mt.transpose(-2, -1)) * scale).softmax(dim=-1)
    x_mt = (a @ v_mt).transpose(1, 2).reshape(B, n_t, C)
    a = ((q_s @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    x_s = (a @ v).transpose(1, 2).reshape(B, n_s, C)
    return F.linear(torch.cat([x_mt, x_s], dim=1), p["proj.weight"], p["proj.bias"])


def cross_modal_attention(x_v, x_i, p, heads, n_t, n_s, return_attention=False):
    """Asym_Attention.forward lib/models/mixformer_vit_rgbt/asymmetric_shared_ce.py:146-207
    (== Attention.forward asymmetric_shared.py:55-104 when return_attention is False)."""
    B, N, C = x_v.shape
    scale = (C // heads) ** -0.5
    qkv = F.linear(torch.cat([x_v, x_i], dim=0), p["qkv.weight"], p["qkv.bias"]).reshape(2 * B, N, 3, heads, C // heads)
    qV, kV, vV = qkv[:B].permute(2, 0, 3, 1, 4).unbind(0)
    qI, kI, vI = qkv[B:].permute(2, 0, 3, 1, 4).unbind(0)
    sp = lambda t: torch.split(t, [n_t, n_s], dim=2)
    (q_mt_V, q_s_V), (k_mt_V, k_s_V), (v_mt_V, v_s_V) = sp(qV), sp(kV), sp(vV)
    (q_mt_I, q_s_I), (k_mt_I, k_s_I), (v_mt_I, v_s_I) = sp(qI), sp(kI), sp(vI)
    k_mt = torch.cat([k_mt_V, k_mt_I], dim=2)
    v_mt = torch.cat([v_mt_V, v_mt_I], dim=2)

    def att(q, k, v, n):
        a = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
        return (a @ v).transpose(1, 2).reshape(B, n, C)

    x_mt_V = att(q_mt_V, k_mt_V, v_mt_V, n_t)
    x_mt_I = att(q_mt_I, k_mt_I, v_mt_I, n_t)
    x_s_V = att(q_s_V, torch.cat([k_mt, k_s_V], dim=2), torch.cat([v_mt, v_s_V], dim=2), n_s)
    x_s_I = att(q_s_I, torch.cat([k_mt, k_s_I], dim=2), torch.cat([v_mt, v_s_I], dim=2), n_s)
    x = F.linear(torch.cat([torch.cat([x_mt_V, x_s_V], dim=1), torch.cat([x_mt_I, x_s_I], dim=1)], dim=0),
                 p["proj.weight"], p["proj.bias"])
    attn_t2s = None
    if return_attention:
        attn_t2s = ((torch.cat([q_mt_V, q_mt_I], dim=2) @ torch.cat([k_s_V, k_s_I], dim=2).transpose(-2, -1))
                    * scale).softmax(dim=-1)
    return x[:B], x[B:], attn_t2s


def mlp(x, p):
    """timm Mlp (fc1 -> erf GELU -> fc2), used at lib/models/mixformer_vit/mixformer.py:123."""
    return F.linear(F.gelu(F.linear(x, p["fc1.weight"], p["fc1.bias"])), p["fc2.weight"], p["fc2.bias"])


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _ln(x, p, name, eps):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


# ------------------------------------------------------------------------------------------------- backbones
def vit_dims(vit_type):
    if vit_type == "large_patch16":
        return dict(dim=1024, depth=24, heads=16)
    if vit_type == "base_patch16":
        return dict(dim=768, depth=12, heads=12)
    if vit_type == "convmae_base":       # lib/models/mixformer_convmae/mixformer_online.py:395-410
        return dict(dim=768, depth=11, heads=12, convmae=True)
    if vit_type == "convmae_large":
        return dict(dim=1024, depth=20, heads=16, convmae=True)
    raise KeyError("VIT_TYPE shoule set to 'large_patch16' or 'base_patch16'")


def _chan_ln(x, w, b, eps=1e-5):
    """nn.LayerNorm over channels of an NCHW map: norm(x.permute(0,2,3,1)).permute(0,3,1,2)."""
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def _conv_patch_embed(sd, name, x, stride):
    """PatchEmbed.forward lib/models/mixformer_convmae/mixformer_online.py:48-51: proj -> channel LN -> GELU."""
    x = F.conv2d(x, sd[name + ".proj.weight"], sd[name + ".proj.bias"], stride=stride)
    return F.gelu(_chan_ln(x, sd[name + ".norm.weight"], sd[name + ".norm.bias"]))


def _cblock(p, x):
    """CBlock.forward (mask=None) lib/models/mixformer_convmae/mixformer_online.py:181-189."""
    h = _chan_ln(x, p["norm1.weight"], p["norm1.bias"])
    h = F.conv2d(h, p["conv1.weight"], p["conv1.bias"])
    h = F.conv2d(h, p["attn.weight"], p["attn.bias"], padding=2, groups=h.shape[1])
    x = x + F.conv2d(h, p["conv2.weight"], p["conv2.bias"])
    h = _chan_ln(x, p["norm2.weight"], p["norm2.bias"])
    h = F.conv2d(F.gelu(F.conv2d(h, p["mlp.fc1.weight"], p["mlp.fc1.bias"])), p["mlp.fc2.weight"], p["mlp.fc2.bias"])
    return x + h


def convmae_stem(sd, img):
    """ConvViT.forward, one crop: lib/models/mixformer_convmae/mixformer_online.py:266-276 -> tokens [B, n, C]."""
    x = _conv_patch_embed(sd, "patch_embed1", img, 4)
    i = 0
    while f"blocks1.{i}.conv1.weight" in sd:
        x = _cblock(_sub(sd, f"blocks1.{i}."), x)
        i += 1
    x = _conv_patch_embed(sd, "patch_embed2", x, 2)
    i = 0
    while f"blocks2.{i}.conv1.weight" in sd:
        x = _cblock(_sub(sd, f"blocks2.{i}."), x)
        i += 1
    x = _conv_patch_embed(sd, "patch_embed3", x, 2)
    x = x.flatten(2).permute(0, 2, 1)
    return F.linear(x, sd["patch_embed4.weight"], sd["patch_embed4.bias"])


def _tokens_of(sd, img):
    """Token embedding of one crop WITHOUT the positional table: 16x16 patch conv (MixViT) or the ConvMAE stem."""
    if "patch_embed4.weight" in sd:
        return convmae_stem(sd, img)
    return patch_embed(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"])


def _block_prefix(sd):
    return "blocks3." if "patch_embed4.weight" in sd else "blocks."


def _embed_tokens(sd, x_t, x_ot, x_s):
    """VisionTransformer.forward lib/models/mixformer_vit/mixformer.py:192-203 (ConvViT.forward :266-311)."""
    t = _tokens_of(sd, x_t) + sd["pos_embed_t"]
    ot = _tokens_of(sd, x_ot) + sd["pos_embed_t"]
    s = _tokens_of(sd, x_s) + sd["pos_embed_s"]
    return torch.cat([t, ot, s], dim=1), t.shape[1] * 2, s.shape[1]


def backbone_plain(sd, x_t, x_ot, x_s, heads, depth, per_modality_ln=False):
    """Single-stream / batch-stacked backbone.  per_modality_ln=False: Block.forward
    lib/models/mixformer_vit/mixformer.py:126-129 (also mixformer_unibackbone.py:113-139);
    True: Block_Shared.forward lib/models/mixformer_vit_rgbt/mixformer_shared.py:143-159 (first half of the
    batch = RGB uses norm*_v, second half norm*_i).  Returns the search tokens [B, n_s, C] (no final norm)."""
    x, n_t, n_s = _embed_tokens(sd, x_t, x_ot, x_s)
    eps = 1e-6
    bp = _block_prefix(sd)
    for i in range(depth):
        p = _sub(sd, f"{bp}{i}.")
        a = _sub(p, "attn.")
        m = _sub(p, "mlp.")
        if per_modality_ln:
            n = x.shape[0] // 2
            h = torch.cat([_ln(x[:n], p, "norm1_v", eps), _ln(x[n:], p, "norm1_i", eps)], dim=0)
            x = x + mixed_attention(h, a, heads, n_t, n_s)
            h = torch.cat([_ln(x[:n], p, "norm2_v", eps), _ln(x[n:], p, "norm2_i", eps)], dim=0)
            x = x + mlp(h, m)
        else:
            x = x + mixed_attention(_ln(x, p, "norm1", eps), a, heads, n_t, n_s)
            x = x + mlp(_ln(x, p, "norm2", eps), m)
    return x[:, n_t:], x


def _forced_order(gi, forced):
    """Local positions (in the current global-index table `gi` [B, Ls]) of the tokens whose GLOBAL indices are listed
    in `forced` [B, keep], in the listed order, followed by the remaining local positions (ascending)."""
    B, Ls = forced.shape[0], gi.shape[1]     # gi may carry more rows than sequences (first stage: full-batch table)
    idx = torch.empty((B, Ls), dtype=torch.int64)
    for b in range(B):
        pos = {int(g): i for i, g in enumerate(gi[b].tolist())}
        top = [pos[int(g)] for g in forced[b].tolist()]      # KeyError: the forced token was removed at an earlier stage
        assert len(set(top)) == len(top), "forced keep list names a token twice"
        rest = sorted(set(range(Ls)) - set(top))
        idx[b] = torch.tensor(top + rest, dtype=torch.int64)
    return idx


def candidate_elimination(attn, tokens_v, tokens_i, keep_ratio, gidx_v, gidx_i, n_t, forced=None):
    """asymmetric_shared_ce.py:49-101 with box_mask_z=None (the test-time call,
    lib/test/tracker/asymmetric_shared_ce.py:96-98); get_token_from_attn :22-46.
    forced = (keep_v, keep_i) global indices [B, keep]: TEST AID ("forced-keep" mode) - the scores are computed as
    the reference does, but the kept set is the given one instead of the top-k of the scores, so that an
    implementation whose low-precision scores flipped near-ties across the keep boundary can be compared stage by
    stage on the SAME token population (tests/test_forward_gpu.py)."""
    lens_s = attn.shape[-1] // 2
    lens_keep = math.ceil(keep_ratio * lens_s)
    if lens_keep == lens_s:
        return tokens_v, tokens_i, gidx_v, gidx_i, None, None, None
    score = attn.mean(dim=2).mean(dim=1)
    outs = []
    for m, (sc, tok, gi) in enumerate(((score[:, :lens_s], tokens_v, gidx_v), (score[:, lens_s:], tokens_i, gidx_i))):
        if forced is None:
            _, idx = torch.sort(sc, dim=1, descending=True)
        else:
            assert forced[m].shape == (sc.shape[0], lens_keep), (forced[m].shape, lens_keep)
            idx = _forced_order(gi, forced[m])
        top, non = idx[:, :lens_keep], idx[:, lens_keep:]
        keep_i = gi.gather(1, top)
        rem_i = gi.gather(1, non)
        tt, ts = tok[:, :n_t], tok[:, n_t:]
        new = torch.cat([tt, ts.gather(1, top.unsqueeze(-1).expand(-1, -1, tok.shape[-1]))], dim=1)
        outs.append((new, keep_i, rem_i))
    return outs[0][0], outs[1][0], outs[0][1], outs[1][1], outs[0][2], outs[1][2], score


def backbone_asymmetric(sd, x_t, x_ot, x_s, heads, depth, ce_loc=None, ce_keep=None, forced_keep=None):
    """asymmetric_shared.py:137-154 blocks / asymmetric_shared_ce.py CE_Block_Shared.forward :247-282,
    VisionTransformer.forward :377-425, _recover_search :427-447.  Inputs are batch-stacked [RGB x B, TIR x B].
    Returns (search tokens [2B, n_s, C] in original positions, aux dict with CE scores / kept indices)."""
    x, n_t, n_s0 = _embed_tokens(sd, x_t, x_ot, x_s)
    B2 = x.shape[0]
    B = B2 // 2
    eps = 1e-6
    gidx_v = torch.linspace(0, n_s0 - 1, n_s0).repeat(B2, 1)
    gidx_i = gidx_v.clone()
    x_v, x_i = x[:B], x[B:]
    removed_v, removed_i = [], []
    aux = {"ce_scores": [], "ce_keep_v": [], "ce_keep_i": []}
    ce_idx = 0
    for i in range(depth):
        p = _sub(sd, f"blocks.{i}.")
        keep = 1.0
        if ce_loc is not None and i in ce_loc:
            keep = ce_keep[ce_idx]
            ce_idx += 1
        exe_ce = keep < 1
        n_s = gidx_v.shape[1]
        hv, hi = _ln(x_v, p, "norm1_v", eps), _ln(x_i, p, "norm1_i", eps)
        av, ai, attn_t2s = cross_modal_attention(hv, hi, _sub(p, "attn."), heads, n_t, n_s, exe_ce)
        x_v, x_i = x_v + av, x_i + ai
        if exe_ce:
            forced = None if forced_keep is None else forced_keep[len(aux["ce_scores"])]
            aux.setdefault("ce_gidx_in_v", []).append(gidx_v)       # token population the stage scored (global indices)
            aux.setdefault("ce_gidx_in_i", []).append(gidx_i)
            x_v, x_i, gidx_v, gidx_i, rv, ri, score = candidate_elimination(attn_t2s, x_v, x_i, keep, gidx_v, gidx_i, n_t,
                                                                            forced)
            aux["ce_scores"].append(score)
            aux["ce_keep_v"].append(gidx_v)
            aux["ce_keep_i"].append(gidx_i)
        else:
            rv = ri = None
        if ce_loc is not None and i in ce_loc:
            removed_v.append(rv)
            removed_i.append(ri)
        m = _sub(p, "mlp.")
        h = torch.cat([_ln(x_v, p, "norm2_v", eps), _ln(x_i, p, "norm2_i", eps)], dim=0)
        y = mlp(h, m)
        x_v, x_i = x_v + y[:B], x_i + y[B:]

    def recover(xm, removed, gidx):
        z, s = xm[:, :n_t], xm[:, n_t:]
        if removed and removed[0] is not None:
            rem = torch.cat(removed, dim=1)
            pad = torch.zeros(s.shape[0], n_s0 - s.shape[1], s.shape[2])
            s = torch.cat([s, pad], dim=1)
            index_all = torch.cat([gidx, rem], dim=1)
            s = torch.zeros_like(s).scatter_(1, index_all.unsqueeze(-1).expand(-1, -1, s.shape[-1]).to(torch.int64), s)
        return torch.cat([z, s], dim=1)

    # note the reference passes the full-batch index tensors; only the first B rows were ever gathered
    x_v = recover(x_v, removed_v, gidx_v)
    x_i = recover(x_i, removed_i, gidx_i)
    x = torch.cat([x_v, x_i], dim=0)
    aux["template_tokens"] = x[:, :n_t // 2]            # first template's tokens (asymmetric_shared_online.py:263-269)
    return x[:, n_t:], aux


# ------------------------------------------------------------------------------------------------- a7 fusion
def sine_position_embedding(B, H, W, num_pos_feats, temperature=10000.0):
    """PositionEmbeddingSine(normalize=True) deformable_attention/position_encoding.py:24-56 with an all-valid mask."""
    not_mask = torch.ones(B, H, W, dtype=torch.bool)
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = (y_embed - 0.5) / (y_embed[:, -1:, :] + eps) * scale
    x_embed = (x_embed - 0.5) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def msda_core(value, shapes, loc, attn):
    """MSDA forward semantics, restated per sample point from the CUDA kernel
    (deformable_attention/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, bilinear :33-84): pixel coords
    h = loc_y*H - 0.5, w = loc_x*W - 0.5; a sample contributes iff -1 < h < H and -1 < w < W; the four
    neighbours are weighted bilinearly with zeros outside the map.
    value [N,S,M,D], loc [N,Lq,M,L,P,2], attn [N,Lq,M,L,P] -> [N,Lq,M*D]."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = torch.zeros(N, Lq, M, D, dtype=value.dtype)
    start = 0
    n_idx = torch.arange(N).view(N, 1, 1, 1).expand(N, Lq, M, P)
    m_idx = torch.arange(M).view(1, 1, M, 1).expand(N, Lq, M, P)
    for l, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W].reshape(N, H, W, M, D)
        start += H * W
        h_im = loc[:, :, :, l, :, 1] * H - 0.5
        w_im = loc[:, :, :, l, :, 0] * W - 0.5
        ok = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low, w_low = torch.floor(h_im), torch.floor(w_im)
        lh, lw = h_im - h_low, w_im - w_low
        hh, hw = 1 - lh, 1 - lw
        h_low, w_low = h_low.long(), w_low.long()
        acc = torch.zeros(N, Lq, M, P, D, dtype=value.dtype)
        for dh, dw, wt in ((0, 0, hh * hw), (0, 1, hh * lw), (1, 0, lh * hw), (1, 1, lh * lw)):
            hi, wi = h_low + dh, w_low + dw
            inb = ok & (hi >= 0) & (hi <= H - 1) & (wi >= 0) & (wi <= W - 1)
            g = v[n_idx, hi.clamp(0, H - 1), wi.clamp(0, W - 1), m_idx]  # [N,Lq,M,P,D]
            acc = acc + g * (wt * inb).unsqueeze(-1)
        out = out + (acc * attn[:, :, :, l, :].unsqueeze(-1)).sum(dim=3)
    return out.reshape(N, Lq, M * D)


def msdeform_attn_bimodal(query, ref, src, shapes, p, n_heads=8, n_points=4):
    """MSDeformAttn_Bimodal.forward deformable_attention/ops/modules/ms_deform_attn_bimodal.py:83-130."""
    N, Lq, C = query.shape
    L = 2
    qv, qi = torch.chunk(query, 2, 1)
    qb = torch.cat([qv, qi], dim=2)
    value = F.linear(src, p["value_proj.weight"], p["value_proj.bias"]).view(N, -1, n_heads, C // n_heads)
    off = F.linear(qb, p["sampling_offsets.weight"], p["sampling_offsets.bias"]).view(N, Lq // 2, n_heads, L, n_points, 2)
    off = torch.cat([off, off], dim=1)
    aw = F.linear(qb, p["attention_weights.weight"], p["attention_weights.bias"]).view(N, Lq // 2, n_heads, L * n_points)
    aw = torch.cat([aw, aw], dim=1)
    aw = F.softmax(aw, -1).view(N, Lq, n_heads, L, n_points)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    out = msda_core(value, shapes, loc, aw)
    return F.linear(out, p["output_proj.weight"], p["output_proj.bias"])


def fusion_encoder_lnspecific(sd, src_v, src_i):
    """DeformableAttentionFusion_LNSpecific.forward deformable_attention/deformable_encoder_lnspecific.py:70-108,
    encoder layer :143-160, reference points :167-184.  sd keys are relative to `fusion_attention.`."""
    B, C, H, W = src_v.shape
    pos = sine_position_embedding(B, H, W, C // 2).flatten(2).transpose(1, 2)
    lvl = sd["level_embed"]
    src = torch.cat([src_v.flatten(2).transpose(1, 2), src_i.flatten(2).transpose(1, 2)], dim=1)
    pos = torch.cat([pos + lvl[0].view(1, 1, -1), pos + lvl[1].view(1, 1, -1)], dim=1)
    shapes = [(H, W), (H, W)]
    ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
    r = torch.stack((rx.reshape(-1)[None] / W, ry.reshape(-1)[None] / H), -1)   # valid ratios are 1
    ref = torch.cat([r, r], dim=1)[:, :, None].expand(B, -1, 2, -1)              # [B, 2HW, L=2, 2]
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layers."))
    for i in range(n_layers):
        p = _sub(sd, f"encoder.layers.{i}.")
        a = msdeform_attn_bimodal(src + pos, ref, src, shapes, _sub(p, "self_attn."))
        src = src + a
        spec = "norm1_v.weight" in p     # LN-specific encoder; else DeformableTransformerEncoderLayer deformable_encoder.py:143-158

        def norm(t, name):
            if not spec:
                return _ln(t, p, name, 1e-5)
            tv, ti = torch.chunk(t, 2, 1)
            return torch.cat([_ln(tv, p, name + "_v", 1e-5), _ln(ti, p, name + "_i", 1e-5)], dim=1)
        src = norm(src, "norm1")
        y = F.linear(F.relu(F.linear(src, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"], p["linear2.bias"])
        src = norm(src + y, "norm2")
    return src


def _conv_gn(x, p, name):
    x = F.conv2d(x, p[name + ".0.weight"], p[name + ".0.bias"])
    return F.group_norm(x, 32, p[name + ".1.weight"], p[name + ".1.bias"], 1e-5)


def fusion_vi(sd, search_v, search_i, fusion_class):
    """Attention_Fusion_Bimodal_LNSpecific{,_Sum,_2}.forward lib/models/mixformer_vit_rgbt/fusion_utils.py:270-279,
    :309-318, :344-353.  sd keys relative to `fusion_vi.`."""
    b, c, h, w = search_v.shape
    if fusion_class == "RGBT_Fusion_Cat":      # fusion_utils.py:105-110 (eval BatchNorm2d, eps 1e-5)
        out = torch.cat([search_v, search_i], dim=1)
        for j in (1, 2, 3):
            out = F.conv2d(out, sd[f"fusion{j}.weight"], None, padding=1)
            out = F.relu(F.batch_norm(out, sd[f"fusion{j}_bn.running_mean"], sd[f"fusion{j}_bn.running_var"],
                                      sd[f"fusion{j}_bn.weight"], sd[f"fusion{j}_bn.bias"], False, 0.0, 1e-5))
        return out
    if fusion_class == "Attention_Fusion_Bimodal_LNSpecific_2":
        iv, ii = _conv_gn(search_v, sd, "adjust_in"), _conv_gn(search_i, sd, "adjust_in")
    else:
        iv, ii = _conv_gn(search_v, sd, "adjust_v"), _conv_gn(search_i, sd, "adjust_i")
    out = fusion_encoder_lnspecific(_sub(sd, "fusion_attention."), iv, ii)
    ov, oi = torch.chunk(out, 2, 1)
    if fusion_class in ("Attention_Fusion_Bimodal_LNSpecific", "Attention_Fusion_Bimodal"):
        ov = ov.permute(0, 2, 1).reshape(b, -1, h, w)
        oi = oi.permute(0, 2, 1).reshape(b, -1, h, w)
        return _conv_gn(torch.cat([ov, oi], dim=1), sd, "adjust_cat")
    o = (ov + oi).permute(0, 2, 1).reshape(b, -1, h, w)
    if fusion_class == "Attention_Fusion_Bimodal_LNSpecific_Sum":
        return _conv_gn(o, sd, "adjust_sum")
    if fusion_class == "Attention_Fusion_Bimodal_LNSpecific_2":
        return _conv_gn(o, sd, "adjust_out")
    raise KeyError(f"fusion class {fusion_class} is not on the accelerated path")


# ------------------------------------------------------------------------------------------------- a8/a9 head
def _conv_bn_relu(x, p, name):
    """conv() lib/models/mixformer_cvt/head.py:7-20 in eval mode.  BatchNorm2d (running stats, eps 1e-5) and
    FrozenBatchNorm2d (lib/models/mixformer_cvt/utils.py:47-57) are the same affine map."""
    x = F.conv2d(x, p[name + ".0.weight"], p[name + ".0.bias"], padding=1)
    scale = p[name + ".1.weight"] * (p[name + ".1.running_var"] + 1e-5).rsqrt()
    bias = p[name + ".1.bias"] - p[name + ".1.running_mean"] * scale
    return F.relu(x * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))


def corner_head_pyramid(sd, x, feat_sz, stride):
    """Pyramid_Corner_Predictor.forward/get_score_map/soft_argmax lib/models/mixformer_cvt/head.py:147-212.
    Returns (xyxy / img_sz [B,4], score maps [B,2,feat_sz*feat_sz])."""
    maps = []
    for c in ("tl", "br"):
        x1 = _conv_bn_relu(x, sd, f"conv1_{c}")
        x2 = _conv_bn_relu(x1, sd, f"conv2_{c}")
        up1 = F.interpolate(_conv_bn_relu(x, sd, f"adjust1_{c}"), scale_factor=2) + F.interpolate(x2, scale_factor=2)
        x3 = _conv_bn_relu(up1, sd, f"conv3_{c}")
        up2 = F.interpolate(_conv_bn_relu(x, sd, f"adjust2_{c}"), scale_factor=4) + F.interpolate(x3, scale_factor=2)
        x4 = _conv_bn_relu(up2, sd, f"conv4_{c}")
        a3 = x2
        for j in range(3):
            a3 = _conv_bn_relu(a3, sd, f"adjust3_{c}.{j}")
        a4 = x3
        for j in range(2):
            a4 = _conv_bn_relu(a4, sd, f"adjust4_{c}.{j}")
        score = F.conv2d(x4, sd[f"conv5_{c}.weight"], sd[f"conv5_{c}.bias"]) + F.interpolate(a3, scale_factor=4) + \
            F.interpolate(a4, scale_factor=2)
        maps.append(score.reshape(score.shape[0], -1))
    return _soft_argmax_pair(maps, feat_sz, stride), torch.stack(maps, dim=1)


def corner_head_plain(sd, x, feat_sz, stride):
    """Corner_Predictor lib/models/mixformer_cvt/head.py:54-94 (HEAD_TYPE == CORNER)."""
    maps = []
    for c in ("tl", "br"):
        y = x
        for j in range(1, 5):
            y = _conv_bn_relu(y, sd, f"conv{j}_{c}")
        s = F.conv2d(y, sd[f"conv5_{c}.weight"], sd[f"conv5_{c}.bias"])
        maps.append(s.reshape(s.shape[0], -1))
    return _soft_argmax_pair(maps, feat_sz, stride), torch.stack(maps, dim=1)


def _soft_argmax_pair(maps, feat_sz, stride):
    idx = torch.arange(0, feat_sz).view(-1, 1) * stride
    coord_x = idx.repeat((feat_sz, 1)).view(-1).float()
    coord_y = idx.repeat((1, feat_sz)).view(-1).float()
    out = []
    for m in maps:
        prob = F.softmax(m, dim=1)
        out += [torch.sum(coord_x * prob, dim=1), torch.sum(coord_y * prob, dim=1)]
    return torch.stack(out, dim=1) / (feat_sz * stride)


def box_xyxy_to_cxcywh(x):
    """lib/utils/box_ops.py:27-31."""
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1)


def box_head(sd, feat, cfg):
    """build_box_head lib/models/mixformer_cvt/head.py:235-258 + forward_box_head mixformer.py:325-338."""
    size = cfg["search_size"]
    if cfg["head_type"] == "CORNER_UP":
        xyxy, maps = corner_head_pyramid(sd, feat, size // 4, 4)
    elif cfg["head_type"] == "CORNER":
        xyxy, maps = corner_head_plain(sd, feat, size // 16, 16)
    else:
        raise ValueError("HEAD TYPE %s is not supported." % cfg["head_type"])
    return box_xyxy_to_cxcywh(xyxy).view(-1, 1, 4), maps


# ------------------------------------------------------------------------------------------------- a10 SPM
def box_cxcywh_to_xyxy(x):
    """lib/utils/box_ops.py:8-12."""
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def score_decoder(sd, search_feat, template_feat, search_box, num_heads):
    """ScoreDecoder.forward lib/models/mixformer_cvt/score_decoder.py:32-66.  search_feat [B,C,H,W], template_feat
    [B,C,Ht,Wt], search_box [B,4] normalised xyxy.  PrRoIPool (GPU-only in the reference) is the restatement in
    oracle/native_ops_oracle.py.  Returns raw logits [B]."""
    from oracle import native_ops_oracle as NO
    b, c, h, w = search_feat.shape
    scale = c ** -0.5                                         # hidden_dim ** -0.5, NOT head_dim (score_decoder.py:18)
    bb = search_box.clone().view(-1, 4) * w
    rois = torch.cat([torch.arange(b, dtype=torch.float32).view(-1, 1), bb], dim=1)
    pooled = torch.from_numpy(NO.prroi_pool_forward(search_feat.numpy(), rois.numpy(), 4, 4, 1.0))   # [B,C,4,4]
    x = sd["score_token"].expand(b, -1, -1)
    x = F.layer_norm(x, (c,), sd["norm1.weight"], sd["norm1.bias"], 1e-5)
    mem = [pooled.flatten(2).transpose(1, 2), template_feat.flatten(2).transpose(1, 2)]
    hd = c // num_heads
    for i in range(2):
        q = F.linear(x, sd[f"proj_q.{i}.weight"], sd[f"proj_q.{i}.bias"]).view(b, -1, num_heads, hd).transpose(1, 2)
        k = F.linear(mem[i], sd[f"proj_k.{i}.weight"], sd[f"proj_k.{i}.bias"]).view(b, -1, num_heads, hd).transpose(1, 2)
        v = F.linear(mem[i], sd[f"proj_v.{i}.weight"], sd[f"proj_v.{i}.bias"]).view(b, -1, num_heads, hd).transpose(1, 2)
        a = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
        x = (a @ v).transpose(1, 2).reshape(b, -1, c)
        x = F.linear(x, sd[f"proj.{i}.weight"], sd[f"proj.{i}.bias"])
        x = F.layer_norm(x, (c,), sd[f"norm2.{i}.weight"], sd[f"norm2.{i}.bias"], 1e-5)
    n = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("score_head.layers."))
    for i in range(n):                                         # MLP head.py:229-232
        x = F.linear(x, sd[f"score_head.layers.{i}.weight"], sd[f"score_head.layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x.view(-1)


def forward_head_online(sd, search_feat, template_feat, mc, heads, run_score_head=True, gt_bboxes=None):
    """MixFormerOnlineScore.forward_head lib/models/mixformer_vit/mixformer_online.py:326-343."""
    boxes, maps = box_head(_sub(sd, "box_head."), search_feat, mc)
    out = dict(pred_boxes=boxes, score_maps=maps, feat=search_feat)
    if run_score_head:
        bb = gt_bboxes if gt_bboxes is not None else box_cxcywh_to_xyxy(boxes.clone().view(-1, 4))
        out["pred_scores"] = score_decoder(_sub(sd, "score_branch."), search_feat, template_feat, bb, heads)
    return out


class OnlineState:
    """Cached-template state of one sequence (Attention.set_online / VisionTransformer.set_online,
    lib/models/mixformer_vit/mixformer_online.py:96-113,243-262): per-layer qkv of the template tokens and the
    final feature of the first template."""

    def __init__(self):
        self.qkv_mem = []
        self.template = None


def online_set(sd, cfg, template, online_template):
    """model.set_online(template [1,3,T,T], online_template [n,3,T,T])."""
    mc = cfg if "variant" in cfg else model_cfg("mixformer_vit_online", cfg)
    d = vit_dims(mc["vit_type"])
    bsd = _sub(sd, "backbone.")
    x_t = _tokens_of(bsd, template) + bsd["pos_embed_t"]
    x_ot = _tokens_of(bsd, online_template) + bsd["pos_embed_t"]
    x = torch.cat([x_t, x_ot.reshape(1, -1, x_ot.shape[-1])], dim=1)
    st = OnlineState()
    H, C = d["heads"], x.shape[-1]
    bp = _block_prefix(bsd)
    for i in range(d["depth"]):
        p = _sub(bsd, f"{bp}{i}.")
        a = _sub(p, "attn.")
        h = _ln(x, p, "norm1", 1e-6)
        B, N, _ = h.shape
        qkv = F.linear(h, a["qkv.weight"], a["qkv.bias"]).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        st.qkv_mem.append(qkv)
        q, k, v = qkv.unbind(0)
        att = ((q @ k.transpose(-2, -1)) * (C // H) ** -0.5).softmax(dim=-1)
        y = (att @ v).transpose(1, 2).reshape(B, N, C)
        x = x + F.linear(y, a["proj.weight"], a["proj.bias"])
        x = x + mlp(_ln(x, p, "norm2", 1e-6), _sub(p, "mlp."))
    g = mc["template_size"] // 16
    st.template = _tokens_to_map(x[:, :g * g], g)
    return st


def online_forward_test(sd, cfg, st, search, run_score_head=True):
    """model.forward_test(search [1,3,S,S]) against the cache (Attention.forward_test :80-94, backbone :229-241)."""
    mc = cfg if "variant" in cfg else model_cfg("mixformer_vit_online", cfg)
    d = vit_dims(mc["vit_type"])
    bsd = _sub(sd, "backbone.")
    x = _tokens_of(bsd, search) + bsd["pos_embed_s"]
    H, C = d["heads"], x.shape[-1]
    bp = _block_prefix(bsd)
    for i in range(d["depth"]):
        p = _sub(bsd, f"{bp}{i}.")
        a = _sub(p, "attn.")
        h = _ln(x, p, "norm1", 1e-6)
        B, N, _ = h.shape
        qkv_s = F.linear(h, a["qkv.weight"], a["qkv.bias"]).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
        q_s = qkv_s[0]
        _, k, v = torch.cat([st.qkv_mem[i], qkv_s], dim=3).unbind(0)
        att = ((q_s @ k.transpose(-2, -1)) * (C // H) ** -0.5).softmax(dim=-1)
        y = (att @ v).transpose(1, 2).reshape(B, N, C)
        x = x + F.linear(y, a["proj.weight"], a["proj.bias"])
        x = x + mlp(_ln(x, p, "norm2", 1e-6), _sub(p, "mlp."))
    g = mc["search_size"] // 16
    return forward_head_online(sd, _tokens_to_map(x, g), st.template, mc, d["heads"], run_score_head)


# ------------------------------------------------------------------------------------------------- whole forward
def model_cfg(variant, cfg):
    """The handful of config keys the forward depends on, from a reference-style cfg tree."""
    m = cfg["MODEL"]
    out = dict(variant=variant, vit_type=m["VIT_TYPE"], head_type=m["HEAD_TYPE"], hidden_dim=m["HIDDEN_DIM"],
               search_size=cfg["DATA"]["SEARCH"]["SIZE"], template_size=cfg["DATA"]["TEMPLATE"]["SIZE"],
               fusion_class=m.get("FUSION_CLASS"), fusion_layers=m.get("FUSION_LAYERS"))
    bb = m.get("BACKBONE", {})
    out["ce_loc"] = list(bb["CE_LOC"]) if "CE_LOC" in bb else None
    out["ce_keep"] = list(bb["CE_KEEP_RATIO"]) if "CE_KEEP_RATIO" in bb else None
    return out


def _tokens_to_map(tok, g):
    B, n, C = tok.shape
    return tok.transpose(1, 2).reshape(B, C, g, g)


@torch.no_grad()
def forward(variant, sd, cfg, template, online_template, search, forced_keep=None):
    """The reference `model(template, online_template, search)` for the non-online variants.
    RGB-T inputs are 2-lists [v, i].  Returns dict(pred_boxes [B,1,4], score_maps [B,2,S*S], + aux).
    forced_keep (asymmetric_shared_ce only, test aid): per CE stage a (keep_v, keep_i) pair of global-index tensors
    [B, keep] that replaces the stage's top-k selection (see candidate_elimination)."""
    mc = cfg if "variant" in cfg else model_cfg(variant, cfg)
    d = vit_dims(mc["vit_type"])
    g = mc["search_size"] // 16
    aux = {}
    if variant in ("mixformer_vit_online", "mixformer_convmae_online"):   # mixformer_online.py:297-311 (+ SPM)
        s, x = backbone_plain(_sub(sd, "backbone."), template, online_template, search, d["heads"], d["depth"])
        gt = mc["template_size"] // 16
        return forward_head_online(sd, _tokens_to_map(s, g), _tokens_to_map(x[:, :gt * gt], gt), mc, d["heads"])
    if variant == "mixformer_vit":          # lib/models/mixformer_vit/mixformer.py:294-306
        s, _ = backbone_plain(_sub(sd, "backbone."), template, online_template, search, d["heads"], d["depth"])
        feat = _tokens_to_map(s, g)
    else:
        if variant == "mixformer_vit_rgbt":  # two-stream, lib/models/mixformer_vit_rgbt/mixformer.py:366-395
            sv, _ = backbone_plain(_sub(sd, "backbone_v."), template[0], online_template[0], search[0], d["heads"], d["depth"])
            si, _ = backbone_plain(_sub(sd, "backbone_i."), template[1], online_template[1], search[1], d["heads"], d["depth"])
        else:                                 # batch-stacked, mixformer_shared.py:400-424
            t = torch.cat(template, dim=0)
            ot = torch.cat(online_template, dim=0)
            sr = torch.cat(search, dim=0)
            bsd = _sub(sd, "backbone.")
            if variant == "mixformer_vit_rgbt_shared":
                s, _ = backbone_plain(bsd, t, ot, sr, d["heads"], d["depth"], per_modality_ln=True)
            elif variant == "mixformer_vit_rgbt_unibackbone":
                s, _ = backbone_plain(bsd, t, ot, sr, d["heads"], d["depth"])
            elif variant in ("asymmetric_shared", "asymmetric_shared_online"):
                s, aux = backbone_asymmetric(bsd, t, ot, sr, d["heads"], d["depth"])
            elif variant == "asymmetric_shared_ce":
                s, aux = backbone_asymmetric(bsd, t, ot, sr, d["heads"], d["depth"], mc["ce_loc"], mc["ce_keep"],
                                             forced_keep)
            else:
                raise KeyError(variant)
            n = s.shape[0] // 2
            sv, si = s[:n], s[n:]
        fv, fi = _tokens_to_map(sv, g).contiguous(), _tokens_to_map(si, g).contiguous()
        feat = fusion_vi(_sub(sd, "fusion_vi."), fv, fi, mc["fusion_class"])
        aux["search_v"], aux["search_i"] = fv, fi
    templ_tok = aux.pop("template_tokens", None)
    if variant == "asymmetric_shared_online":
        # MixFormer_RGBT_OnlineScore.forward lib/models/mixformer_vit_rgbt/asymmetric_shared_online.py:352-373: the SPM sees
        # the FUSED search map and both modalities' first-template maps stacked along H (cat(split(template), dim=2))
        gt = mc["template_size"] // 16
        n = templ_tok.shape[0] // 2
        tmap = torch.cat([_tokens_to_map(templ_tok[:n], gt), _tokens_to_map(templ_tok[n:], gt)], dim=2)
        out = forward_head_online(sd, feat, tmap, mc, d["heads"])
        out.update(aux)
        return out
    boxes, maps = box_head(_sub(sd, "box_head."), feat, mc)
    return dict(pred_boxes=boxes, score_maps=maps, feat=feat, **aux)
