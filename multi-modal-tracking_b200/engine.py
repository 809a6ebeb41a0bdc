"""Forward engine: packs a reference-layout state_dict into a device weight arena and runs the per-frame
tracker forward as a sequence of C-ABI kernel launches (ops.py) on torch's current CUDA stream.

Data layout in HBM (B = sequences in the batched step, N = tokens per sequence-modality):
  * token tensors are row-major [rows, C]; the residual stream `x` is fp32 [nseq*N, C] with
    nseq = B (RGB-only) or 2B (RGB-T, modality-major: RGB sequences first, TIR second - the reference's
    torch.cat([v, i], dim=0), mixformer_shared.py:414-416); inside a sequence the order is
    [template(64) | online template(64) | search(324)] (mixformer.py:202);
  * GEMM inputs are `act` dtype = bf16 (fast mode) or fp32 (parity mode); LayerNorm/GroupNorm statistics, the
    residual stream, softmax state, sampling offsets and the corner soft-argmax are always fp32;
  * feature maps are NHWC (== token rows); conv weights are packed [O, (ky, kx, c)] with eval-BatchNorm folded;
  * all workspaces are allocated once per batch size and reused (no allocation inside a step);
  * bf16 mode: the blocks' LayerNorms are folded into the GEMMs around them (_block_folded) - the residual stream has a
    bf16 shadow copy `xb` and per-row partial sums written by the proj / fc2 epilogues, the normalised rows never exist.

No torch compute op is on the path: torch provides memory (torch.empty) and the stream only.
"""
from __future__ import annotations

import math

import torch

from . import ops
from .pos_embed import fusion_pos_table

VIT_DIMS = {"base_patch16": dict(dim=768, depth=12, heads=12), "large_patch16": dict(dim=1024, depth=24, heads=16),
            # ConvMAE (lib/models/mixformer_convmae/mixformer_online.py:395-410): conv stem (4,2,2) then MixViT blocks
            "convmae_base": dict(dim=768, depth=11, heads=12, stem=(256, 384)),
            "convmae_large": dict(dim=1024, depth=20, heads=16, stem=(384, 768))}
STACKED = ("mixformer_vit_rgbt_shared", "mixformer_vit_rgbt_unibackbone", "asymmetric_shared", "asymmetric_shared_ce",
           "asymmetric_shared_online")
CROSS_MODAL = ("asymmetric_shared", "asymmetric_shared_ce", "asymmetric_shared_online")
PER_MODALITY_LN = ("mixformer_vit_rgbt_shared", "asymmetric_shared", "asymmetric_shared_ce", "asymmetric_shared_online")


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class ForwardEngine:
    def __init__(self, variant, cfg, state_dict, device, precision="bf16"):
        self.variant = variant
        self.dev = device
        self.bf16 = precision == "bf16"
        self.act = torch.bfloat16 if self.bf16 else torch.float32
        m = cfg["MODEL"]
        d = VIT_DIMS[m["VIT_TYPE"]]
        self.vit_type = m["VIT_TYPE"]
        self.dim, self.depth, self.heads = d["dim"], d["depth"], d["heads"]
        self.stem_dims = d.get("stem")           # ConvMAE: channel widths of the two conv stages
        self.block_prefix = "blocks3." if self.stem_dims else "blocks."
        self.search_size = int(cfg["DATA"]["SEARCH"]["SIZE"])
        self.template_size = int(cfg["DATA"]["TEMPLATE"]["SIZE"])
        self.gs, self.gt = self.search_size // 16, self.template_size // 16
        self.Ls0, self.Lt = self.gs * self.gs, 2 * self.gt * self.gt
        self.N0 = self.Lt + self.Ls0
        self.head_type = m["HEAD_TYPE"]
        if self.head_type not in ("CORNER_UP", "CORNER"):
            raise NotImplementedError(f"HEAD_TYPE {self.head_type!r}: only the corner heads (CORNER_UP = pyramid, used by "
                                      "every shipped MixViT / ConvMAE YAML, and the plain CORNER) are on the accelerated path")
        self.rgbt = variant not in ("mixformer_vit", "mixformer_vit_online", "mixformer_convmae_online")
        self.fusion_class = m.get("FUSION_CLASS") if self.rgbt else None
        bb = m.get("BACKBONE", {})
        self.ce_loc = list(bb["CE_LOC"]) if (variant == "asymmetric_shared_ce" and "CE_LOC" in bb) else []
        self.ce_keep = list(bb["CE_KEEP_RATIO"]) if self.ce_loc else []
        self.scale = (self.dim // self.heads) ** -0.5
        # LayerNorm folded into the GEMMs around it (bf16 tensor-core path; _block): the normalised rows never exist in HBM.
        # MMT_LN_FOLD=0 keeps the stand-alone LayerNorm kernel (A/B measurements); the fp32 parity mode always does.
        import os
        self.ln_fold = self.bf16 and os.environ.get("MMT_LN_FOLD", "1") != "0"
        self._ws = {}
        self._tiles = {}
        self.aux = {}
        self._pack(state_dict)

    # ------------------------------------------------------------------------------------------ packing
    def _w(self, t):
        return t.detach().to(device=self.dev, dtype=self.act).contiguous()

    def _pack_backbone(self, sd, prefix, per_modality_ln):
        dev = self.dev
        g = lambda k: sd[prefix + k]
        bb = {}
        if self.stem_dims:      # the token-embedding GEMM is patch_embed4 (Linear) on the conv stem's output
            bb["pe_w"] = self._w(g("patch_embed4.weight"))
            bb["pe_b"] = _f32(g("patch_embed4.bias"), dev)
        else:
            bb["pe_w"] = self._w(g("patch_embed.proj.weight").reshape(self.dim, -1))
            bb["pe_b"] = _f32(g("patch_embed.proj.bias"), dev)
        pt, ps = g("pos_embed_t")[0], g("pos_embed_s")[0]
        bb["pos"] = _f32(torch.cat([pt, pt, ps], dim=0), dev)          # [N0, dim]
        blocks = []
        for i in range(self.depth):
            p = f"{self.block_prefix}{i}."
            b = {}
            for j in (1, 2):
                if per_modality_ln:
                    b[f"ln{j}"] = (_f32(g(p + f"norm{j}_v.weight"), dev), _f32(g(p + f"norm{j}_v.bias"), dev),
                                   _f32(g(p + f"norm{j}_i.weight"), dev), _f32(g(p + f"norm{j}_i.bias"), dev))
                else:
                    b[f"ln{j}"] = (_f32(g(p + f"norm{j}.weight"), dev), _f32(g(p + f"norm{j}.bias"), dev), None, None)
            for name, key in (("qkv", "attn.qkv"), ("proj", "attn.proj"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
                folded = self.ln_fold and name in ("qkv", "fc1")
                if not folded:
                    b[name + "_w"] = self._w(g(p + key + ".weight"))
                    b[name + "_b"] = _f32(g(p + key + ".bias"), dev)
                    continue
                # Linear(LN(x)) = rs * (x W'^T - mu * colsum) + bias'  with  W' = W diag(gamma), colsum = rowsum(bf16 W'),
                # bias' = bias + W beta (one set per modality-specific norm): include/mmt_b200.h, mmt_gemm_bf16_ex
                W = g(p + key + ".weight").detach().float().cpu()
                bias = g(p + key + ".bias").detach().float().cpu()
                ln = b["ln1" if name == "qkv" else "ln2"]
                sets = []
                for s_ in range(2 if ln[2] is not None else 1):
                    gam, bet = ln[2 * s_].float().cpu(), ln[2 * s_ + 1].float().cpu()
                    Wf = (W * gam[None, :]).to(torch.bfloat16)
                    sets.append((Wf.to(dev).contiguous(), _f32(bias + W @ bet, dev), _f32(Wf.float().sum(dim=1), dev)))
                b[name + "_f"] = sets
            blocks.append(b)
        bb["blocks"] = blocks
        return bb

    @staticmethod
    def _fold_bn(sd, name):
        """conv3x3 + eval BatchNorm2d / FrozenBatchNorm2d -> (W[O, (ky,kx,c)], b[O]) fp32
        (head.py:7-20; utils.py:47-57: scale = w * rsqrt(var + 1e-5))."""
        w = sd[name + ".0.weight"].detach().float()
        b = sd[name + ".0.bias"].detach().float()
        scale = sd[name + ".1.weight"].detach().float() * (sd[name + ".1.running_var"].detach().float() + 1e-5).rsqrt()
        shift = sd[name + ".1.bias"].detach().float() - sd[name + ".1.running_mean"].detach().float() * scale
        wp = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
        return wp, b * scale + shift

    def _pack_head(self, sd):
        dev = self.dev
        H = {}
        fold = lambda n: self._fold_bn(sd, "box_head." + n)
        ws, bs, self.s1_cols = [], [], {}
        col = 0
        self.head_ch = sd["box_head.conv1_tl.0.weight"].shape[0]
        if self.head_type == "CORNER":       # Corner_Predictor head.py:23-94: conv1..conv4 (3x3 + BN + ReLU) + conv5 (1x1) at stride 16
            for n in ("conv1_tl", "conv1_br"):
                w, b = fold(n)
                self.s1_cols[n] = (col, col + w.shape[0])
                col += w.shape[0]
                ws.append(w)
                bs.append(b)
            H["s1_w"], H["s1_b"] = self._w(torch.cat(ws, 0)), _f32(torch.cat(bs, 0), dev)
            self.s1_width = col
            for c in ("tl", "br"):
                for n in (f"conv2_{c}", f"conv3_{c}", f"conv4_{c}"):
                    w, b = fold(n)
                    H[n + "_w"], H[n + "_b"] = self._w(w), _f32(b, dev)
                H[f"w5_{c}"] = _f32(sd[f"box_head.conv5_{c}.weight"].reshape(-1), dev)
                H[f"b5_{c}"] = float(sd[f"box_head.conv5_{c}.bias"].detach().float().reshape(-1)[0])
            return H
        for n in ("conv1_tl", "conv1_br", "adjust1_tl", "adjust1_br", "adjust2_tl", "adjust2_br"):
            w, b = fold(n)
            self.s1_cols[n] = (col, col + w.shape[0])
            col += w.shape[0]
            ws.append(w)
            bs.append(b)
        H["s1_w"], H["s1_b"] = self._w(torch.cat(ws, 0)), _f32(torch.cat(bs, 0), dev)
        self.s1_width = col
        for c in ("tl", "br"):
            for n in (f"conv2_{c}", f"conv3_{c}", f"conv4_{c}", f"adjust3_{c}.0", f"adjust3_{c}.1", f"adjust3_{c}.2",
                      f"adjust4_{c}.0", f"adjust4_{c}.1"):
                w, b = fold(n)
                H[n + "_w"], H[n + "_b"] = self._w(w), _f32(b, dev)
            H[f"w5_{c}"] = _f32(sd[f"box_head.conv5_{c}.weight"].reshape(-1), dev)
            H[f"b5_{c}"] = float(sd[f"box_head.conv5_{c}.bias"].detach().float().reshape(-1)[0])
        return H

    def _pack_fusion(self, sd):
        dev = self.dev
        F = {}
        g = lambda k: sd["fusion_vi." + k]
        cls = self.fusion_class
        from .builders import FUSION_CLASSES
        if cls not in FUSION_CLASSES:
            raise KeyError(f"FUSION_CLASS {cls!r} is not on the accelerated path")
        if cls == "RGBT_Fusion_Cat":          # three conv3x3 (no bias) + eval BatchNorm + ReLU, fusion_utils.py:86-110
            convs = []
            for j in (1, 2, 3):
                w = g(f"fusion{j}.weight").detach().float()
                scale = g(f"fusion{j}_bn.weight").detach().float() * (g(f"fusion{j}_bn.running_var").detach().float() + 1e-5).rsqrt()
                shift = g(f"fusion{j}_bn.bias").detach().float() - g(f"fusion{j}_bn.running_mean").detach().float() * scale
                wp = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
                convs.append((self._w(wp), _f32(shift, dev)))
            return {"cat_convs": convs}
        self.d_model = g("fusion_attention.level_embed").shape[1]
        names_in = ("adjust_in", "adjust_in") if cls.endswith("_2") else ("adjust_v", "adjust_i")
        F["in"] = [dict(w=self._w(g(n + ".0.weight").reshape(self.d_model, -1)), b=_f32(g(n + ".0.bias"), dev),
                        gn_w=_f32(g(n + ".1.weight"), dev), gn_b=_f32(g(n + ".1.bias"), dev)) for n in names_in]
        F["pos"] = fusion_pos_table(self.gs, self.gs, self.d_model, g("fusion_attention.level_embed")).to(dev)
        layers = []
        i = 0
        while f"fusion_vi.fusion_attention.encoder.layers.{i}.linear1.weight" in sd:
            p = f"fusion_attention.encoder.layers.{i}."
            L = {}
            L["offw_w"] = self._w(torch.cat([g(p + "self_attn.sampling_offsets.weight"),
                                             g(p + "self_attn.attention_weights.weight")], 0))
            L["offw_b"] = _f32(torch.cat([g(p + "self_attn.sampling_offsets.bias"),
                                          g(p + "self_attn.attention_weights.bias")], 0), dev)
            for nm, key in (("val", "self_attn.value_proj"), ("out", "self_attn.output_proj"), ("l1", "linear1"),
                            ("l2", "linear2")):
                L[nm + "_w"], L[nm + "_b"] = self._w(g(p + key + ".weight")), _f32(g(p + key + ".bias"), dev)
            for j in (1, 2):
                if ("fusion_vi." + p + f"norm{j}_v.weight") in sd:
                    L[f"ln{j}"] = tuple(_f32(g(p + f"norm{j}_{mm}.{wb}"), dev) for mm in ("v", "i") for wb in ("weight", "bias"))
                else:       # Attention_Fusion_Bimodal: one LayerNorm for both modalities (deformable_encoder.py:129,137)
                    L[f"ln{j}"] = (_f32(g(p + f"norm{j}.weight"), dev), _f32(g(p + f"norm{j}.bias"), dev), None, None)
            layers.append(L)
            i += 1
        F["layers"] = layers
        self.n_heads_f = 8
        self.n_points_f = L["offw_w"].shape[0] // (self.n_heads_f * 2 * 3)
        self.d_ffn = layers[0]["l1_w"].shape[0]
        if cls.endswith("_Sum") or cls.endswith("_2"):
            n = "adjust_sum" if cls.endswith("_Sum") else "adjust_out"
            w = g(n + ".0.weight").reshape(-1, self.d_model)
            w = torch.cat([w, w], dim=1)        # conv(out_v + out_i) == [out_v | out_i] @ [W | W]^T
        else:
            n = "adjust_cat"
            w = g(n + ".0.weight").reshape(-1, 2 * self.d_model)
        F["out"] = dict(w=self._w(w), b=_f32(g(n + ".0.bias"), dev), gn_w=_f32(g(n + ".1.weight"), dev),
                        gn_b=_f32(g(n + ".1.bias"), dev))
        return F

    def _pack(self, sd):
        if self.variant in ("mixformer_vit", "mixformer_vit_online", "mixformer_convmae_online"):
            self.bbs = [self._pack_backbone(sd, "backbone.", False)]
        elif self.variant == "mixformer_vit_rgbt":
            self.bbs = [self._pack_backbone(sd, "backbone_v.", False), self._pack_backbone(sd, "backbone_i.", False)]
        elif self.variant in STACKED:
            self.bbs = [self._pack_backbone(sd, "backbone.", self.variant in PER_MODALITY_LN)]
        else:
            raise KeyError(self.variant)
        self.head = self._pack_head(sd)
        self.fusion = self._pack_fusion(sd) if self.rgbt else None

    # ------------------------------------------------------------------------------------------ workspaces
    def _buf(self, B, name, shape, dtype):
        key = (B, name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(shape, device=self.dev, dtype=dtype)
            self._ws[key] = t
        return t

    def _attn_tiles(self, kind, nseq, N, Ls):
        """Host-built per-tile key-segment table (see mmt_mixattn_fwd).  kind: 'sym' or 'cross'."""
        key = (kind, nseq, N, Ls)
        hit = self._tiles.get(key)
        if hit is not None:
            return hit
        Lt = self.Lt
        recs = []

        def add(q0, qn, segs):
            for o in range(0, qn, 128):
                r = [q0 + o, min(128, qn - o), q0 + o, len(segs)]
                rows = [s[0] for s in segs] + [0] * (3 - len(segs))
                lens = [s[1] for s in segs] + [0] * (3 - len(segs))
                recs.append(r + rows + lens + [0, 0, 0] + [0, 0, 0])

        if kind == "sym":
            for s in range(nseq):
                base = s * N
                add(base, Lt, [(base, Lt)])
                add(base + Lt, Ls, [(base, Lt + Ls)])
            max_keys = Lt + Ls
        else:
            B = nseq // 2
            for m in range(2):
                for b in range(B):
                    base = (m * B + b) * N
                    add(base, Lt, [(base, Lt)])
                    add(base + Lt, Ls, [(b * N, Lt), ((B + b) * N, Lt), (base + Lt, Ls)])
            max_keys = 2 * Lt + Ls
        t = torch.from_numpy(ops.order_tiles(recs)).to(self.dev)
        self._tiles[key] = (t, max_keys)
        return self._tiles[key]

    # ------------------------------------------------------------------------------------------ backbone
    def _embed_buf(self, rows, lane=0):
        """A-operand buffer of the token-embedding GEMM: one row per token, in final token order (one per concurrent
        stream `lane`)."""
        return self._buf((rows, lane), "patches", (rows, 3 * 256), self.act)

    def _stage_tokens(self, bb, img, buf, tok_off, tok_per_seq):
        """Write the embedding-GEMM input rows of the crops `img` [n,3,S,S] at rows b*tok_per_seq + tok_off + patch.
        MixViT: the 16x16 patch matrix.  (ConvMAE: the conv stem's output, engine_online.ConvMAEOnlineEngine.)"""
        ops.patchify(img, buf, tok_off, tok_per_seq)

    def _embed(self, bb, B, imgs_t, imgs_ot, imgs_s, x, lane=0):
        """Embed the three crops of `B` sequences into x [B*N0, dim] (rows b*N0 + [t | ot | s]): staging of the GEMM
        input in token order, then ONE GEMM with bias and the positional table in the epilogue."""
        buf = self._embed_buf(x.shape[0], lane)
        n_t = self.gt * self.gt
        self._stage_tokens(bb, imgs_t, buf, 0, self.N0)
        self._stage_tokens(bb, imgs_ot, buf, n_t, self.N0)
        self._stage_tokens(bb, imgs_s, buf, 2 * n_t, self.N0)
        ops.gemm(buf, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, bb["pos"], out=x)

    def _block(self, blk, x, nseq, N, Ls, ln_period, cross, tag, ce_keep=None, gidx=None, tiles=None, qkv1=None,
               qkv_out=None, first=True, last=True):
        """One pre-LN block on x [nseq*N, dim] fp32 (in place).  Returns (x, N, Ls, gidx) - changed by CE.
        tiles / qkv1 / qkv_out: explicit attention tile table, second (cached) qkv buffer and the buffer that receives
        this block's qkv (the cached-template paths of engine_online.py).
        first / last: position of the block in its chain of `_block` calls on the same x (folded LayerNorm: the first
        block has no producing GEMM in front of it, the last one no consumer behind it)."""
        if self.ln_fold:
            return self._block_folded(blk, x, nseq, N, Ls, ln_period, cross, tag, ce_keep, gidx, tiles, qkv1, qkv_out,
                                      first, last)
        M = nseq * N
        dim = self.dim
        h = self._buf(tag, "h", (M, dim), self.act)
        qkv = qkv_out if qkv_out is not None else self._buf(tag, "qkv", (M, 3 * dim), self.act)
        att = self._buf(tag, "att", (M, dim), self.act)
        g0, b0, g1, b1 = blk["ln1"]
        self._ln(x, g0, b0, g1, b1, ln_period, 1e-6, h)
        ops.gemm(h, blk["qkv_w"], blk["qkv_b"], out=qkv)
        tiles, max_keys = tiles if tiles is not None else self._attn_tiles("cross" if cross else "sym", nseq, N, Ls)
        ops.mixattn(qkv, qkv1, dim, self.heads, tiles, max_keys, att, self.scale)
        ops.gemm(att, blk["proj_w"], blk["proj_b"], ops.ACT_NONE, x, None, out=x)
        if ce_keep is not None:
            x, N, Ls, gidx = self._candidate_elimination(qkv, x, nseq, N, Ls, ce_keep, gidx, tag,
                                                         q_src=qkv1 if N == Ls else None)
            M = nseq * N
            h = self._buf(tag, "h", (M, dim), self.act)
        g0, b0, g1, b1 = blk["ln2"]
        self._ln(x, g0, b0, g1, b1, nseq * N // 2 if ln_period else 0, 1e-6, h)
        hid = self._buf(tag, "hid", (M, 4 * dim), self.act)
        ops.gemm(h, blk["fc1_w"], blk["fc1_b"], ops.ACT_GELU, out=hid)
        ops.gemm(hid, blk["fc2_w"], blk["fc2_b"], ops.ACT_NONE, x, None, out=x)
        return x, N, Ls, gidx

    # ---- the same block with both LayerNorms folded into the GEMMs around them (bf16 mode; mmt_gemm_bf16_ex):
    #   proj / fc2 epilogue   : x += ... (fp32, as before) AND a bf16 copy xb of the new rows + their partial (sum, sumsq)
    #   qkv / fc1 GEMM        : A = xb (raw rows), W' = W diag(gamma); epilogue = rs * (acc - mu * colsum) + bias', GELU
    # so the normalised activations are never written to or read from HBM: 2 launches and ~270 MB of traffic per block
    # and modality at 64 sequences.  Rows without a producing GEMM (first block of a chain, rows re-gathered by candidate
    # elimination) get xb / sums from mmt_rowstats_cast.  Modality-specific norms = one consumer launch per row range.
    def _ln_bufs(self, tag, M):
        return (self._buf(tag, "xb", (M, self.dim), torch.bfloat16),
                self._buf(tag, "ln_sums", (self.dim // 128, M, 2), torch.float32))       # slot-major partial sums

    def _gemm_ln(self, xb, sums, sets, period, act, out):
        M = xb.shape[0]
        if period and len(sets) > 1:
            for (w, b, cs), (r0, r1) in zip(sets, ((0, period), (period, M))):
                ops.gemm(xb[r0:r1], w, b, act, out=out[r0:r1], ln_stats=sums[:, r0:r1], ln_eps=1e-6, colsum=cs)
        else:
            w, b, cs = sets[0]
            ops.gemm(xb, w, b, act, out=out, ln_stats=sums, ln_eps=1e-6, colsum=cs)

    def _block_folded(self, blk, x, nseq, N, Ls, ln_period, cross, tag, ce_keep, gidx, tiles, qkv1, qkv_out, first, last):
        M = nseq * N
        dim = self.dim
        qkv = qkv_out if qkv_out is not None else self._buf(tag, "qkv", (M, 3 * dim), self.act)
        att = self._buf(tag, "att", (M, dim), self.act)
        xb, sums = self._ln_bufs(tag, M)
        if first:
            ops.rowstats_cast(x, xb, sums)
        self._gemm_ln(xb, sums, blk["qkv_f"], ln_period, ops.ACT_NONE, qkv)
        tiles, max_keys = tiles if tiles is not None else self._attn_tiles("cross" if cross else "sym", nseq, N, Ls)
        ops.mixattn(qkv, qkv1, dim, self.heads, tiles, max_keys, att, self.scale)
        if ce_keep is not None:
            ops.gemm(att, blk["proj_w"], blk["proj_b"], ops.ACT_NONE, x, None, out=x)
            x, N, Ls, gidx = self._candidate_elimination(qkv, x, nseq, N, Ls, ce_keep, gidx, tag,
                                                         q_src=qkv1 if N == Ls else None)
            M = nseq * N
            xb, sums = self._ln_bufs(tag, M)
            ops.rowstats_cast(x, xb, sums)
        else:
            ops.gemm(att, blk["proj_w"], blk["proj_b"], ops.ACT_NONE, x, None, out=x, xb_out=xb, stats_out=sums)
        hid = self._buf(tag, "hid", (M, 4 * dim), self.act)
        self._gemm_ln(xb, sums, blk["fc1_f"], nseq * N // 2 if ln_period else 0, ops.ACT_GELU, hid)
        if last:
            ops.gemm(hid, blk["fc2_w"], blk["fc2_b"], ops.ACT_NONE, x, None, out=x)
        else:
            ops.gemm(hid, blk["fc2_w"], blk["fc2_b"], ops.ACT_NONE, x, None, out=x, xb_out=xb, stats_out=sums)
        return x, N, Ls, gidx

    def _ln(self, x, g0, b0, g1, b1, period, eps, out):
        if self.bf16:
            ops.layernorm(x, g0, b0, g1, b1, period, eps, out_bf16=out)
        else:
            ops.layernorm(x, g0, b0, g1, b1, period, eps, out_f32=out)

    def _candidate_elimination(self, qkv, x, nseq, N, Ls, keep_ratio, gidx, tag, q_src=None):
        """asymmetric_shared_ce.py:49-101 (test-time branch: no template mask).
        q_src: the block's cached template q/k/v [nseq*Lt, 3*dim] when x / qkv hold the search rows only (forward_search);
        the template rows of x (N - Ls of them per sequence: Lt or 0) are carried over unchanged."""
        keep = math.ceil(keep_ratio * Ls)
        if keep == Ls:
            return x, N, Ls, gidx
        B = nseq // 2
        lt_x = N - Ls                       # template rows inside x: Lt (full forward) or 0 (search-only rows)
        nqt = 2 * self.Lt // 32
        partial = self._buf(tag, "ce_partial", (B * self.heads * nqt * 2 * Ls,), torch.float32)
        # per-stage workspaces, allocated once per batch size (no allocation inside a step; the aux outputs below are
        # views of them, valid until the next forward)
        stage = len(self.aux["ce_scores"])
        scores = self._buf(tag, f"ce_scores{stage}", (B, 2 * Ls), torch.float32)
        if q_src is None:
            ops.ce_scores(qkv, self.dim, self.heads, B, N, self.Lt, Ls, self.scale, partial, scores)
        else:
            ops.ce_scores(qkv, self.dim, self.heads, B, N, self.Lt, Ls, self.scale, partial, scores, q_src=q_src,
                          q_seq_rows=self.Lt, k_row_off=lt_x)
        g_keep = self._buf(tag, f"ce_keep{stage}", (nseq, keep), torch.float32)
        g_rem = self._buf(tag, f"ce_rem{stage}", (nseq, Ls - keep), torch.float32)
        order = self._buf(tag, f"ce_order{stage}", (nseq, Ls), torch.int32)
        ops.ce_topk(scores, B, Ls, keep, gidx, g_keep, g_rem, order)
        n_new = lt_x + keep
        x_new = self._buf(tag, f"ce_x{stage}", (nseq * n_new, self.dim), torch.float32)
        ops.ce_gather_tokens(x, nseq, N, lt_x, order, Ls, keep, x_new)
        self.aux["ce_scores"].append(scores)
        self.aux["ce_keep"].append(g_keep)
        self.aux["ce_removed"].append(g_rem)
        return x_new, n_new, keep, g_keep

    def _run_backbone(self, bb, x, nseq, tag, stacked):
        """All blocks on x [nseq*N0, dim]; returns the search-token rows [nseq*Ls0, dim] in `act` dtype."""
        st = self._backbone_begin(x, nseq, stacked)
        for i in range(len(bb["blocks"])):
            self._backbone_block(bb, st, i, nseq, tag)
        return self._backbone_end(st, nseq, tag)

    def _backbone_begin(self, x, nseq, stacked):
        N, Ls = self.N0, self.Ls0
        st = dict(x=x, N=N, Ls=Ls, gidx=None, ce_i=0, cross=stacked and self.variant in CROSS_MODAL,
                  per_ln=stacked and self.variant in PER_MODALITY_LN)
        if self.ce_loc:
            key = ("gidx0", nseq, Ls)
            gidx = self._ws.get(key)        # global search-token indices 0..Ls-1 per sequence (float32 like the
            if gidx is None:                # reference, asymmetric_shared_ce.py:397-399); built once per batch size
                gidx = torch.arange(Ls, device=self.dev, dtype=torch.float32).repeat(nseq, 1).contiguous()
                self._ws[key] = gidx
            st["gidx"] = gidx
            self.aux.update(ce_scores=[], ce_keep=[], ce_removed=[])
        return st

    def _backbone_block(self, bb, st, i, nseq, tag):
        ce_keep = None
        if i in self.ce_loc:
            ce_keep = self.ce_keep[st["ce_i"]]
            st["ce_i"] += 1
            if not ce_keep < 1:
                ce_keep = None
        st["x"], st["N"], st["Ls"], st["gidx"] = self._block(
            bb["blocks"][i], st["x"], nseq, st["N"], st["Ls"], (nseq * st["N"] // 2) if st["per_ln"] else 0, st["cross"],
            tag, ce_keep, st["gidx"], first=(i == 0), last=(i == len(bb["blocks"]) - 1))

    def _backbone_end(self, st, nseq, tag):
        x, N, Ls, gidx = st["x"], st["N"], st["Ls"], st["gidx"]
        self._last_x = (x, N)            # residual stream after the last block (template rows: engine_online.py)
        feat = self._buf(tag, "search_rows", (nseq * self.Ls0, self.dim), self.act)
        if Ls != self.Ls0:
            ops.ce_recover(x, nseq, N, self.Lt, gidx, Ls, self.Ls0, feat)
        else:
            ops.copy_rows(x, N, self.Lt, Ls, nseq, feat)
        return feat

    # Two independent backbones (mixformer_vit_rgbt: backbone_v / backbone_i never exchange data before the fusion,
    # mixformer.py:379-380) run on TWO CUDA streams, launches interleaved block by block.  Every kernel of the chain
    # leaves SMs idle in its last partial wave (the persistent GEMMs: 339 tiles on 74 CTA pairs = 4.58 waves) and the
    # LayerNorm / attention kernels leave the tensor pipe idle altogether; with a second, independent chain queued
    # those SMs pick up the other modality's CTAs at once.  Same kernels on the same data: results are bit-identical
    # to the sequential order.  MMT_TWO_STREAMS=0 restores it (A/B measurements).
    def _lanes(self):
        if getattr(self, "_side", None) is None:
            import os
            self._side = torch.cuda.Stream(device=self.dev)
            self._fork_ev, self._join_ev = torch.cuda.Event(), torch.cuda.Event()
            self._two_streams = os.environ.get("MMT_TWO_STREAMS", "1") != "0"
            self._head_lanes = os.environ.get("MMT_HEAD_LANES", "1") != "0"      # A/B switch of _run_head's two lanes
            self._half_sm_small = os.environ.get("MMT_HALF_SM_SMALL", "1") != "0"  # A/B switch of the small-GEMM SM budget
            self._n_sm = torch.cuda.get_device_properties(self.dev).multi_processor_count
        return self._side

    def set_two_streams(self, on: bool):
        """Developer / bench switch: run the two modality backbones of mixformer_vit_rgbt sequentially on one stream
        (per-launch CUDA-event brackets are only meaningful without a concurrent second chain)."""
        self._lanes()
        self._two_streams = bool(on)
        return self

    def _run_two_backbones(self, B, t, ot, s, x, wait):
        M1 = B * self.N0
        side = self._lanes()
        main = torch.cuda.current_stream()
        if not self._two_streams:
            feats = []
            for m in range(2):
                xm = x[m * M1:(m + 1) * M1]
                wait(m)
                self._embed(self.bbs[m], B, t[m], ot[m], s[m], xm)
                feats.append(self._run_backbone(self.bbs[m], xm, B, ("bb", B, m), False))
            return feats
        self._fork_ev.record(main)
        side.wait_event(self._fork_ev)
        lanes = (main, side)
        sts = []
        # few rows per modality (single-sequence latency): every GEMM is one partial wave of CTAs that each own an SM, so two
        # whole-GPU grids would alternate instead of overlapping - give each chain half the SMs (ops.config_small_gemm_sms)
        small = self.bf16 and self._half_sm_small and M1 <= 8 * 128
        prev_budget = ops.config_small_gemm_sms(self._n_sm // 2) if small else None
        try:
            for m in range(2):
                with torch.cuda.stream(lanes[m]):
                    xm = x[m * M1:(m + 1) * M1]
                    wait(m)
                    self._embed(self.bbs[m], B, t[m], ot[m], s[m], xm, lane=m)
                    sts.append(self._backbone_begin(xm, B, False))
            for i in range(self.depth):
                for m in range(2):
                    with torch.cuda.stream(lanes[m]):
                        self._backbone_block(self.bbs[m], sts[m], i, B, ("bb", B, m))
        finally:
            if small:
                ops.config_small_gemm_sms(prev_budget)
        feats = []
        for m in range(2):
            with torch.cuda.stream(lanes[m]):
                feats.append(self._backbone_end(sts[m], B, ("bb", B, m)))
        self._join_ev.record(side)
        main.wait_event(self._join_ev)
        return feats

    # ------------------------------------------------------------------------------------------ fusion
    def _conv1x1_gn(self, a, p, B, HW, out, tag, out_seq_rows=0, out_row_off=0):
        """Conv2d(k=1) + GroupNorm(32) (fusion_utils.py:252-268); `out` dtype selects the fp32 / bf16 output."""
        C = p["w"].shape[0]
        pre = self._buf(tag, "gn_pre", (B * HW, C), torch.float32)
        ops.gemm(a, p["w"], p["b"], out=pre)
        if out.dtype == torch.float32:
            ops.groupnorm(pre, B, HW, 32, p["gn_w"], p["gn_b"], 1e-5, out_f32=out, out_seq_rows=out_seq_rows,
                          out_row_off=out_row_off)
        else:
            ops.groupnorm(pre, B, HW, 32, p["gn_w"], p["gn_b"], 1e-5, out_bf16=out, out_seq_rows=out_seq_rows,
                          out_row_off=out_row_off)

    def _run_fusion(self, sv, si, B):
        """fusion_vi (fusion_utils.py:270-279 and variants) on search-token rows sv, si [B*HW, 768] -> [B*HW, 768]."""
        F_ = self.fusion
        if "cat_convs" in F_:                 # RGBT_Fusion_Cat
            tag = ("fuscat", B)
            a = ops.concat_cols(sv, si, self._buf(tag, "cat", (B * self.Ls0, 2 * self.dim), self.act))
            for j, (w, b) in enumerate(F_["cat_convs"]):
                o = self._buf(tag, f"o{j}", (B * self.Ls0, w.shape[0]), self.act)
                self._conv3x3(a, B, self.gs, self.gs, a.shape[1], w, b, o, tag)
                a = o
            return a
        HW, d, L = self.Ls0, self.d_model, self.Ls0
        tag = ("fus", B)
        src = self._buf(tag, "src", (B, 2 * HW, d), torch.float32)
        # per-modality 1x1 conv + GroupNorm written into the two halves of each sequence's token block
        for m, a in enumerate((sv, si)):
            self._conv1x1_gn(a, F_["in"][m], B, HW, src, tag, out_seq_rows=2 * HW, out_row_off=m * HW)
        src2 = src.view(B * 2 * HW, d)
        val_in = self._buf(tag, "val_in", (B * 2 * HW, d), self.act)
        q_in = self._buf(tag, "q_in", (B * HW, 2 * d), self.act)
        value = self._buf(tag, "value", (B * 2 * HW, d), self.act)
        offw = self._buf(tag, "offw", (B * HW, F_["layers"][0]["offw_w"].shape[0]), torch.float32)
        samp = self._buf(tag, "samp", (B * 2 * HW, d), self.act)
        hid = self._buf(tag, "ffn", (B * 2 * HW, self.d_ffn), self.act)
        src_act = self._buf(tag, "src_act", (B * 2 * HW, d), self.act)
        for Lw in F_["layers"]:
            ops.fusion_prep(src2, F_["pos"], B, HW, val_in, q_in)
            ops.gemm(val_in, Lw["val_w"], Lw["val_b"], out=value)
            ops.gemm(q_in, Lw["offw_w"], Lw["offw_b"], out=offw)
            ops.msda_bimodal(value, offw, samp, B, self.gs, self.gs, self.n_heads_f, d // self.n_heads_f, self.n_points_f)
            ops.gemm(samp, Lw["out_w"], Lw["out_b"], ops.ACT_NONE, src2, None, out=src2)
            g0, b0, g1, b1 = Lw["ln1"]
            if self.bf16:
                ops.layernorm(src2, g0, b0, g1, b1, HW, 1e-5, out_f32=src2, out_bf16=src_act)
                a_in = src_act
            else:
                ops.layernorm(src2, g0, b0, g1, b1, HW, 1e-5, out_f32=src2)
                a_in = src2
            ops.gemm(a_in, Lw["l1_w"], Lw["l1_b"], ops.ACT_RELU, out=hid)
            ops.gemm(hid, Lw["l2_w"], Lw["l2_b"], ops.ACT_NONE, src2, None, out=src2)
            g0, b0, g1, b1 = Lw["ln2"]
            ops.layernorm(src2, g0, b0, g1, b1, HW, 1e-5, out_f32=src2)
        cat_in = self._buf(tag, "cat_in", (B * HW, 2 * d), self.act)
        ops.fusion_prep(src2, None, B, HW, None, cat_in)
        fused = self._buf(tag, "fused", (B * HW, 768), self.act)
        self._conv1x1_gn(cat_in, F_["out"], B, HW, fused, (tag, "o"))
        return fused

    # ------------------------------------------------------------------------------------------ head
    def _conv3x3(self, src, B, H, W, C, w, b, out, tag, up=None):
        """Conv2d(k=3, pad=1) + folded BN + ReLU on NHWC rows.  `up` = (src1, s1, src2, s2): the conv input is
        up_s1(src1) + up_s2(src2) (nearest), head.py:166-178.  bf16: implicit GEMM over 4-D TMA boxes (the upsampled
        sum is materialised once); fp32 parity mode: im2col (upsampling folded into the gather) + SIMT GEMM."""
        if self.bf16:
            if up is not None:
                src1, s1, src2, s2 = up
                src = ops.upsample_add(src1, s1, B, H, W, C, self._buf(tag, f"up{H}", (B * H * W, C), self.act), src2, s2)
            return ops.conv3x3(src, B, H, W, C, w, b, ops.ACT_RELU, out)
        need = B * H * W * 9 * C
        col = self._ws.get("im2col")
        if col is None or col.numel() < need:       # one growing im2col buffer (fp32 parity mode only)
            col = torch.empty((need,), device=self.dev, dtype=self.act)
            self._ws["im2col"] = col
        col = col[:need].view(B * H * W, 9 * C)
        if up is not None:
            src1, s1, src2, s2 = up
            ops.im2col3x3(src1, s1, B, H, W, C, col, src2, s2)
        else:
            ops.im2col3x3(src, 1, B, H, W, C, col)
        return ops.gemm(col, w, b, ops.ACT_RELU, out=out)

    def _run_head(self, feat, B, want_maps=True):
        """Corner head on NHWC rows feat [B*gs*gs, C] -> boxes cxcywh [B,4] (+ raw score maps)."""
        if self.head_type == "CORNER":
            return self._run_head_plain(feat, B, want_maps)
        H = self.head
        gs = self.gs
        ch = self.head_ch                       # 384
        C = feat.shape[1]
        tag = ("head", B)
        n18, n36, n72 = B * gs * gs, B * 4 * gs * gs, B * 16 * gs * gs
        s1 = self._buf(tag, "s1", (n18, self.s1_width), self.act)
        self._conv3x3(feat, B, gs, gs, C, H["s1_w"], H["s1_b"], s1, tag)
        sl = lambda n: s1[:, self.s1_cols[n][0]: self.s1_cols[n][1]]
        x4s, a3s, a4s = [], [], []
        # the two corner branches share only s1: the bottom-right chain runs on the side stream (small-N convolutions
        # with one-wave grids - each alone leaves most SMs idle)
        side = self._lanes()
        main = torch.cuda.current_stream()
        lanes = (main, side) if (self.bf16 and self._two_streams and self._head_lanes) else (main, main)
        if lanes[1] is not main:
            self._fork_ev.record(main)
            side.wait_event(self._fork_ev)
        for ci, c in enumerate(("tl", "br")):
            with torch.cuda.stream(lanes[ci]):
                tag = ("head", B, c)
                x2 = self._buf(tag, "x2" + c, (n18, ch // 2), self.act)
                self._conv3x3(sl(f"conv1_{c}"), B, gs, gs, ch, H[f"conv2_{c}_w"], H[f"conv2_{c}_b"], x2, tag)
                # up-1: conv3(up2(adjust1(x)) + up2(x2)) at 2gs x 2gs
                x3 = self._buf(tag, "x3" + c, (n36, ch // 4), self.act)
                self._conv3x3(None, B, 2 * gs, 2 * gs, ch // 2, H[f"conv3_{c}_w"], H[f"conv3_{c}_b"], x3, tag,
                              up=(sl(f"adjust1_{c}"), 2, x2, 2))
                # up-2: conv4(up4(adjust2(x)) + up2(x3)) at 4gs x 4gs
                x4 = self._buf(tag, "x4" + c, (n72, ch // 8), self.act)
                self._conv3x3(None, B, 4 * gs, 4 * gs, ch // 4, H[f"conv4_{c}_w"], H[f"conv4_{c}_b"], x4, tag,
                              up=(sl(f"adjust2_{c}"), 4, x3, 2))
                # side branches: adjust3 on x2 (gs), adjust4 on x3 (2gs)
                a = x2
                for j, co in enumerate((ch // 4, ch // 8, 1)):
                    o = self._buf(tag, f"a3{c}{j}", (n18, co), self.act)
                    self._conv3x3(a, B, gs, gs, a.shape[1], H[f"adjust3_{c}.{j}_w"], H[f"adjust3_{c}.{j}_b"], o, tag)
                    a = o
                a3s.append(a)
                a = x3
                for j, co in enumerate((ch // 8, 1)):
                    o = self._buf(tag, f"a4{c}{j}", (n36, co), self.act)
                    self._conv3x3(a, B, 2 * gs, 2 * gs, a.shape[1], H[f"adjust4_{c}.{j}_w"], H[f"adjust4_{c}.{j}_b"], o,
                                  tag)
                    a = o
                a4s.append(a)
                x4s.append(x4)
        if lanes[1] is not main:
            self._join_ev.record(side)
            main.wait_event(self._join_ev)
        S = 4 * gs
        maps = torch.empty((B, 2, S * S), device=self.dev, dtype=torch.float32) if want_maps else None
        xyxy = torch.empty((B, 4), device=self.dev, dtype=torch.float32)
        boxes = torch.empty((B, 4), device=self.dev, dtype=torch.float32)
        ops.corner_decode(x4s, (H["w5_tl"], H["w5_br"]), (H["b5_tl"], H["b5_br"]), a3s, a4s, B, S, 4.0,
                          float(self.search_size), xyxy, boxes, maps)
        self._last_xyxy = xyxy
        return boxes, maps

    def _run_head_plain(self, feat, B, want_maps=True):
        """Corner_Predictor (head.py:23-94): both corners' conv1 as one launch, then conv2..conv4 per corner at the
        backbone resolution, conv5 (1x1) + softmax + soft-argmax (stride 16) in the decode kernel."""
        H, gs, ch, C = self.head, self.gs, self.head_ch, feat.shape[1]
        tag = ("headp", B)
        n = B * gs * gs
        s1 = self._buf(tag, "s1", (n, self.s1_width), self.act)
        self._conv3x3(feat, B, gs, gs, C, H["s1_w"], H["s1_b"], s1, tag)
        x4s = []
        for c in ("tl", "br"):
            a = s1[:, self.s1_cols[f"conv1_{c}"][0]: self.s1_cols[f"conv1_{c}"][1]]
            for j, co in ((2, ch // 2), (3, ch // 4), (4, ch // 8)):
                o = self._buf(tag, f"x{j}{c}", (n, co), self.act)
                self._conv3x3(a, B, gs, gs, a.shape[1], H[f"conv{j}_{c}_w"], H[f"conv{j}_{c}_b"], o, (tag, c))
                a = o
            x4s.append(a)
        maps = torch.empty((B, 2, gs * gs), device=self.dev, dtype=torch.float32) if want_maps else None
        xyxy = torch.empty((B, 4), device=self.dev, dtype=torch.float32)
        boxes = torch.empty((B, 4), device=self.dev, dtype=torch.float32)
        ops.corner_decode(x4s, (H["w5_tl"], H["w5_br"]), (H["b5_tl"], H["b5_br"]), None, None, B, gs, 16.0,
                          float(self.search_size), xyxy, boxes, maps)
        self._last_xyxy = xyxy
        return boxes, maps

    # ------------------------------------------------------------------------------------------ forward
    def _check_img(self, t, size):
        if not t.is_cuda:
            raise NotImplementedError("mmt_b200 forward is CUDA-only (no CPU fallback)")
        if t.dim() != 4 or t.shape[1] != 3 or t.shape[2] != size or t.shape[3] != size:
            raise RuntimeError(f"expected a [B,3,{size},{size}] crop, got {tuple(t.shape)}")
        return t.float().contiguous() if (t.dtype != torch.float32 or not t.is_contiguous()) else t

    def forward(self, template, online_template, search, want_maps=True, ready_events=None):
        """ready_events: optional [event_v, event_i] recorded on a copy stream after that modality's crops arrived
        (runner.FrameStep): the compute stream waits for a modality only right before its patch embedding, so the
        second modality's host->device copy overlaps the first one's work."""
        self.aux = {}
        wait = (lambda m: torch.cuda.current_stream().wait_event(ready_events[m])) if ready_events else (lambda m: None)
        if not self.rgbt:
            t, ot, s = (self._check_img(template, self.template_size), self._check_img(online_template, self.template_size),
                        self._check_img(search, self.search_size))
            B = s.shape[0]
            x = self._buf(B, "x", (B * self.N0, self.dim), torch.float32)
            wait(0)
            self._embed(self.bbs[0], B, t, ot, s, x)
            feat = self._run_backbone(self.bbs[0], x, B, ("bb", B), False)
            boxes, maps = self._run_head(feat, B, want_maps)
            return dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps, feat_rows=feat)
        t = [self._check_img(v, self.template_size) for v in template]
        ot = [self._check_img(v, self.template_size) for v in online_template]
        s = [self._check_img(v, self.search_size) for v in search]
        B = s[0].shape[0]
        M1 = B * self.N0
        x = self._buf(B, "x", (2 * M1, self.dim), torch.float32)
        if self.variant == "mixformer_vit_rgbt":      # two independent backbones (mixformer.py:379-380)
            sv, si = self._run_two_backbones(B, t, ot, s, x, wait)
        else:                                          # batch-stacked modalities, shared weights
            for m in range(2):
                wait(m)
                self._embed(self.bbs[0], B, t[m], ot[m], s[m], x[m * M1:(m + 1) * M1])
            f = self._run_backbone(self.bbs[0], x, 2 * B, ("bb", B), True)
            sv, si = f[: B * self.Ls0], f[B * self.Ls0:]
        fused = self._run_fusion(sv, si, B)
        boxes, maps = self._run_head(fused, B, want_maps)
        res = dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps, feat_rows=fused, search_rows=(sv, si), **self.aux)
        return res

    # ------------------------------------------------------------------------------------------ template-side reuse
    # in every variant template rows attend template keys only, so the templates'
    # per-layer q/k/v do not depend on the search crop and are identical from one template update to the next.
    # cache_templates() runs the template tokens once and keeps every layer's packed qkv; forward_search() then runs
    # the search tokens only, reading the cached rows as a second key/value source (tile records with k_buf = 1).
    # Symmetric variants (template/search interaction only through the search rows' keys): results are bit-identical
    # to forward() - same GEMM K-order per row, same key-block boundaries in the attention kernels.
    def _seq_tiles(self, kind, nseq, rows_per_seq, key_segs_fn):
        key = (kind, nseq, rows_per_seq)
        hit = self._tiles.get(key)
        if hit is None:
            recs, max_keys = [], 0
            for sq in range(nseq):
                segs = key_segs_fn(sq)
                max_keys = max(max_keys, sum(sg[2] for sg in segs))
                pad = segs + [(0, 0, 0)] * (3 - len(segs))
                for o in range(0, rows_per_seq, 128):
                    q0 = sq * rows_per_seq + o
                    recs.append([q0, min(128, rows_per_seq - o), q0, len(segs)] + [sg[1] for sg in pad] +
                                [sg[2] for sg in pad] + [sg[0] for sg in pad] + [0, 0, 0])
            hit = (torch.from_numpy(ops.order_tiles(recs)).to(self.dev), max_keys)
            self._tiles[key] = hit
        return hit

    def _check_cacheable(self):
        """Every variant qualifies: template rows attend template keys of their own modality only - also in the cross-modal
        blocks (asymmetric_shared.py:55-104: q_mt x k_mt per modality) and with candidate elimination, which prunes search
        tokens only (asymmetric_shared_ce.py:49-101)."""
        return None

    def cache_templates(self, template, online_template):
        self._check_cacheable()
        Lt, T = self.Lt, self.gt * self.gt
        if not self.rgbt:
            groups = [(self.bbs[0], [self._check_img(template, self.template_size)],
                       [self._check_img(online_template, self.template_size)], False)]
        else:
            t = [self._check_img(v, self.template_size) for v in template]
            ot = [self._check_img(v, self.template_size) for v in online_template]
            if self.variant == "mixformer_vit_rgbt":
                groups = [(self.bbs[m], [t[m]], [ot[m]], False) for m in range(2)]
            else:
                groups = [(self.bbs[0], t, ot, self.variant in PER_MODALITY_LN)]
        self._tcache = []
        for gi, (bb, ts, ots, per_ln) in enumerate(groups):
            B = ts[0].shape[0]
            nseq = B * len(ts)
            tag = ("tcache", B, gi)
            x = self._buf(tag, "x", (nseq * Lt, self.dim), torch.float32)
            buf = self._embed_buf(nseq * Lt)
            for m in range(len(ts)):
                sub = buf[m * B * Lt:(m + 1) * B * Lt]
                self._stage_tokens(bb, ts[m], sub, 0, Lt)
                self._stage_tokens(bb, ots[m], sub, T, Lt)
            if "pos_tt" not in bb:       # per backbone: backbone_v / backbone_i carry their own positional tables
                bb["pos_tt"] = bb["pos"][:Lt].contiguous()
                bb["pos_ss"] = bb["pos"][Lt:].contiguous()
            ops.gemm(buf, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, bb["pos_tt"], out=x)
            tiles = self._seq_tiles("tcache", nseq, Lt, lambda sq: [(0, sq * Lt, Lt)])
            cache = [self._buf(tag, f"qkv{i}", (nseq * Lt, 3 * self.dim), self.act) for i in range(self.depth)]
            for i, blk in enumerate(bb["blocks"]):
                self._block(blk, x, nseq, Lt, 0, (nseq * Lt // 2) if per_ln else 0, False, tag, tiles=tiles,
                            qkv_out=cache[i], first=(i == 0), last=(i == self.depth - 1))
            self._tcache.append((cache, nseq, B))

    def forward_search(self, search, want_maps=True):
        """The per-frame forward of forward() for the search crops alone, against cache_templates()' cache."""
        self._check_cacheable()
        if getattr(self, "_tcache", None) is None:
            raise RuntimeError("forward_search called before cache_templates")
        self.aux = {}
        Lt, Ls = self.Lt, self.Ls0
        ss = [self._check_img(search, self.search_size)] if not self.rgbt else \
            [self._check_img(v, self.search_size) for v in search]
        B = ss[0].shape[0]
        if self.variant == "mixformer_vit_rgbt":
            groups = [(self.bbs[m], [ss[m]], False) for m in range(2)]
        else:
            groups = [(self.bbs[0], ss, self.variant in PER_MODALITY_LN)]
        feats = []
        for gi, (bb, imgs, per_ln) in enumerate(groups):
            cache, nseq_c, Bc = self._tcache[gi]
            nseq = B * len(imgs)
            if nseq != nseq_c:
                raise RuntimeError(f"cached templates are for {Bc} sequences, got {B} search crops")
            tag = ("scache", B, gi)
            x = self._buf(tag, "x", (nseq * Ls, self.dim), torch.float32)
            buf = self._embed_buf(nseq * Ls)
            for m in range(len(imgs)):
                self._stage_tokens(bb, imgs[m], buf[m * B * Ls:(m + 1) * B * Ls], 0, Ls)
            ops.gemm(buf, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, bb["pos_ss"], out=x)
            cross = self.variant in CROSS_MODAL
            Bh = nseq // 2

            def tiles_for(ls):
                if not cross:      # keys: the sequence's cached template rows, then its own search rows
                    return self._seq_tiles(("scache", ls), nseq, ls, lambda sq: [(1, sq * Lt, Lt), (0, sq * ls, ls)])
                # cross-modal blocks: cached template rows of BOTH modalities (RGB first, as in _attn_tiles("cross")),
                # then the own search rows
                return self._seq_tiles(("scache_x", ls), nseq, ls,
                                       lambda sq: [(1, (sq % Bh) * Lt, Lt), (1, (Bh + sq % Bh) * Lt, Lt), (0, sq * ls, ls)])

            ls, gidx, ce_i = Ls, None, 0
            if self.ce_loc:
                gidx = self._ws.get(("gidx0", nseq, Ls))
                if gidx is None:
                    gidx = torch.arange(Ls, device=self.dev, dtype=torch.float32).repeat(nseq, 1).contiguous()
                    self._ws[("gidx0", nseq, Ls)] = gidx
                self.aux.update(ce_scores=[], ce_keep=[], ce_removed=[])
            for i, blk in enumerate(bb["blocks"]):
                ce_keep = None
                if i in self.ce_loc:
                    ce_keep = self.ce_keep[ce_i] if self.ce_keep[ce_i] < 1 else None
                    ce_i += 1
                x, _, ls, gidx = self._block(blk, x, nseq, ls, ls, (nseq * ls // 2) if per_ln else 0, cross, tag, ce_keep,
                                             gidx, tiles=tiles_for(ls), qkv1=cache[i], first=(i == 0),
                                             last=(i == self.depth - 1))
            rows = self._buf(tag, "search_rows", (nseq * Ls, self.dim), self.act)
            if ls != Ls:
                ops.ce_recover(x, nseq, ls, 0, gidx, ls, Ls, rows)
            else:
                ops.copy_rows(x, Ls, 0, Ls, nseq, rows)
            feats.append(rows)
        if not self.rgbt:
            boxes, maps = self._run_head(feats[0], B, want_maps)
            return dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps, feat_rows=feats[0])
        if len(feats) == 2:
            sv, si = feats
        else:
            sv, si = feats[0][: B * Ls], feats[0][B * Ls:]
        fused = self._run_fusion(sv, si, B)
        boxes, maps = self._run_head(fused, B, want_maps)
        return dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps, feat_rows=fused, search_rows=(sv, si), **self.aux)

    def forward_head_only(self, search_feat):
        """Corner head on an NCHW feature map (forward_box_head, mixformer.py:325-338)."""
        B, C, H, W = search_feat.shape
        rows = search_feat.permute(0, 2, 3, 1).reshape(B * H * W, C).to(self.act).contiguous()   # layout plumbing
        boxes, maps = self._run_head(rows, B, True)
        return dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps)

    def rows_to_map(self, rows, B):
        """Token rows [B*gs*gs, C] -> the reference's [B, C, gs, gs] fp32 feature map (output formatting for
        return_features=True; not on the box path)."""
        return rows.view(B, self.gs, self.gs, -1).permute(0, 3, 1, 2).float().contiguous()
