"""Engine of the online (SPM) MixViT tracker: the full forward with the score head, and the cached-template test
path `set_online` / `forward_test` (lib/models/mixformer_vit/mixformer_online.py:80-113, 229-262, 286-360;
lib/models/mixformer_cvt/score_decoder.py:32-66).  Same kernels as engine.py; the cache is one packed qkv buffer
per layer that the attention kernels read as a second key/value source (tile records with k_buf = 1)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .engine import ForwardEngine, _f32


class OnlineEngine(ForwardEngine):
    def _pack(self, sd):
        super()._pack(sd)
        dev = self.dev
        g = lambda k: sd["score_branch." + k]
        S = {"token": _f32(g("score_token").reshape(1, -1), dev),
             "norm1": (_f32(g("norm1.weight"), dev), _f32(g("norm1.bias"), dev)), "layers": [], "mlp": []}
        for i in range(2):
            S["layers"].append(dict(
                q_w=self._w(g(f"proj_q.{i}.weight")), q_b=_f32(g(f"proj_q.{i}.bias"), dev),
                kv_w=self._w(torch.cat([g(f"proj_k.{i}.weight"), g(f"proj_v.{i}.weight")], 0)),
                kv_b=_f32(torch.cat([g(f"proj_k.{i}.bias"), g(f"proj_v.{i}.bias")], 0), dev),
                o_w=self._w(g(f"proj.{i}.weight")), o_b=_f32(g(f"proj.{i}.bias"), dev),
                ln=(_f32(g(f"norm2.{i}.weight"), dev), _f32(g(f"norm2.{i}.bias"), dev))))
        i = 0
        while f"score_branch.score_head.layers.{i}.weight" in sd:
            S["mlp"].append((self._w(g(f"score_head.layers.{i}.weight")), _f32(g(f"score_head.layers.{i}.bias"), dev)))
            i += 1
        self.spm = S
        self.spm_heads = self.dim // 64
        self.spm_scale = self.dim ** -0.5            # hidden_dim ** -0.5 (score_decoder.py:18), not head_dim
        self.qkv_mem = None                          # per-layer cached template qkv [Tm, 3C]
        self.templ_rows = None                       # cached feature rows of the first template [T, C]

    # ------------------------------------------------------------------------------------------ SPM
    def _spm_tiles(self, B, T):
        key = ("spm", B, T)
        hit = self._tiles.get(key)
        if hit is None:
            recs = [[b, 1, b, 1, b * T, 0, 0, T, 0, 0, 1, 0, 0, 0, 0, 0] for b in range(B)]
            hit = (torch.from_numpy(np.asarray(recs, dtype=np.int32)).to(self.dev), T)
            self._tiles[key] = hit
        return hit

    def _run_spm(self, feat32, templ_rows, xyxy, B):
        """ScoreDecoder.forward: feat32 fp32 NHWC rows [B*gs*gs, C]; templ_rows `act` [B*T, C]; xyxy fp32 [B,4]
        normalised -> logits fp32 [B]."""
        C, gs, S = self.dim, self.gs, self.spm
        tag = ("spm", B)
        rois = ops.spm_rois(xyxy, float(gs), self._buf(tag, "rois", (B, 5), torch.float32))
        pooled = self._buf(tag, "pooled", (B, 16, C), torch.float32)
        ops.prroi_pool(feat32.view(B, gs, gs, C), rois, 4, 4, 1.0, channels_last=True, out=pooled)
        pooled_a = ops.copy_rows(pooled.view(B * 16, C), B * 16, 0, B * 16, 1, self._buf(tag, "pooled_a", (B * 16, C), self.act))
        tok = self._ws.get(("spm_tok", B))
        if tok is None:
            tok = S["token"].expand(B, -1).contiguous()
            self._ws[("spm_tok", B)] = tok
        x = self._buf(tag, "x", (B, C), self.act)
        self._ln(tok, S["norm1"][0], S["norm1"][1], None, None, 0, 1e-5, x)
        qkv0 = self._buf(tag, "qkv0", (B, 3 * C), self.act)
        att = self._buf(tag, "att", (B, C), self.act)
        y = self._buf(tag, "y", (B, C), torch.float32)
        for i, (mem, T) in enumerate(((pooled_a, 16), (templ_rows, templ_rows.shape[0] // B))):
            L = S["layers"][i]
            qkv1 = self._buf(tag, f"qkv1_{i}", (B * T, 3 * C), self.act)
            ops.gemm(x, L["q_w"], L["q_b"], out=qkv0[:, :C])
            ops.gemm(mem, L["kv_w"], L["kv_b"], out=qkv1[:, C:])
            tiles, mk = self._spm_tiles(B, T)
            ops.mixattn(qkv0, qkv1, C, self.spm_heads, tiles, mk, att, self.spm_scale)
            ops.gemm(att, L["o_w"], L["o_b"], out=y)
            self._ln(y, L["ln"][0], L["ln"][1], None, None, 0, 1e-5, x)
        h = x
        n = len(S["mlp"])
        for i, (w, b) in enumerate(S["mlp"]):
            last = i == n - 1
            if last:
                o = torch.empty((B, w.shape[0]), device=self.dev, dtype=torch.float32)
            else:
                o = self._buf(tag, f"mlp{i}", (B, w.shape[0]), self.act)
            ops.gemm(h, w, b, ops.ACT_NONE if last else ops.ACT_RELU, out=o)
            h = o
        return h.view(B)

    def _scores(self, x, N, row_off_s, B, templ_rows, gt_bboxes):
        """fp32 search rows out of the residual stream -> SPM."""
        feat32 = ops.copy_rows(x, N, row_off_s, self.Ls0, B, self._buf(("spm", B), "feat32", (B * self.Ls0, self.dim), torch.float32))
        if gt_bboxes is not None:
            xyxy = gt_bboxes.detach().to(device=self.dev, dtype=torch.float32).reshape(B, 4).contiguous()
        else:
            xyxy = self._last_xyxy
        return self._run_spm(feat32, templ_rows, xyxy, B)

    # ------------------------------------------------------------------------------------------ full forward
    def forward(self, template, online_template, search, want_maps=True, run_score_head=True, gt_bboxes=None):
        self.aux = {}
        t, ot, s = (self._check_img(template, self.template_size), self._check_img(online_template, self.template_size),
                    self._check_img(search, self.search_size))
        B = s.shape[0]
        x = self._buf(B, "x", (B * self.N0, self.dim), torch.float32)
        self._embed(self.bbs[0], B, t, ot, s, x)
        feat = self._run_backbone(self.bbs[0], x, B, ("bb", B), False)
        boxes, maps = self._run_head(feat, B, want_maps)
        res = dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps, feat_rows=feat)
        if run_score_head:
            T = self.gt * self.gt
            templ = ops.copy_rows(x, self.N0, 0, T, B, self._buf(("spm", B), "templ", (B * T, self.dim), self.act))
            res["pred_scores"] = self._scores(x, self.N0, self.Lt, B, templ, gt_bboxes)
        return res

    # ------------------------------------------------------------------------------------------ cached-template path
    def _full_tiles(self, rows, key_segs, tag):
        """Query rows [0, rows) in 128-row tiles, every tile reading the same key segments [(buf, row0, len), ...]."""
        key = (tag, rows, tuple(key_segs))
        hit = self._tiles.get(key)
        if hit is None:
            recs = []
            for o in range(0, rows, 128):
                segs = list(key_segs) + [(0, 0, 0)] * (3 - len(key_segs))
                recs.append([o, min(128, rows - o), o, len(key_segs)] + [sg[1] for sg in segs] + [sg[2] for sg in segs] +
                            [sg[0] for sg in segs] + [0, 0, 0])
            hit = (torch.from_numpy(ops.order_tiles(recs)).to(self.dev), sum(sg[2] for sg in key_segs))
            self._tiles[key] = hit
        return hit

    def set_online(self, template, online_template):
        """VisionTransformer.set_online / Attention.set_online (mixformer_online.py:243-262, 96-113): ONE template and
        n online templates form one token sequence with full self-attention; every layer's qkv is kept."""
        t = self._check_img(template, self.template_size)
        ot = self._check_img(online_template, self.template_size)
        if t.shape[0] != 1:
            raise RuntimeError("set_online caches ONE sequence (template batch must be 1), like the reference "
                               "(x_ot.reshape(1, -1, C), mixformer_online.py:252)")
        T = self.gt * self.gt
        Tm = (1 + ot.shape[0]) * T
        bb = self.bbs[0]
        tag = ("online_t", Tm)
        x = self._buf(tag, "x", (Tm, self.dim), torch.float32)
        patches = self._embed_buf(Tm)
        self._stage_tokens(bb, t, patches, 0, Tm)
        self._stage_tokens(bb, ot, patches, T, T)
        pos_t = self._ws.get("pos_t")
        if pos_t is None:
            pos_t = bb["pos"][:T].contiguous()
            self._ws["pos_t"] = pos_t
            self._ws["pos_s"] = bb["pos"][2 * T:].contiguous()
        ops.gemm(patches, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, pos_t, out=x)
        self.qkv_mem = [self._buf(tag, f"qkv_mem{i}", (Tm, 3 * self.dim), self.act) for i in range(self.depth)]
        tiles = self._full_tiles(Tm, [(0, 0, Tm)], "set_online")
        for i, blk in enumerate(bb["blocks"]):
            self._block(blk, x, 1, Tm, 0, 0, False, tag, tiles=tiles, qkv_out=self.qkv_mem[i], first=(i == 0),
                        last=(i == self.depth - 1))
        self.templ_rows = ops.copy_rows(x, Tm, 0, T, 1, self._buf(tag, "templ_rows", (T, self.dim), self.act))
        self.mem_rows = Tm

    def forward_test(self, search, want_maps=True, run_score_head=True, gt_bboxes=None):
        """VisionTransformer.forward_test / Attention.forward_test (mixformer_online.py:229-241, 80-94): search tokens
        only; keys/values = cached template rows + own rows."""
        if self.qkv_mem is None:
            raise RuntimeError("forward_test called before set_online")
        s = self._check_img(search, self.search_size)
        if s.shape[0] != 1:
            raise RuntimeError("forward_test runs ONE search crop against the cached templates, like the reference")
        Ls, Tm, bb = self.Ls0, self.mem_rows, self.bbs[0]
        tag = ("online_s", Tm)
        x = self._buf(tag, "x", (Ls, self.dim), torch.float32)
        patches = self._embed_buf(Ls)
        self._stage_tokens(bb, s, patches, 0, Ls)
        ops.gemm(patches, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, self._ws["pos_s"], out=x)
        tiles = self._full_tiles(Ls, [(1, 0, Tm), (0, 0, Ls)], "forward_test")
        for i, blk in enumerate(bb["blocks"]):
            self._block(blk, x, 1, Ls, 0, 0, False, tag, tiles=tiles, qkv1=self.qkv_mem[i], first=(i == 0),
                        last=(i == self.depth - 1))
        feat = ops.copy_rows(x, Ls, 0, Ls, 1, self._buf(tag, "search_rows", (Ls, self.dim), self.act))
        boxes, maps = self._run_head(feat, 1, want_maps)
        res = dict(pred_boxes=boxes.view(1, 1, 4), score_maps=maps, feat_rows=feat)
        if run_score_head:
            res["pred_scores"] = self._scores(x, Ls, 0, 1, self.templ_rows, gt_bboxes)
        return res


    # ------------------------------------------------------------------------------------------ cached path, B sequences
    # The reference's set_online / forward_test assume ONE sequence (x_ot.reshape(1, -1, C), mixformer_online.py:252).
    # The same computation for B independent sequences at once: every sequence keeps its own (1 + n) * T cached template
    # rows per layer, its search tokens read that block and their own rows.  Per sequence the result is bit-identical to
    # the batch-1 path above (same 128-row tile boundaries inside a sequence, row-independent GEMMs).
    def set_online_batch(self, template, online_template):
        """template [B, 3, T, T]; online_template [B, n, 3, T, T] (n online templates per sequence)."""
        t = self._check_img(template, self.template_size)
        if online_template.dim() != 5 or online_template.shape[0] != t.shape[0]:
            raise RuntimeError(f"online_template must be [B, n, 3, T, T] with B = {t.shape[0]}, got {tuple(online_template.shape)}")
        B, n = online_template.shape[0], online_template.shape[1]
        T = self.gt * self.gt
        Tm = (1 + n) * T
        bb = self.bbs[0]
        tag = ("online_tb", B, Tm)
        x = self._buf(tag, "x", (B * Tm, self.dim), torch.float32)
        patches = self._embed_buf(B * Tm)
        self._stage_tokens(bb, t, patches, 0, Tm)
        for k in range(n):
            self._stage_tokens(bb, self._check_img(online_template[:, k], self.template_size), patches, (1 + k) * T, Tm)
        pos_t = self._ws.get("pos_t")
        if pos_t is None:
            pos_t = bb["pos"][:T].contiguous()
            self._ws["pos_t"] = pos_t
            self._ws["pos_s"] = bb["pos"][2 * T:].contiguous()
        ops.gemm(patches, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, pos_t, out=x)
        mem = [self._buf(tag, f"qkv_mem{i}", (B * Tm, 3 * self.dim), self.act) for i in range(self.depth)]
        tiles = self._seq_tiles(f"set_online_b{Tm}", B, Tm, lambda sq: [(0, sq * Tm, Tm)])
        for i, blk in enumerate(bb["blocks"]):
            self._block(blk, x, B, Tm, 0, 0, False, tag, tiles=tiles, qkv_out=mem[i], first=(i == 0), last=(i == self.depth - 1))
        templ = ops.copy_rows(x, Tm, 0, T, B, self._buf(tag, "templ_rows", (B * T, self.dim), self.act))
        self._batch_cache = dict(B=B, Tm=Tm, mem=mem, templ=templ)

    def forward_test_batch(self, search, want_maps=True, run_score_head=True, gt_bboxes=None):
        """search [B, 3, S, S] against the cache of set_online_batch (same B)."""
        c = getattr(self, "_batch_cache", None)
        if c is None:
            raise RuntimeError("forward_test_batch called before set_online_batch")
        s = self._check_img(search, self.search_size)
        B, Tm = c["B"], c["Tm"]
        if s.shape[0] != B:
            raise RuntimeError(f"cached templates are for {B} sequences, got {s.shape[0]} search crops")
        Ls, bb = self.Ls0, self.bbs[0]
        tag = ("online_sb", B, Tm)
        x = self._buf(tag, "x", (B * Ls, self.dim), torch.float32)
        patches = self._embed_buf(B * Ls)
        self._stage_tokens(bb, s, patches, 0, Ls)
        ops.gemm(patches, bb["pe_w"], bb["pe_b"], ops.ACT_NONE, None, self._ws["pos_s"], out=x)
        tiles = self._seq_tiles(f"forward_test_b{Tm}", B, Ls, lambda sq: [(1, sq * Tm, Tm), (0, sq * Ls, Ls)])
        for i, blk in enumerate(bb["blocks"]):
            self._block(blk, x, B, Ls, 0, 0, False, tag, tiles=tiles, qkv1=c["mem"][i], first=(i == 0),
                        last=(i == self.depth - 1))
        feat = ops.copy_rows(x, Ls, 0, Ls, B, self._buf(tag, "search_rows", (B * Ls, self.dim), self.act))
        boxes, maps = self._run_head(feat, B, want_maps)
        res = dict(pred_boxes=boxes.view(B, 1, 4), score_maps=maps, feat_rows=feat)
        if run_score_head:
            res["pred_scores"] = self._scores(x, Ls, 0, B, c["templ"], gt_bboxes)
        return res


class ConvMAEOnlineEngine(OnlineEngine):
    """mixformer_convmae_online: the same online engine behind the ConvMAE conv stem
    (lib/models/mixformer_convmae/mixformer_online.py:266-392): per crop, Conv 4x4/4 + LN + GELU, 2 CBlocks,
    Conv 2x2/2 + LN + GELU, 2 CBlocks, Conv 2x2/2 + LN + GELU, then patch_embed4 (Linear) + pos-embed as the
    token-embedding GEMM.  Stem residual streams are fp32 NHWC rows; every conv except the depthwise 5x5 is a GEMM."""

    def _pack(self, sd):
        super()._pack(sd)
        dev = self.dev
        g = lambda k: sd["backbone." + k]
        E0, E1 = self.stem_dims
        st = {}
        st["pe1_w"] = self._w(g("patch_embed1.proj.weight").reshape(E0, -1))                       # (c, ky, kx)
        st["pe2_w"] = self._w(g("patch_embed2.proj.weight").permute(0, 2, 3, 1).reshape(E1, -1))   # (ky, kx, c)
        st["pe3_w"] = self._w(g("patch_embed3.proj.weight").permute(0, 2, 3, 1).reshape(self.dim, -1))
        for j in (1, 2, 3):
            st[f"pe{j}_b"] = _f32(g(f"patch_embed{j}.proj.bias"), dev)
            st[f"pe{j}_ln"] = (_f32(g(f"patch_embed{j}.norm.weight"), dev), _f32(g(f"patch_embed{j}.norm.bias"), dev))
        for stage, E in ((1, E0), (2, E1)):
            blocks = []
            i = 0
            while f"backbone.blocks{stage}.{i}.conv1.weight" in sd:
                p = f"blocks{stage}.{i}."
                blocks.append(dict(
                    ln1=(_f32(g(p + "norm1.weight"), dev), _f32(g(p + "norm1.bias"), dev)),
                    ln2=(_f32(g(p + "norm2.weight"), dev), _f32(g(p + "norm2.bias"), dev)),
                    c1_w=self._w(g(p + "conv1.weight").reshape(E, E)), c1_b=_f32(g(p + "conv1.bias"), dev),
                    c2_w=self._w(g(p + "conv2.weight").reshape(E, E)), c2_b=_f32(g(p + "conv2.bias"), dev),
                    dw_w=_f32(g(p + "attn.weight").reshape(E, 25).t(), dev), dw_b=_f32(g(p + "attn.bias"), dev),
                    f1_w=self._w(g(p + "mlp.fc1.weight").reshape(-1, E)), f1_b=_f32(g(p + "mlp.fc1.bias"), dev),
                    f2_w=self._w(g(p + "mlp.fc2.weight").reshape(E, -1)), f2_b=_f32(g(p + "mlp.fc2.bias"), dev)))
                i += 1
            st[f"blocks{stage}"] = blocks
        self.stem = st

    def _embed_buf(self, rows, lane=0):
        return self._buf((rows, lane), "pe4_in", (rows, self.dim), self.act)

    def _ln_act(self, x, ln, gelu, out, **remap):
        if out.dtype == torch.float32:
            ops.layernorm_act(x, ln[0], ln[1], 1e-5, gelu, out_f32=out, **remap)
        else:
            ops.layernorm_act(x, ln[0], ln[1], 1e-5, gelu, out_bf16=out, **remap)

    def _cblock(self, blk, x, n, H, W, E, tag):
        """CBlock.forward (:181-189, mask=None): x += conv2(dw5x5(conv1(LN(x)))); x += fc2(GELU(fc1(LN(x))))."""
        rows = n * H * W
        h = self._buf(tag, f"h{E}", (rows, E), self.act)
        c1 = self._buf(tag, f"c1{E}", (rows, E), self.act)
        dw = self._buf(tag, f"dw{E}", (rows, E), self.act)
        self._ln_act(x, blk["ln1"], False, h)
        ops.gemm(h, blk["c1_w"], blk["c1_b"], out=c1)
        ops.dwconv5x5(c1, blk["dw_w"], blk["dw_b"], n, H, W, dw)
        ops.gemm(dw, blk["c2_w"], blk["c2_b"], ops.ACT_NONE, x, None, out=x)
        self._ln_act(x, blk["ln2"], False, h)
        hid = self._buf(tag, f"hid{E}", (rows, blk["f1_w"].shape[0]), self.act)
        ops.gemm(h, blk["f1_w"], blk["f1_b"], ops.ACT_GELU, out=hid)
        ops.gemm(hid, blk["f2_w"], blk["f2_b"], ops.ACT_NONE, x, None, out=x)

    def _stage_tokens(self, bb, img, buf, tok_off, tok_per_seq):
        st = self.stem
        E0, E1 = self.stem_dims
        n, S = img.shape[0], img.shape[2]
        H1, H2, H3 = S // 4, S // 8, S // 16
        tag = ("stem", n, S)
        p1 = self._buf(tag, "p1", (n * H1 * H1, 48), self.act)
        ops.patchify(img, p1, 0, H1 * H1, patch=4)
        y1 = self._buf(tag, "y1", (n * H1 * H1, E0), torch.float32)
        ops.gemm(p1, st["pe1_w"], st["pe1_b"], out=y1)
        self._ln_act(y1, st["pe1_ln"], True, y1)
        for blk in st["blocks1"]:
            self._cblock(blk, y1, n, H1, H1, E0, tag)
        p2 = ops.patchify2x2(y1, n, H1, H1, self._buf(tag, "p2", (n * H2 * H2, 4 * E0), self.act))
        y2 = self._buf(tag, "y2", (n * H2 * H2, E1), torch.float32)
        ops.gemm(p2, st["pe2_w"], st["pe2_b"], out=y2)
        self._ln_act(y2, st["pe2_ln"], True, y2)
        for blk in st["blocks2"]:
            self._cblock(blk, y2, n, H2, H2, E1, tag)
        p3 = ops.patchify2x2(y2, n, H2, H2, self._buf(tag, "p3", (n * H3 * H3, 4 * E1), self.act))
        y3 = self._buf(tag, "y3", (n * H3 * H3, self.dim), torch.float32)
        ops.gemm(p3, st["pe3_w"], st["pe3_b"], out=y3)
        # LN + GELU of patch_embed3, written straight into the token order of the embedding GEMM (patch_embed4)
        self._ln_act(y3, st["pe3_ln"], True, buf, seg_rows=H3 * H3, out_seq_rows=tok_per_seq, out_row_off=tok_off)


class AsymOnlineEngine(OnlineEngine):
    """asymmetric_shared backbone + fusion + corner head (engine.ForwardEngine, batch-stacked modalities, cross-modal
    attention) with the SPM score head on top (lib/models/mixformer_vit_rgbt/asymmetric_shared_online.py:352-401): the
    score decoder pools the FUSED search map at the predicted box and attends the first-template tokens of BOTH
    modalities (`torch.cat(torch.split(template, [N, N], dim=0), dim=2)`: RGB tokens first, then infrared)."""

    def _spm_tiles(self, B, T):
        """One query row per sequence; keys = the sequence's 16 pooled tokens (T = 16) or its two template segments
        (T = 2 * gt^2: rows [b*Tm, +Tm) of the RGB half and of the infrared half of the stacked template buffer)."""
        Tm = self.gt * self.gt
        if T != 2 * Tm:
            return super()._spm_tiles(B, T)
        key = ("spm2", B, T)
        hit = self._tiles.get(key)
        if hit is None:
            recs = [[b, 1, b, 2, b * Tm, (B + b) * Tm, 0, Tm, Tm, 0, 1, 1, 0, 0, 0, 0] for b in range(B)]
            hit = (torch.from_numpy(np.asarray(recs, dtype=np.int32)).to(self.dev), T)
            self._tiles[key] = hit
        return hit

    def forward(self, template, online_template, search, want_maps=True, run_score_head=False, gt_bboxes=None,
                ready_events=None):
        res = ForwardEngine.forward(self, template, online_template, search, want_maps=want_maps,
                                    ready_events=ready_events)
        if not run_score_head:
            return res
        B = res["pred_boxes"].shape[0]
        HW, C, Tm = self.Ls0, self.dim, self.gt * self.gt
        tag = ("spm", B)
        # fused search map as fp32 rows: the GroupNorm of fusion_vi's output conv once more, from its fp32 pre-activation
        # (bf16 mode keeps only a bf16 copy of the fused map; PrRoIPool integrates fp32 like the reference)
        F_ = self.fusion
        if "cat_convs" in F_:
            raise NotImplementedError("asymmetric_shared_online with RGBT_Fusion_Cat: no shipped YAML uses it")
        pre = self._buf((("fus", B), "o"), "gn_pre", (B * HW, 768), torch.float32)
        feat32 = self._buf(tag, "feat32", (B * HW, C), torch.float32)
        ops.groupnorm(pre, B, HW, 32, F_["out"]["gn_w"], F_["out"]["gn_b"], 1e-5, out_f32=feat32)
        # first-template rows of both modalities out of the residual stream: [RGB x B | infrared x B] blocks of Tm rows
        x, N = self._last_x
        templ = self._buf(tag, "templ2", (2 * B * Tm, C), self.act)
        ops.copy_rows(x, N, 0, Tm, 2 * B, templ)
        if gt_bboxes is not None:
            xyxy = gt_bboxes.detach().to(device=self.dev, dtype=torch.float32).reshape(B, 4).contiguous()
        else:
            xyxy = self._last_xyxy
        res["pred_scores"] = self._run_spm(feat32, templ, xyxy, B)
        return res

    def set_online(self, *a, **k):
        # the reference's own set_online of this class references attributes that do not exist (backbone_v / backbone_i,
        # asymmetric_shared_online.py:386-387): the RGB-T models support the full forward only
        raise NotImplementedError("asymmetric_shared_online supports the full forward only (as in the reference)")

    forward_test = set_online
