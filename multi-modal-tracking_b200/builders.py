"""Drop-in model builders: same names, signatures, config objects and checkpoint layout as the reference's
`lib/models` builders, with the forward executed by the sm_100a engine (engine.py) instead of torch ops.

    build_mixformer_vit(cfg, train=False)                lib/models/mixformer_vit/mixformer.py:341-366
    build_mixformer_vit_rgbt(cfg, train=False)           lib/models/mixformer_vit_rgbt/mixformer.py:433-466
    build_mixformer_vit_rgbt_shared(cfg, train=False)    lib/models/mixformer_vit_rgbt/mixformer_shared.py:461-507
    build_mixformer_vit_rgbt_uni(cfg, train=False)       lib/models/mixformer_vit_rgbt/mixformer_unibackbone.py
    build_asymmetric_shared(cfg, train=False)            lib/models/mixformer_vit_rgbt/asymmetric_shared.py
    build_asymmetric_shared_ce(cfg, train=False)         lib/models/mixformer_vit_rgbt/asymmetric_shared_ce.py:590-640
    build_mixformer_vit_online_score(cfg, train=False)   lib/models/mixformer_vit/mixformer_online.py:363-385
        (SPM score head + cached-template set_online / forward_test)
    build_mixformer_convmae_online_score(cfg, train=False)  lib/models/mixformer_convmae/mixformer_online.py:506-526

The returned nn.Module owns nn.Parameters / buffers under EXACTLY the reference's state_dict keys and shapes, so
`load_state_dict(torch.load(ckpt)["net"], strict=True)` works on reference checkpoints.  The torch sub-modules
(nn.Linear, nn.Conv2d, ...) are used purely as named parameter containers - their forward is never called.
`forward(template, online_template, search, ...)` returns `(out_dict, coords)` like the reference.
Only inference is supported (train=True raises): training is out of scope of this library.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .pos_embed import sincos_pos_embed_2d

VIT_DIMS = {"base_patch16": dict(dim=768, depth=12, heads=12), "large_patch16": dict(dim=1024, depth=24, heads=16)}


# ------------------------------------------------------------------------------------------------ containers
class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    """Parameter layout of Block (mixformer.py:112-124) or, with per_modality_ln, Block_Shared /
    CE_Block_Shared (mixformer_shared.py:113-141, asymmetric_shared_ce.py:210-245)."""

    def __init__(self, dim, per_modality_ln):
        super().__init__()
        ln = partial(nn.LayerNorm, eps=1e-6)
        if per_modality_ln:
            self.norm1_v, self.norm1_i = ln(dim), ln(dim)
        else:
            self.norm1 = ln(dim)
        self.attn = _Attention(dim)
        if per_modality_ln:
            self.norm2_v, self.norm2_i = ln(dim), ln(dim)
        else:
            self.norm2 = ln(dim)
        self.mlp = _Mlp(dim, dim * 4)


class _PatchEmbed(nn.Module):
    def __init__(self, dim, patch=16):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class _Backbone(nn.Module):
    """Parameter layout of the MixViT VisionTransformer subclasses.  `timm_leftovers` keeps the unused timm
    base-class parameters that RGB-only checkpoints still carry (cls_token, pos_embed, norm, head: SURVEY
    section 7 'State-dict fidelity'); the RGB-T builders drop them (mixformer_vit_rgbt/mixformer.py:329-333)."""

    def __init__(self, vit_type, img_size_s, img_size_t, per_modality_ln=False, timm_leftovers=False):
        super().__init__()
        if vit_type not in VIT_DIMS:
            raise KeyError("VIT_TYPE shoule set to 'large_patch16' or 'base_patch16'")
        d = VIT_DIMS[vit_type]
        dim = d["dim"]
        self.embed_dim, self.depth, self.num_heads = dim, d["depth"], d["heads"]
        if timm_leftovers:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
            self.pos_embed = nn.Parameter(torch.zeros(1, (224 // 16) ** 2 + 1, dim))
            self.norm = nn.LayerNorm(dim, eps=1e-6)
            self.head = nn.Linear(dim, 1000)
        self.patch_embed = _PatchEmbed(dim)
        self.blocks = nn.Sequential(*[_Block(dim, per_modality_ln) for _ in range(d["depth"])])
        self.grid_size_s, self.grid_size_t = img_size_s // 16, img_size_t // 16
        self.pos_embed_s = nn.Parameter(torch.zeros(1, self.grid_size_s ** 2, dim), requires_grad=False)
        self.pos_embed_t = nn.Parameter(torch.zeros(1, self.grid_size_t ** 2, dim), requires_grad=False)
        self.pos_embed_s.data.copy_(sincos_pos_embed_2d(dim, self.grid_size_s).unsqueeze(0))
        self.pos_embed_t.data.copy_(sincos_pos_embed_2d(dim, self.grid_size_t).unsqueeze(0))
        for m in self.modules():      # timm init_weights(''): trunc_normal(.02) linears, unit LayerNorms
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


class _CMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Conv2d(dim, hidden, 1)
        self.fc2 = nn.Conv2d(hidden, dim, 1)


class _CBlock(nn.Module):
    """Parameter layout of CBlock (lib/models/mixformer_convmae/mixformer_online.py:167-180)."""

    def __init__(self, dim, mlp_ratio=4):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.conv1 = nn.Conv2d(dim, dim, 1)
        self.conv2 = nn.Conv2d(dim, dim, 1)
        self.attn = nn.Conv2d(dim, dim, 5, padding=2, groups=dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _CMlp(dim, int(dim * mlp_ratio))


class _ConvPatchEmbed(nn.Module):
    def __init__(self, patch, inp, dim):
        super().__init__()
        self.proj = nn.Conv2d(inp, dim, kernel_size=patch, stride=patch)
        self.norm = nn.LayerNorm(dim)


class _ConvViT(nn.Module):
    """Parameter layout of ConvViT (lib/models/mixformer_convmae/mixformer_online.py:192-262)."""
    DIMS = {"convmae_base": ((256, 384, 768), (2, 2, 11), 12), "convmae_large": ((384, 768, 1024), (2, 2, 20), 16)}

    def __init__(self, vit_type, img_size_s, img_size_t):
        super().__init__()
        if vit_type not in self.DIMS:
            raise KeyError("VIT_TYPE shoule set to 'convmae_base' or 'convmae_large'")
        (e0, e1, e2), depth, heads = self.DIMS[vit_type]
        self.embed_dim, self.num_heads = e2, heads
        self.patch_embed1 = _ConvPatchEmbed(4, 3, e0)
        self.patch_embed2 = _ConvPatchEmbed(2, e0, e1)
        self.patch_embed3 = _ConvPatchEmbed(2, e1, e2)
        self.patch_embed4 = nn.Linear(e2, e2)
        self.blocks1 = nn.ModuleList([_CBlock(e0) for _ in range(depth[0])])
        self.blocks2 = nn.ModuleList([_CBlock(e1) for _ in range(depth[1])])
        self.blocks3 = nn.ModuleList([_Block(e2, False) for _ in range(depth[2])])
        self.norm = nn.LayerNorm(e2, eps=1e-6)
        for m in self.modules():      # ConvViT._init_weights :236-244
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        self.grid_size_s, self.grid_size_t = img_size_s // 16, img_size_t // 16
        self.pos_embed_s = nn.Parameter(torch.zeros(1, self.grid_size_s ** 2, e2), requires_grad=False)
        self.pos_embed_t = nn.Parameter(torch.zeros(1, self.grid_size_t ** 2, e2), requires_grad=False)
        self.pos_embed_s.data.copy_(sincos_pos_embed_2d(e2, self.grid_size_s).unsqueeze(0))
        self.pos_embed_t.data.copy_(sincos_pos_embed_2d(e2, self.grid_size_t).unsqueeze(0))


class FrozenBatchNorm2d(nn.Module):
    """Buffers of lib/models/mixformer_cvt/utils.py:21-45 (no num_batches_tracked)."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


def _conv(inp, out, freeze_bn):
    return nn.Sequential(nn.Conv2d(inp, out, kernel_size=3, padding=1, bias=True),
                         FrozenBatchNorm2d(out) if freeze_bn else nn.BatchNorm2d(out), nn.ReLU(inplace=True))


class _CornerHead(nn.Module):
    """Parameter layout of Corner_Predictor / Pyramid_Corner_Predictor (lib/models/mixformer_cvt/head.py:23-52,
    97-145).  Unlike the reference it can be constructed without CUDA."""

    def __init__(self, inplanes, channel, feat_sz, stride, freeze_bn, pyramid):
        super().__init__()
        self.feat_sz, self.stride, self.img_sz, self.pyramid = feat_sz, stride, feat_sz * stride, pyramid
        for c in ("tl", "br"):
            setattr(self, f"conv1_{c}", _conv(inplanes, channel, freeze_bn))
            setattr(self, f"conv2_{c}", _conv(channel, channel // 2, freeze_bn))
            setattr(self, f"conv3_{c}", _conv(channel // 2, channel // 4, freeze_bn))
            setattr(self, f"conv4_{c}", _conv(channel // 4, channel // 8, freeze_bn))
            setattr(self, f"conv5_{c}", nn.Conv2d(channel // 8, 1, kernel_size=1))
            if pyramid:
                setattr(self, f"adjust1_{c}", _conv(inplanes, channel // 2, freeze_bn))
                setattr(self, f"adjust2_{c}", _conv(inplanes, channel // 4, freeze_bn))
                setattr(self, f"adjust3_{c}", nn.Sequential(_conv(channel // 2, channel // 4, freeze_bn),
                                                            _conv(channel // 4, channel // 8, freeze_bn),
                                                            _conv(channel // 8, 1, freeze_bn)))
                setattr(self, f"adjust4_{c}", nn.Sequential(_conv(channel // 4, channel // 8, freeze_bn),
                                                            _conv(channel // 8, 1, freeze_bn)))


def build_box_head(cfg):
    """build_box_head lib/models/mixformer_cvt/head.py:235-258 (corner heads only)."""
    if "CORNER" not in cfg.MODEL.HEAD_TYPE:
        raise ValueError("HEAD TYPE %s is not supported." % cfg.MODEL.HEAD_TYPE)
    channel = cfg.MODEL.get("HEAD_DIM", 384)
    freeze_bn = cfg.MODEL.get("HEAD_FREEZE_BN", False)
    if cfg.MODEL.HEAD_TYPE == "CORNER":
        stride, pyramid = 16, False
    elif cfg.MODEL.HEAD_TYPE == "CORNER_UP":
        stride, pyramid = 4, True
    else:
        raise ValueError()
    return _CornerHead(cfg.MODEL.HIDDEN_DIM, channel, int(cfg.DATA.SEARCH.SIZE / stride), stride, freeze_bn, pyramid)


class _MSDeformAttnBimodal(nn.Module):
    def __init__(self, d_model=512, n_levels=2, n_heads=8, n_points=4):
        super().__init__()
        self.sampling_offsets = nn.Linear(2 * d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(2 * d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        # MSDeformAttn_Bimodal._reset_parameters ms_deform_attn_bimodal.py:64-81
        import math
        nn.init.zeros_(self.sampling_offsets.weight)
        thetas = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
        for i in range(n_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.zeros_(self.value_proj.bias)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.zeros_(self.output_proj.bias)


class _FusionEncoderLayer(nn.Module):
    def __init__(self, d_model, d_ffn, ln_specific=True):
        super().__init__()
        self.self_attn = _MSDeformAttnBimodal(d_model)
        if ln_specific:
            self.norm1_v, self.norm1_i = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        else:           # DeformableTransformerEncoderLayer deformable_encoder.py:111-137: one norm for both modalities
            self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        if ln_specific:
            self.norm2_v, self.norm2_i = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        else:
            self.norm2 = nn.LayerNorm(d_model)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)


class _FusionEncoder(nn.Module):
    def __init__(self, d_model, layers, ln_specific=True):
        super().__init__()
        self.layers = nn.ModuleList([_FusionEncoderLayer(d_model, 4 * d_model, ln_specific) for _ in range(layers)])


class _FusionAttention(nn.Module):
    """Parameter layout of DeformableAttentionFusion_LNSpecific (deformable_encoder_lnspecific.py:23-57)."""

    def __init__(self, d_model, layers, ln_specific=True):
        super().__init__()
        self.encoder = _FusionEncoder(d_model, layers, ln_specific)
        self.level_embed = nn.Parameter(torch.randn(2, d_model))


def _conv_gn(inp, out):
    return nn.Sequential(nn.Conv2d(inp, out, kernel_size=1), nn.GroupNorm(32, out))


FUSION_CLASSES = ("Attention_Fusion_Bimodal_LNSpecific", "Attention_Fusion_Bimodal_LNSpecific_Sum",
                  "Attention_Fusion_Bimodal_LNSpecific_2", "Attention_Fusion_Bimodal", "RGBT_Fusion_Cat")


class _FusionCat(nn.Module):
    """Parameter layout of RGBT_Fusion_Cat (fusion_utils.py:86-110): three bias-free 3x3 convs + BatchNorm + ReLU on
    the channel concatenation of the two modalities."""

    def __init__(self, channels=768):
        super().__init__()
        self.fusion_class = "RGBT_Fusion_Cat"
        for j, (i, o) in enumerate(((2 * channels, 2 * channels), (2 * channels, channels), (channels, channels)), 1):
            setattr(self, f"fusion{j}", nn.Conv2d(i, o, kernel_size=3, stride=1, padding=1, bias=False))
            setattr(self, f"fusion{j}_bn", nn.BatchNorm2d(o))


class _Fusion(nn.Module):
    """Parameter layout of Attention_Fusion_Bimodal_LNSpecific{,_Sum,_2} (fusion_utils.py:243-353)."""

    def __init__(self, fusion_class, channels=768, d_model=512, layers=2):
        super().__init__()
        if fusion_class == "Attention_Fusion_512":
            # The reference's own builders cannot construct this class either: they call
            # globals()[cfg.MODEL.FUSION_CLASS](768, d_model=512, num_feature_levels=2, num_encoder_layers=...)
            # (asymmetric_shared.py:418, mixformer_shared.py:474, ...) and Attention_Fusion_512.__init__
            # (fusion_utils.py:128-150) takes no num_encoder_layers - the same TypeError, verbatim (tests/test_boundary.py
            # pins it against the reference's builder where the reference tree is present).
            raise TypeError("Attention_Fusion_512.__init__() got an unexpected keyword argument 'num_encoder_layers'")
        if fusion_class not in FUSION_CLASSES:
            raise KeyError(f"FUSION_CLASS {fusion_class!r} is not on the accelerated path; supported: {FUSION_CLASSES}")
        self.fusion_class, self.d_model = fusion_class, d_model
        if fusion_class.endswith("_2"):
            self.adjust_in = _conv_gn(channels, d_model)
        else:
            self.adjust_v = _conv_gn(channels, d_model)
            self.adjust_i = _conv_gn(channels, d_model)
        self.fusion_attention = _FusionAttention(d_model, layers, ln_specific="LNSpecific" in fusion_class)
        if fusion_class.endswith("_Sum"):
            self.adjust_sum = _conv_gn(d_model, channels)
        elif fusion_class.endswith("_2"):
            self.adjust_out = _conv_gn(d_model, channels)
        else:
            self.adjust_cat = _conv_gn(2 * d_model, channels)


# ------------------------------------------------------------------------------------------------ models
class _EngineModule(nn.Module):
    """Common behaviour: lazily (re)pack weights into the engine's device arena, run the engine forward."""

    variant = ""
    precision = "bf16"      # "bf16" (tcgen05 path) or "fp32" (parity mode)

    def _init_engine_state(self, cfg):
        self.head_type = cfg.MODEL.HEAD_TYPE
        self._cfg = cfg
        self._engine = None
        self._packed_version = -1
        self._version = 0
        self._use_graph = False
        self._graphs = {}

    def enable_cuda_graph(self, on: bool = True):
        """Replay the whole forward (~240 kernel launches) as ONE CUDA graph per batch size: inputs are copied into
        static device buffers, the graph is replayed on the current stream and fresh copies of the boxes are returned.
        Removes the per-launch host cost, which is what bounds the bs=1 per-frame latency."""
        self._use_graph = bool(on)
        if not on:
            self._graphs = {}
        return self

    def _graphed_forward(self, template, online_template, search):
        eng = self.engine()
        flat = lambda a: list(a) if isinstance(a, (list, tuple)) else [a]
        ins = flat(template) + flat(online_template) + flat(search)
        key = (id(eng), tuple(tuple(t.shape) for t in ins))
        g = self._graphs.get(key)
        if g is None:
            for t in ins:
                if not t.is_cuda:
                    raise NotImplementedError("mmt_b200 forward is CUDA-only (no CPU fallback)")
            static = [torch.empty_like(t, dtype=torch.float32).contiguous() for t in ins]
            for d, t in zip(static, ins):
                d.copy_(t)
            n = len(flat(template))
            regroup = (lambda xs: xs) if isinstance(template, (list, tuple)) else (lambda xs: xs[0])
            args = (regroup(static[:n]), regroup(static[n:2 * n]), regroup(static[2 * n:]))
            eng.forward(*args, want_maps=False)          # warm-up: workspaces, function attributes, tile tables
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                res = eng.forward(*args, want_maps=False)
            g = (graph, static, res)
            self._graphs[key] = g
        graph, static, res = g
        for d, t in zip(static, ins):
            d.copy_(t, non_blocking=True)
        graph.replay()
        out = dict(res)
        out["pred_boxes"] = res["pred_boxes"].clone()
        return out

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._version += 1
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._version += 1      # .cuda()/.to(): weights moved, arena must be rebuilt
        return r

    def invalidate(self):
        """Re-pack the device weight arena (and drop captured CUDA graphs) at the next forward.  load_state_dict on the
        model, .cuda() / .to() and set_precision do this by themselves; call it after what they cannot see: in-place edits
        of parameters (`p.data.copy_()`), or load_state_dict on a SUB-module (`model.backbone.load_state_dict(...)`).
        (An automatic check - comparing every parameter's version counter per forward - was measured at 0.5 ms of host time
        per call, 40 % of the single-sequence latency, and is therefore not done.)"""
        self._version += 1
        return self

    def set_precision(self, precision: str):
        if precision not in ("bf16", "fp32"):
            raise ValueError(precision)
        if precision != self.precision:
            self.precision = precision
            self._version += 1
        return self

    def engine(self):
        from .engine import ForwardEngine
        if self.variant == "mixformer_vit_online":
            from .engine_online import OnlineEngine as ForwardEngine
        elif self.variant == "mixformer_convmae_online":
            from .engine_online import ConvMAEOnlineEngine as ForwardEngine
        elif self.variant == "asymmetric_shared_online":
            from .engine_online import AsymOnlineEngine as ForwardEngine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            # same convention as the reference's native ops (prroi_pool/functional.py:62-63)
            raise NotImplementedError("mmt_b200 models run on CUDA only: call .cuda() first (no CPU fallback)")
        if self._engine is None or self._packed_version != self._version:
            self._graphs = {}
            self._engine = ForwardEngine(self.variant, self._cfg, self.state_dict(), dev, self.precision)
            self._packed_version = self._version
        return self._engine

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("mmt_b200 provides the inference forward only (training is out of scope)")
        return super().train(False)

    # Template-side reuse (an extension, not a reference API): the trackers re-send
    # unchanged templates with every frame (lib/test/tracker/mixformer_vit.py:71); cache_templates() once per template
    # update + forward_search() per frame give the same boxes as forward() for 72 % of the token rows.
    @torch.no_grad()
    def cache_templates(self, template, online_template):
        sq = lambda t: t.squeeze(0) if (torch.is_tensor(t) and t.dim() == 5) else t
        self.engine().cache_templates(sq(template), sq(online_template))

    @torch.no_grad()
    def forward_search(self, search):
        sq = lambda t: t.squeeze(0) if (torch.is_tensor(t) and t.dim() == 5) else t
        res = self.engine().forward_search(sq(search))
        return {"pred_boxes": res["pred_boxes"]}, res["pred_boxes"]

    def _finish(self, res, return_features=False):
        coords = res["pred_boxes"]
        out_dict = {"pred_boxes": coords}
        if return_features:
            eng, B = self._engine, coords.shape[0]
            sv, si = res["search_rows"]
            return out_dict, coords, eng.rows_to_map(sv, B), eng.rows_to_map(si, B)
        return out_dict, coords


class MixFormer(_EngineModule):
    """RGB-only MixViT tracker (lib/models/mixformer_vit/mixformer.py:285-338)."""
    variant = "mixformer_vit"

    def __init__(self, backbone, box_head, cfg):
        super().__init__()
        self.backbone, self.box_head = backbone, box_head
        self._init_engine_state(cfg)

    @torch.no_grad()
    def forward(self, template, online_template, search, run_score_head=False, gt_bboxes=None):
        sq = lambda t: t.squeeze(0) if t.dim() == 5 else t
        if self._use_graph:
            return self._finish(self._graphed_forward(sq(template), sq(online_template), sq(search)))
        return self._finish(self.engine().forward(sq(template), sq(online_template), sq(search)))

    def forward_box_head(self, search):
        res = self.engine().forward_head_only(search)
        return {"pred_boxes": res["pred_boxes"]}, res["pred_boxes"]


class MixFormer_RGBT(_EngineModule):
    """RGB-T trackers: two-stream (lib/models/mixformer_vit_rgbt/mixformer.py:350-431) when built with two
    backbones, batch-stacked shared-backbone family otherwise (mixformer_shared.py:385-459,
    asymmetric_shared_ce.py:540-588)."""

    def __init__(self, variant, backbone, box_head, fusion_vi, cfg):
        super().__init__()
        self.variant = variant
        if isinstance(backbone, (list, tuple)):
            self.backbone_v, self.backbone_i = backbone
        else:
            self.backbone = backbone
        self.fusion_vi, self.box_head = fusion_vi, box_head
        self._init_engine_state(cfg)

    @torch.no_grad()
    def forward(self, template, online_template, search, run_score_head=False, gt_bboxes=None,
                ce_template_mask=None, ce_keep_rate=None, return_features=False, ready_events=None):
        if ce_template_mask is not None or ce_keep_rate is not None:
            # training-time CE schedule / template mask (lib/utils/ce_utils.py); the test-time trackers never
            # pass them (lib/test/tracker/asymmetric_shared_ce.py:96-98)
            raise NotImplementedError("ce_template_mask / ce_keep_rate are training-only arguments")
        if self._use_graph and not return_features:
            for e in (ready_events or ()):
                torch.cuda.current_stream().wait_event(e)
            return self._finish(self._graphed_forward(list(template), list(online_template), list(search)))
        res = self.engine().forward(list(template), list(online_template), list(search), ready_events=ready_events)
        return self._finish(res, return_features)


class _MLP(nn.Module):
    """Parameter layout of MLP (lib/models/mixformer_cvt/head.py:215-232, BN=False)."""

    def __init__(self, inp, hidden, out, num_layers):
        super().__init__()
        h = [hidden] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([inp] + h, h + [out]))


class _ScoreDecoder(nn.Module):
    """Parameter layout of ScoreDecoder, the SPM (lib/models/mixformer_cvt/score_decoder.py:12-30)."""

    def __init__(self, num_heads, hidden_dim, nlayer_head=3, pool_size=4):
        super().__init__()
        self.num_heads, self.pool_size = num_heads, pool_size
        self.score_head = _MLP(hidden_dim, hidden_dim, 1, nlayer_head)
        lin = lambda: nn.ModuleList(nn.Linear(hidden_dim, hidden_dim, bias=True) for _ in range(2))
        self.proj_q, self.proj_k, self.proj_v, self.proj = lin(), lin(), lin(), lin()
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.norm2 = nn.ModuleList(nn.LayerNorm(hidden_dim) for _ in range(2))
        self.score_token = nn.Parameter(torch.zeros(1, 1, hidden_dim))
        nn.init.trunc_normal_(self.score_token, std=.02)


class MixFormerOnlineScore(_EngineModule):
    """MixViT tracker with the SPM score head and the cached-template test path
    (lib/models/mixformer_vit/mixformer_online.py:286-360).  forward() is the full forward (+ pred_scores);
    set_online(template, online_template) caches the templates' per-layer K/V and the template feature;
    forward_test(search) runs the search tokens only against that cache.  Like the reference, the cached path is
    per sequence: set_online takes ONE template and n online templates, forward_test ONE search crop."""
    variant = "mixformer_vit_online"

    def __init__(self, backbone, box_head, score_branch, cfg):
        super().__init__()
        self.backbone, self.box_head, self.score_branch = backbone, box_head, score_branch
        self._init_engine_state(cfg)

    @staticmethod
    def _sq(t):
        return t.squeeze(0) if t.dim() == 5 else t

    def _finish_online(self, res, run_score_head):
        coords = res["pred_boxes"]
        out = {"pred_boxes": coords}
        if run_score_head:
            out["pred_scores"] = res["pred_scores"].view(-1)
        return out, coords

    @torch.no_grad()
    def forward(self, template, online_template, search, run_score_head=True, gt_bboxes=None):
        res = self.engine().forward(self._sq(template), self._sq(online_template), self._sq(search),
                                    run_score_head=run_score_head, gt_bboxes=gt_bboxes)
        return self._finish_online(res, run_score_head)

    @torch.no_grad()
    def set_online(self, template, online_template):
        self.engine().set_online(self._sq(template), self._sq(online_template))

    @torch.no_grad()
    def forward_test(self, search, run_score_head=True, gt_bboxes=None):
        res = self.engine().forward_test(self._sq(search), run_score_head=run_score_head, gt_bboxes=gt_bboxes)
        return self._finish_online(res, run_score_head)

    # Batched form of the cached path (an extension: the reference's set_online / forward_test hold ONE sequence):
    # template [B,3,T,T], online_template [B,n,3,T,T], search [B,3,S,S]; per sequence bit-identical to the calls above.
    @torch.no_grad()
    def set_online_batch(self, template, online_template):
        self.engine().set_online_batch(template, online_template)

    @torch.no_grad()
    def forward_test_batch(self, search, run_score_head=True, gt_bboxes=None):
        res = self.engine().forward_test_batch(search, run_score_head=run_score_head, gt_bboxes=gt_bboxes)
        return self._finish_online(res, run_score_head)


def _require_inference(train):
    if train:
        raise NotImplementedError("mmt_b200 builders construct inference models only: call build_*(cfg, train=False)")


def build_mixformer_vit(cfg, train=False) -> MixFormer:
    _require_inference(train)
    backbone = _Backbone(cfg.MODEL.VIT_TYPE, cfg.DATA.SEARCH.SIZE, cfg.DATA.TEMPLATE.SIZE, timm_leftovers=True)
    return MixFormer(backbone, build_box_head(cfg), cfg).eval()


def _fusion(cfg):
    if cfg.MODEL.FUSION_CLASS == "RGBT_Fusion_Cat":
        return _FusionCat(768)
    return _Fusion(cfg.MODEL.FUSION_CLASS, 768, 512, cfg.MODEL.FUSION_LAYERS)


def build_mixformer_vit_rgbt(cfg, train=False) -> MixFormer_RGBT:
    _require_inference(train)
    mk = lambda: _Backbone(cfg.MODEL.VIT_TYPE, cfg.DATA.SEARCH.SIZE, cfg.DATA.TEMPLATE.SIZE)
    return MixFormer_RGBT("mixformer_vit_rgbt", [mk(), mk()], build_box_head(cfg), _fusion(cfg), cfg).eval()


def _build_stacked(variant, cfg, train, per_modality_ln):
    _require_inference(train)
    bb = _Backbone(cfg.MODEL.VIT_TYPE, cfg.DATA.SEARCH.SIZE, cfg.DATA.TEMPLATE.SIZE, per_modality_ln=per_modality_ln)
    return MixFormer_RGBT(variant, bb, build_box_head(cfg), _fusion(cfg), cfg).eval()


def build_mixformer_vit_rgbt_shared(cfg, train=False) -> MixFormer_RGBT:
    return _build_stacked("mixformer_vit_rgbt_shared", cfg, train, True)


def build_mixformer_vit_rgbt_uni(cfg, train=False) -> MixFormer_RGBT:
    return _build_stacked("mixformer_vit_rgbt_unibackbone", cfg, train, False)


def build_asymmetric_shared(cfg, train=False) -> MixFormer_RGBT:
    return _build_stacked("asymmetric_shared", cfg, train, True)


def build_asymmetric_shared_ce(cfg, train=False) -> MixFormer_RGBT:
    return _build_stacked("asymmetric_shared_ce", cfg, train, True)


class MixFormer_RGBT_OnlineScore(MixFormer_RGBT):
    """asymmetric_shared + the SPM score head (lib/models/mixformer_vit_rgbt/asymmetric_shared_online.py:337-413):
    the score decoder reads the FUSED search map and the first-template tokens of both modalities."""

    def __init__(self, backbone, box_head, fusion_vi, score_branch, cfg):
        super().__init__("asymmetric_shared_online", backbone, box_head, fusion_vi, cfg)
        self.score_branch = score_branch

    @torch.no_grad()
    def forward(self, template, online_template, search, run_score_head=False, gt_bboxes=None, return_features=False,
                ready_events=None):
        """ready_events: see ForwardEngine.forward (runner.FrameStep passes the per-modality upload events)."""
        res = self.engine().forward(list(template), list(online_template), list(search), run_score_head=run_score_head,
                                    gt_bboxes=gt_bboxes, ready_events=ready_events)
        coords = res["pred_boxes"]
        out = {"pred_boxes": coords}
        if run_score_head:
            out["pred_scores"] = res["pred_scores"].view(-1)
        if return_features:
            eng, B = self._engine, coords.shape[0]
            sv, si = res["search_rows"]
            return out, coords, eng.rows_to_map(sv, B), eng.rows_to_map(si, B), eng.rows_to_map(res["feat_rows"], B)
        return out, coords


def build_asymmetric_shared_online_score(cfg, train=False) -> MixFormer_RGBT_OnlineScore:
    _require_inference(train)
    bb = _Backbone(cfg.MODEL.VIT_TYPE, cfg.DATA.SEARCH.SIZE, cfg.DATA.TEMPLATE.SIZE, per_modality_ln=True)
    score_branch = _ScoreDecoder(num_heads=cfg.MODEL.HIDDEN_DIM // 64, hidden_dim=cfg.MODEL.HIDDEN_DIM, pool_size=4)
    return MixFormer_RGBT_OnlineScore(bb, build_box_head(cfg), _fusion(cfg), score_branch, cfg).eval()


def build_mixformer_vit_online_score(cfg, settings=None, train=False) -> MixFormerOnlineScore:
    _require_inference(train)
    backbone = _Backbone(cfg.MODEL.VIT_TYPE, cfg.DATA.SEARCH.SIZE, cfg.DATA.TEMPLATE.SIZE, timm_leftovers=True)
    score_branch = _ScoreDecoder(num_heads=cfg.MODEL.HIDDEN_DIM // 64, hidden_dim=cfg.MODEL.HIDDEN_DIM, pool_size=4)
    return MixFormerOnlineScore(backbone, build_box_head(cfg), score_branch, cfg).eval()


class MixFormerConvMAEOnlineScore(MixFormerOnlineScore):
    """ConvMAE-backbone online tracker (lib/models/mixformer_convmae/mixformer_online.py:427-504)."""
    variant = "mixformer_convmae_online"


def build_mixformer_convmae_online_score(cfg, settings=None, train=False) -> MixFormerConvMAEOnlineScore:
    """lib/models/mixformer_convmae/mixformer_online.py:506-526."""
    _require_inference(train)
    backbone = _ConvViT(cfg.MODEL.VIT_TYPE, cfg.DATA.SEARCH.SIZE, cfg.DATA.TEMPLATE.SIZE)
    score_branch = _ScoreDecoder(num_heads=cfg.MODEL.HIDDEN_DIM // 64, hidden_dim=cfg.MODEL.HIDDEN_DIM, pool_size=4)
    return MixFormerConvMAEOnlineScore(backbone, build_box_head(cfg), score_branch, cfg).eval()


BUILDERS = {
    "mixformer_convmae_online": build_mixformer_convmae_online_score,
    "mixformer_vit_online": build_mixformer_vit_online_score,
    "mixformer_vit": build_mixformer_vit,
    "mixformer_vit_rgbt": build_mixformer_vit_rgbt,
    "mixformer_vit_rgbt_shared": build_mixformer_vit_rgbt_shared,
    "mixformer_vit_rgbt_unibackbone": build_mixformer_vit_rgbt_uni,
    "asymmetric_shared": build_asymmetric_shared,
    "asymmetric_shared_ce": build_asymmetric_shared_ce,
    "asymmetric_shared_online": build_asymmetric_shared_online_score,
}
