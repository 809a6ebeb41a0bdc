"""TEST INFRASTRUCTURE - import shims that let the UNMODIFIED reference (/root/reference, read-only,
only present in the build container) be imported and run on CPU, so that the oracle restatement in
oracle/mixformer_oracle.py can be pinned against it and golden vectors can be generated
(oracle/gen_golden.py).  Nothing here is used by the product path, by `-m gpu` tests, by smoke() or
by bench.py: /root/reference does not exist on the GPU box.

Shims:
  * timm.models.vision_transformer.VisionTransformer - a stand-in base class exposing the attributes
    the reference subclasses touch (cls_token, pos_embed, pos_drop, norm, head, init_weights);
    timm.models.layers.{Mlp, DropPath, trunc_normal_} (fc1 -> act -> fc2; identity in eval).
  * easydict.EasyDict, empty mmcv.ops classes, empty matplotlib.pyplot.
  * MultiScaleDeformableAttention.ms_deform_attn_forward -> the reference's own pure-PyTorch
    ms_deform_attn_core_pytorch (lib/models/mixformer_vit_rgbt/deformable_attention/ops/functions/
    ms_deform_attn_func.py:41-61).
  * torch.Tensor.cuda -> identity (the corner head calls .cuda() in __init__, head.py:142-145).
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MMT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models"))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) and not isinstance(x, _EasyDict) else x for x in v)
        super().__setattr__(k, v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__


def _trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer() if act_layer is not None else nn.GELU()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training, "shim DropPath is eval-only"
        return x


class _TimmViT(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, weight_init="", norm_layer=None, act_layer=None, **kw):
        super().__init__()
        norm_layer = norm_layer or nn.LayerNorm
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)

    def init_weights(self, mode=""):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)


_installed = False


def install() -> None:
    """Install the shims into sys.modules and put the reference root on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    timm = mod("timm")
    timm.models = mod("timm.models")
    timm.models.vision_transformer = mod("timm.models.vision_transformer", VisionTransformer=_TimmViT)
    timm.models.layers = mod("timm.models.layers", Mlp=_Mlp, DropPath=_DropPath, trunc_normal_=_trunc_normal_,
                             to_2tuple=lambda x: x if isinstance(x, tuple) else (x, x))
    mod("easydict", EasyDict=_EasyDict)
    mmcv = mod("mmcv")
    mmcv.ops = mod("mmcv.ops", ModulatedDeformConv2d=type("ModulatedDeformConv2d", (nn.Module,), {}),
                   ModulatedDeformConv2dPack=type("ModulatedDeformConv2dPack", (nn.Module,), {}))
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")

    def _msda_forward(value, shapes, level_start, loc, weights, im2col_step):
        from lib.models.mixformer_vit_rgbt.deformable_attention.ops.functions.ms_deform_attn_func import \
            ms_deform_attn_core_pytorch
        return ms_deform_attn_core_pytorch(value, [(int(h), int(w)) for h, w in shapes.tolist()], loc, weights)

    mod("MultiScaleDeformableAttention", ms_deform_attn_forward=_msda_forward, ms_deform_attn_backward=None)

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self   # head.py:142-145 / tracker_utils.py:26-27
        torch.cuda.current_device = lambda: 0

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


# (variant name) -> (config module, builder module, builder fn, experiments sub-directory)
VARIANTS = {
    "mixformer_vit": ("lib.config.mixformer_vit.config", "lib.models.mixformer_vit.mixformer",
                      "build_mixformer_vit", "mixformer_vit"),
    "mixformer_vit_rgbt": ("lib.config.mixformer_vit_rgbt.config", "lib.models.mixformer_vit_rgbt.mixformer",
                           "build_mixformer_vit_rgbt", "mixformer_vit_rgbt"),
    "mixformer_vit_rgbt_shared": ("lib.config.mixformer_vit_rgbt_shared.config",
                                  "lib.models.mixformer_vit_rgbt.mixformer_shared",
                                  "build_mixformer_vit_rgbt_shared", "mixformer_vit_rgbt_shared"),
    "mixformer_vit_rgbt_unibackbone": ("lib.config.mixformer_vit_rgbt_unibackbone.config",
                                       "lib.models.mixformer_vit_rgbt.mixformer_unibackbone",
                                       "build_mixformer_vit_rgbt_uni", "mixformer_vit_rgbt_unibackbone"),
    "asymmetric_shared": ("lib.config.asymmetric_shared.config", "lib.models.mixformer_vit_rgbt.asymmetric_shared",
                          "build_asymmetric_shared", "asymmetric_shared"),
    "asymmetric_shared_online": ("lib.config.asymmetric_shared_online.config",
                                 "lib.models.mixformer_vit_rgbt.asymmetric_shared_online",
                                 "build_asymmetric_shared_online_score", "asymmetric_shared_online"),
    "asymmetric_shared_ce": ("lib.config.asymmetric_shared_ce.config",
                             "lib.models.mixformer_vit_rgbt.asymmetric_shared_ce",
                             "build_asymmetric_shared_ce", "asymmetric_shared_ce"),
    "mixformer_vit_online": ("lib.config.mixformer_vit_online.config", "lib.models.mixformer_vit.mixformer_online",
                             "build_mixformer_vit_online_score", "mixformer_vit_online"),
    "mixformer_convmae_online": ("lib.config.mixformer_convmae_online.config",
                                 "lib.models.mixformer_convmae.mixformer_online",
                                 "build_mixformer_convmae_online_score", "mixformer_convmae_online"),
}


def build_reference_model(variant: str, yaml_name: str):
    """Construct the reference nn.Module for `variant` from its shipped YAML, in eval mode, on CPU."""
    import importlib
    install()
    cfg_mod, model_mod, fn, exp_dir = VARIANTS[variant]
    cm = importlib.import_module(cfg_mod)
    cm.update_config_from_file(os.path.join(REFERENCE_ROOT, "experiments", exp_dir, yaml_name + ".yaml"))
    builder = getattr(importlib.import_module(model_mod), fn)
    model = builder(cm.cfg, train=False)
    model.eval()
    return model, cm.cfg
