"""TEST / BASELINE INFRASTRUCTURE - import shims that let the UNMODIFIED reference be imported and run, so that
  * the oracle restatement in oracle/mixformer_oracle.py can be pinned against it and golden vectors can be
    generated (oracle/gen_golden.py; build container, /root/reference), and
  * bench.py can time the reference itself on the GPU box, on the host cores (`--impl reference`) and in eager mode on
    the B200 (`gpu_eager_baseline`), from the unmodified copy oracle/ship_ref.py puts under baseline/_ref.
Nothing here is used by the product path, by `-m gpu` parity tests or by smoke().

Shims for packages the image lacks - the reference's own files are imported as they are:
  * timm.models.vision_transformer.VisionTransformer - a stand-in base class exposing the attributes
    the reference subclasses touch (cls_token, pos_embed, pos_drop, norm, head, init_weights);
    timm.models.layers.{Mlp, DropPath, trunc_normal_} (fc1 -> act -> fc2; identity in eval).
  * easydict.EasyDict, empty mmcv.ops classes, empty matplotlib.pyplot.
  * MultiScaleDeformableAttention.ms_deform_attn_forward (the reference's pybind extension, which does not build
    against torch 2.11): CUDA tensors -> the reference's OWN kernel compiled by oracle/build_ref.py
    (oracle/_ref/libmsda_ref.so); CPU tensors (or library absent) -> the reference's own pure-PyTorch
    ms_deform_attn_core_pytorch (deformable_attention/ops/functions/ms_deform_attn_func.py:41-61).
  * torch.Tensor.cuda -> identity for CPU runs (the corner head calls .cuda() in __init__, head.py:142-145).
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """/root/reference in the build container; on the GPU box the unmodified copy that oracle/ship_ref.py put under
    baseline/_ref (git-ignored, travels with the gpurun snapshot)."""
    env = os.environ.get("MMT_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_REPO, "baseline", "_ref")]:
        if os.path.isdir(os.path.join(cand, "lib", "models")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models"))


def reference_kind() -> str:
    return "shipped copy (baseline/_ref)" if REFERENCE_ROOT.startswith(_REPO) else REFERENCE_ROOT


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


_MSDA_LIB = None


def _msda_ref_lib():
    """The reference's OWN deformable-attention forward kernel, compiled by oracle/build_ref.py (None if not built)."""
    global _MSDA_LIB
    if _MSDA_LIB is None:
        import ctypes
        path = os.path.join(_REPO, "oracle", "_ref", "libmsda_ref.so")
        _MSDA_LIB = False
        if os.path.exists(path):
            try:
                _MSDA_LIB = ctypes.CDLL(path)
                _MSDA_LIB.msda_ref_forward.restype = ctypes.c_int
            except OSError:
                _MSDA_LIB = False
    return _MSDA_LIB or None


MSDA_BACKEND = {"last": None}      # which implementation served the last MSDA call (reported by bench.py)


def install(cpu_model: bool | None = None) -> None:
    """Install the shims into sys.modules and put the reference root on sys.path.
    cpu_model: the reference model will run on the CPU (neutralise the `.cuda()` calls of the head's constructor,
    head.py:142-145); default = True exactly when no GPU is visible.  Pass True for a CPU run on a GPU box."""
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
        lib = _msda_ref_lib() if value.is_cuda else None
        if lib is not None:
            # CUDA tensors: the reference's own kernel (ms_deform_im2col_cuda.cuh) behind oracle/msda_ref_wrapper.cu,
            # called the way ms_deform_attn_cuda.cu:20-80 calls it.  The kernel is fp32 (the reference dispatches
            # float/double only): under autocast the operands are cast up, what custom_fwd(cast_inputs=float32) does.
            import ctypes
            odt = value.dtype
            v, l, w = value.float().contiguous(), loc.float().contiguous(), weights.float().contiguous()
            ss, ls = shapes.to(torch.int64).contiguous(), level_start.to(torch.int64).contiguous()
            N, S, M, D = v.shape
            Lq, L, P = l.shape[1], l.shape[3], l.shape[4]
            out = torch.zeros((N, Lq, M * D), device=v.device, dtype=torch.float32)
            step = min(int(im2col_step), N)
            while N % step:
                step -= 1
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            st = lib.msda_ref_forward(p(v), p(ss), p(ls), p(l), p(w), p(out), N, S, M, D, L, Lq, P, step,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            if st != 0:
                raise RuntimeError(f"msda_ref_forward failed: {st}")
            MSDA_BACKEND["last"] = "reference CUDA kernel (oracle/_ref/libmsda_ref.so)"
            return out.to(odt)
        from lib.models.mixformer_vit_rgbt.deformable_attention.ops.functions.ms_deform_attn_func import \
            ms_deform_attn_core_pytorch
        MSDA_BACKEND["last"] = "reference pure-PyTorch core (ms_deform_attn_core_pytorch)"
        return ms_deform_attn_core_pytorch(value, [(int(h), int(w)) for h, w in shapes.tolist()], loc, weights)

    mod("MultiScaleDeformableAttention", ms_deform_attn_forward=_msda_forward, ms_deform_attn_backward=None)

    if cpu_model is None:
        cpu_model = not torch.cuda.is_available()
    if cpu_model:
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


def build_reference_model(variant: str, yaml_name: str, cpu_model: bool | None = None, overrides: dict | None = None):
    """Construct the reference nn.Module for `variant` from its shipped YAML, in eval mode, on CPU (call .cuda() on the
    result for the eager-GPU baseline; pass cpu_model=True when it will RUN on the CPU of a GPU box).
    overrides: {"MODEL.HEAD_TYPE": "CORNER", ...} applied to the reference's cfg after the YAML (configurations the
    reference supports but ships no experiment file for)."""
    import importlib
    install(cpu_model)
    cfg_mod, model_mod, fn, exp_dir = VARIANTS[variant]
    cm = importlib.import_module(cfg_mod)
    cm.update_config_from_file(os.path.join(REFERENCE_ROOT, "experiments", exp_dir, yaml_name + ".yaml"))
    for k, v in (overrides or {}).items():
        node = cm.cfg
        parts = k.split(".")
        for p_ in parts[:-1]:
            node = getattr(node, p_)
        setattr(node, parts[-1], v)
    builder = getattr(importlib.import_module(model_mod), fn)
    model = builder(cm.cfg, train=False)
    model.eval()
    return model, cm.cfg
