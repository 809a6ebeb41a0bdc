"""The reference's two native extension modules under their own pybind names, served by libmmt_b200.so.

The reference's Python reaches its CUDA kernels through two extension objects:
  * `import MultiScaleDeformableAttention as MSDA` -> `MSDA.ms_deform_attn_forward(value, spatial_shapes,
    level_start_index, sampling_locations, attention_weights, im2col_step)`
    (lib/models/mixformer_vit_rgbt/deformable_attention/ops/functions/ms_deform_attn_func.py:23-28; C++ side
    ops/src/vision.cpp, ops/src/cuda/ms_deform_attn_cuda.cu:20-80);
  * the JIT-built `_prroi_pooling` -> `_prroi_pooling.prroi_pooling_forward_cuda(features, rois, pooled_height,
    pooled_width, spatial_scale)` (external/PreciseRoIPooling/pytorch/prroi_pool/functional.py:21-60; C side
    prroi_pooling_gpu.c:22-44).
`MSDA` and `prroi_pooling` below are drop-in objects with exactly those entry points, so that the UNMODIFIED
reference modules (`MSDeformAttn`, `MSDeformAttn_Bimodal`, `PrRoIPool2D`, `ScoreDecoder`) run on the B200 kernels:

    import sys, mmt_b200.native_ops as native
    native.install()          # registers `MultiScaleDeformableAttention` in sys.modules and patches the PrRoIPool loader

Only the forward ops exist (the backward kernels belong to training, which is out of scope): the `*_backward` names
raise NotImplementedError, so these objects serve `torch.no_grad()` callers - every caller under lib/test.
Error behaviour follows the reference: CPU tensors are rejected (ms_deform_attn.h:38 "Not implemented on the CPU",
functional.py:62-63), a non-zero status of the C ABI becomes RuntimeError.
"""
from __future__ import annotations

import sys
import types

import torch

from . import ops


class _MSDA:
    """Stands in for the `MultiScaleDeformableAttention` extension module."""

    @staticmethod
    def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
        if not value.is_cuda:
            raise RuntimeError("Not implemented on the CPU")                       # ms_deform_attn.h:38
        if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
            raise RuntimeError("ms_deform_attn_forward: value [N,S,M,D], sampling_loc [N,Lq,M,L,P,2], "
                               "attn_weight [N,Lq,M,L,P] expected")
        shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]            # host (H, W) pairs
        S = sum(h * w for h, w in shapes)
        if S != value.shape[1] or len(shapes) != sampling_loc.shape[3]:
            raise RuntimeError("ms_deform_attn_forward: spatial_shapes do not match value / sampling_loc")
        # level_start_index is implied by spatial_shapes (cumulative H*W); no `batch % im2col_step == 0` restriction
        # (ms_deform_attn_cuda.cu:50-52).  fp32 and bf16 values are both served; other dtypes are cast up like autocast.
        v = value if value.dtype in (torch.float32, torch.bfloat16) else value.float()
        out = ops.msda(v, shapes, sampling_loc, attn_weight)
        return out if out.dtype == value.dtype else out.to(value.dtype)

    @staticmethod
    def ms_deform_attn_backward(*args, **kwargs):
        raise NotImplementedError("mmt_b200 provides the inference forward only (no deformable-attention backward)")


class _PrRoIPooling:
    """Stands in for the JIT-built `_prroi_pooling` module of PreciseRoIPooling."""

    @staticmethod
    def prroi_pooling_forward_cuda(features, rois, pooled_height, pooled_width, spatial_scale):
        if not features.is_cuda:
            raise NotImplementedError("Precise RoI Pooling only supports GPU (cuda) implememtations.")   # functional.py:62-63
        if features.dtype != torch.float32 or rois.dtype != torch.float32:
            raise AssertionError("Precise RoI Pooling only takes float input")                           # functional.py:48-49
        return ops.prroi_pool(features, rois, int(pooled_height), int(pooled_width), float(spatial_scale))

    @staticmethod
    def prroi_pooling_backward_cuda(*args, **kwargs):
        raise NotImplementedError("mmt_b200 provides the inference forward only (no PrRoIPool backward)")

    prroi_pooling_coor_backward_cuda = prroi_pooling_backward_cuda


MSDA = _MSDA()
prroi_pooling = _PrRoIPooling()


def install() -> None:
    """Make the reference pick these objects up without editing it: `import MultiScaleDeformableAttention` resolves to
    MSDA, and `prroi_pool.functional._import_prroi_pooling()` (when that module is importable) returns prroi_pooling."""
    mod = types.ModuleType("MultiScaleDeformableAttention")
    mod.ms_deform_attn_forward = MSDA.ms_deform_attn_forward
    mod.ms_deform_attn_backward = MSDA.ms_deform_attn_backward
    sys.modules["MultiScaleDeformableAttention"] = mod
    try:        # the reference's PrRoIPool package, when its tree is importable (lib/models/mixformer_cvt/score_decoder.py:9)
        import importlib
        importlib.import_module("external.PreciseRoIPooling.pytorch.prroi_pool.functional")
    except ImportError:
        pass
    for name, m in list(sys.modules.items()):
        if name.endswith("prroi_pool.functional") and hasattr(m, "_import_prroi_pooling"):
            m._prroi_pooling = prroi_pooling
            m._import_prroi_pooling = lambda: prroi_pooling
        if name.endswith("ms_deform_attn_func") and hasattr(m, "MSDA"):      # already imported: rebind its module global
            m.MSDA = mod
