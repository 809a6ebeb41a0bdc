"""TEST INFRASTRUCTURE - CPU restatement of the reference's PrRoIPool forward.  NOT part of the product.

Follows external/PreciseRoIPooling/src/prroi_pooling_gpu_impl.cu cell by cell: PrRoIPoolingForward :149-212 walks
the unit cells a bin covers and adds PrRoIPoolingMatCalculation :71-106 (the closed-form integral of the bilinear
surface over the part of the cell inside the bin, reads outside the map = 0, PrRoIPoolingGetData :37-42); the sum is
divided by the bin area; a zero-area bin gives 0.  (The reference op is GPU-only - prroi_pool/functional.py:62-63 -
so this restatement cannot be compared with the reference here; it is pinned by the reference's own known-answer
test, external/PreciseRoIPooling/pytorch/tests/test_prroi_pooling2d.py:21-35, restated in
tests/test_oracle_native_ops.py.)  Pure numpy, float64 accumulation; only for small cases.
"""
from __future__ import annotations

import math

import numpy as np


def _get(data, h, w):
    """PrRoIPoolingGetData: data [C,H,W]; zeros outside."""
    C, H, W = data.shape
    if h < 0 or w < 0 or h >= H or w >= W:
        return np.zeros(C, dtype=np.float64)
    return data[:, h, w].astype(np.float64)


def _mat_calculation(data, s_h, s_w, e_h, e_w, y0, x0, y1, x1):
    """PrRoIPoolingMatCalculation :71-106."""
    def term(a, la, b, lb):
        return (la - 0.5 * la * la - a + 0.5 * a * a) * (lb - 0.5 * lb * lb - b + 0.5 * b * b)

    out = _get(data, s_h, s_w) * term(x0 - s_w, x1 - s_w, y0 - s_h, y1 - s_h)
    out = out + _get(data, s_h, e_w) * term(e_w - x1, e_w - x0, y0 - s_h, y1 - s_h)
    out = out + _get(data, e_h, s_w) * term(x0 - s_w, x1 - s_w, e_h - y1, e_h - y0)
    out = out + _get(data, e_h, e_w) * term(e_w - x1, e_w - x0, e_h - y1, e_h - y0)
    return out


def prroi_pool_forward(features: np.ndarray, rois: np.ndarray, pooled_h: int, pooled_w: int, spatial_scale: float):
    """features [N,C,H,W] float32, rois [R,5] (batch_idx, x0, y0, x1, y1) -> [R,C,pooled_h,pooled_w] float32."""
    N, C, H, W = features.shape
    R = rois.shape[0]
    out = np.zeros((R, C, pooled_h, pooled_w), dtype=np.float32)
    for n in range(R):
        b = int(rois[n, 0])
        rsw, rsh, rew, reh = (np.float32(rois[n, k]) * np.float32(spatial_scale) for k in (1, 2, 3, 4))
        rw, rh = max(float(rew - rsw), 0.0), max(float(reh - rsh), 0.0)
        bh, bw = rh / pooled_h, rw / pooled_w
        data = features[b]
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                ws, hs = float(rsw) + bw * pw, float(rsh) + bh * ph
                we, he = ws + bw, hs + bh
                win = max(0.0, bw * bh)
                if win == 0:
                    continue
                acc = np.zeros(C, dtype=np.float64)
                for w_it in range(math.floor(ws), math.ceil(we)):
                    for h_it in range(math.floor(hs), math.ceil(he)):
                        acc += _mat_calculation(data, h_it, w_it, h_it + 1, w_it + 1, max(hs, float(h_it)),
                                                max(ws, float(w_it)), min(he, h_it + 1.0), min(we, w_it + 1.0))
                out[n, :, ph, pw] = (acc / win).astype(np.float32)
    return out
