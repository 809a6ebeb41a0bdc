"""Fixed positional tables of the forward (built once on the host at model-construction / weight-pack time;
they are constants of the model, like the weights).

  * sincos_pos_embed_2d: the MixViT backbone's frozen 2-D sin-cos table - behaviour of
    get_2d_sincos_pos_embed (lib/models/mixformer_vit/pos_utils.py:20-67): fp32 numpy, the x (column) index is
    encoded in the first half of the channels and the y (row) index in the second half; each half is
    [sin(pos*w_k) | cos(pos*w_k)], w_k = 10000^(-k/(D/4)).
  * fusion_pos_table: PositionEmbeddingSine(num_pos_feats=d/2, normalize=True) of an all-valid H x W mask
    (deformable_attention/position_encoding.py:24-56) flattened to tokens, plus level_embed[l], for the two
    "levels" (= modalities) of the fusion encoder (deformable_encoder_lnspecific.py:86-100).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    k = np.arange(dim // 2, dtype=np.float32)
    k /= dim / 2.0
    freq = 1.0 / 10000 ** k
    ang = np.einsum("m,d->md", pos.reshape(-1), freq)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_pos_embed_2d(embed_dim: int, grid_size: int) -> torch.Tensor:
    """[grid_size*grid_size, embed_dim] fp32, tokens in row-major (y outer, x inner) order."""
    ys, xs = np.meshgrid(np.arange(grid_size, dtype=np.float32), np.arange(grid_size, dtype=np.float32),
                         indexing="ij")
    emb = np.concatenate([_sincos_1d(embed_dim // 2, xs), _sincos_1d(embed_dim // 2, ys)], axis=1)
    return torch.from_numpy(emb).float()


def fusion_pos_table(h: int, w: int, d_model: int, level_embed: torch.Tensor) -> torch.Tensor:
    """[2*h*w, d_model] fp32: rows [0, hw) for level 0 (RGB), [hw, 2hw) for level 1 (TIR)."""
    npf = d_model // 2
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32)
    x = torch.arange(1, w + 1, dtype=torch.float32)
    y = (y - 0.5) / (float(h) + eps) * scale
    x = (x - 0.5) / (float(w) + eps) * scale
    i = torch.arange(npf, dtype=torch.float32)
    dim_t = 10000.0 ** (2 * torch.div(i, 2, rounding_mode="floor") / npf)

    def enc(v):          # [n] -> [n, npf], interleaved sin/cos
        a = v[:, None] / dim_t
        return torch.stack((a[:, 0::2].sin(), a[:, 1::2].cos()), dim=2).flatten(1)

    py, px = enc(y), enc(x)                                    # [h, npf], [w, npf]
    pos = torch.cat([py[:, None, :].expand(h, w, npf), px[None, :, :].expand(h, w, npf)], dim=2).reshape(h * w, d_model)
    lvl = level_embed.detach().float().cpu()
    return torch.cat([pos + lvl[0][None], pos + lvl[1][None]], dim=0).contiguous()
