"""Thin torch-tensor wrappers over the C ABI (include/mmt_b200.h).

torch is used here only as the owner of device memory and of the current CUDA stream; every
function forwards raw pointers to libmmt_b200.so and raises RuntimeError on a non-zero status.
Nothing in this module computes on the CPU or with torch ops.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_void_p

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2

# Launch accounting (bench.py): LAUNCHES counts kernels launched through this module; PROFILER, when set to a
# LaunchProfiler, brackets every tensor-core GEMM launch with CUDA events on the launching stream.
LAUNCHES = 0
PROFILER = None


class LaunchProfiler:
    """Per-launch CUDA-event timing by kernel class, with the algorithmic work (FLOPs) of each GEMM launch.
    all_ops=False brackets only the tensor-core GEMM (the roofline kernel of bench.py); True brackets every op."""

    def __init__(self, all_ops=False):
        self.all_ops = all_ops
        self.spans = {}          # class name -> [(start_event, end_event, flops)]

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        return e

    def end(self, start, flops, name="gemm_bf16"):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        self.spans.setdefault(name, []).append((start, e, flops))

    def summary(self, name="gemm_bf16"):
        """(launches, total_ms, total_flops) of one class (a prefix selects several, e.g. "gemm_bf16" = pair + single +
        conv3x3 instantiations); call after a device synchronise."""
        sp = [x for k, v in self.spans.items() if k.startswith(name) for x in v]
        ms = sum(a.elapsed_time(b) for a, b, _ in sp)
        return len(sp), ms, float(sum(f for _, _, f in sp))

    def table(self):
        return {k: dict(zip(("launches", "ms", "flops"), self.summary(k))) for k in self.spans}


_PAIR_MIN = None


def config_pdl(enable: bool) -> bool:
    """Programmatic dependent launch of the GEMM / attention / LayerNorm-statistics kernels on or off (mmt_config_pdl);
    returns the previous setting.  Results do not depend on it."""
    f = _lib.fn("mmt_config_pdl")
    return bool(f(c_int(1 if enable else 0)))


def config_cluster4(enable: bool) -> bool:
    """Clusters of four CTAs (two pairs, multicast weight tile) for the big backbone GEMMs on or off (mmt_config_cluster4);
    returns the previous setting.  Results do not depend on it."""
    f = _lib.fn("mmt_config_cluster4")
    return bool(f(c_int(1 if enable else 0)))


def config_small_gemm_sms(sms: int) -> int:
    """SM budget of the small-GEMM one-wave tile rule (mmt_config_small_gemm_sms; 0 = whole GPU); returns the previous one."""
    f = _lib.fn("mmt_config_small_gemm_sms")
    return int(f(c_int(int(sms))))


def _pair_min_tiles():
    global _PAIR_MIN
    if _PAIR_MIN is None:
        import os
        if os.environ.get("MMT_B200_DEV_LIB") == "1" and os.environ.get("MMT_GEMM_PAIR", "1") == "0":
            _PAIR_MIN = 1 << 60          # developer build with the CTA-pair kernel switched off (A/B)
        else:
            _PAIR_MIN = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count // 2
    return _PAIR_MIN


def _begin():
    p = PROFILER
    return p.begin() if (p is not None and p.all_ops) else None


def _count(n=1, name=None, ev=None):
    global LAUNCHES
    LAUNCHES += n
    if ev is not None:
        PROFILER.end(ev, 0.0, name)


def _ptr(t):
    return c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            # same convention as the reference's native ops (prroi_pool/functional.py:62-63)
            raise NotImplementedError("mmt_b200 ops are CUDA-only (no CPU fallback)")


_gemm_bf16 = _lib.fn("mmt_gemm_bf16_ex")
_gemm_f32 = _lib.fn("mmt_gemm_f32")
_rowstats_cast = _lib.fn("mmt_rowstats_cast")


def _check_ln_sums(t, M):
    """Statistics tensors of the folded LayerNorm are slot-major [slots, rows, 2] fp32 (row slices allowed)."""
    assert t.dtype == torch.float32 and t.dim() == 3 and t.shape[1] == M and t.shape[2] == 2
    assert t.stride(2) == 1 and t.stride(1) == 2 and t.stride(0) % 2 == 0 and t.stride(0) >= 2 * M
    return t.shape[0], t.stride(0) // 2


def rowstats_cast(x, xb, stats):
    """xb = bf16(x), stats[0] = (sum, sum of squares) of every fp32 row, other slots zero (mmt_rowstats_cast): the
    stand-alone producer of the folded LayerNorm.  stats: [slots, rows, 2]."""
    _need_cuda(x, xb, stats)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    assert xb.dtype == torch.bfloat16 and xb.shape == x.shape and xb.stride(1) == 1
    slots, stride = _check_ln_sums(stats, x.shape[0])
    _ev = _begin()
    _lib.check(_rowstats_cast(_ptr(x), c_int(x.shape[0]), c_int(x.shape[1]), _ptr(xb), c_int(xb.stride(0)), _ptr(stats),
                              c_int(slots), c_int(stride), _stream()), "mmt_rowstats_cast")
    _count(1, "rowstats_cast", _ev)


def gemm(a, w, bias=None, act=ACT_NONE, resid=None, rowadd=None, out=None, out_dtype=None, max_ctas=0,
         ln_stats=None, ln_eps=0.0, colsum=None, xb_out=None, stats_out=None):
    """out = act(a @ w.T + bias) + rowadd[row % period] + resid.

    a: [M, K] (may be a row-strided view), w: [N, K]; both bf16 (tcgen05 kernel) or both fp32 (parity kernel).
    bias [N] / rowadd [period, N] / resid [M, N] are fp32.  `out` may be a column-slice view.
    Folded LayerNorm (bf16 only, mmt_gemm_bf16_ex): ln_stats [slots, M, 2] + colsum [N] = `a` holds raw residual rows,
    w / bias carry gamma / beta, the epilogue finishes the normalisation; xb_out [M, N] bf16 + stats_out [N/128, M, 2] =
    also leave the bf16 copy and the partial sums of the fp32 output rows for the next consumer.
    """
    _need_cuda(a, w, bias, resid, rowadd, out, ln_stats, colsum, xb_out, stats_out)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    is_bf16 = a.dtype == torch.bfloat16
    assert w.dtype == a.dtype
    if out is None:
        if out_dtype is None:
            out_dtype = torch.float32 if (resid is not None or not is_bf16) else torch.bfloat16
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    for t in (bias, resid, rowadd):
        assert t is None or t.dtype == torch.float32
    assert bias is None or (bias.numel() == N and bias.is_contiguous())
    assert rowadd is None or (rowadd.shape[1] == N and rowadd.is_contiguous())
    assert resid is None or (resid.shape == (M, N) and resid.stride(1) == 1)
    ldr = resid.stride(0) if resid is not None else 0
    period = rowadd.shape[0] if rowadd is not None else 0
    _count()
    if is_bf16:
        prof = PROFILER
        ev = prof.begin() if prof is not None else None
        slots = ln_stride = out_stride = 0
        if ln_stats is not None:
            assert colsum is not None and colsum.dtype == torch.float32 and colsum.numel() == N
            slots, ln_stride = _check_ln_sums(ln_stats, M)
        if xb_out is not None:
            assert stats_out is not None and xb_out.dtype == torch.bfloat16 and xb_out.shape == (M, N)
            n_out, out_stride = _check_ln_sums(stats_out, M)
            assert n_out == N // 128
        st = _gemm_bf16(_ptr(a), c_int(a.stride(0)), _ptr(w), c_int(w.stride(0)), c_int(M), c_int(N), c_int(K),
                        _ptr(bias), c_int(act), _ptr(resid), c_int(ldr), _ptr(rowadd), c_int(period), _ptr(out),
                        c_int(out.stride(0)), c_int(1 if out.dtype == torch.float32 else 0), c_int(max_ctas),
                        _ptr(ln_stats), c_int(slots), c_int(ln_stride), c_float(ln_eps), _ptr(colsum), _ptr(xb_out),
                        c_int(xb_out.stride(0) if xb_out is not None else 0), _ptr(stats_out), c_int(out_stride), _stream())
        _lib.check(st, "mmt_gemm_bf16_ex")
        if ev is not None:
            # same dispatch rule as gemm_tc.cu::dispatch_gemm: the CTA-pair kernel takes the big N % 256 == 0 GEMMs
            pair = max_ctas <= 0 and N % 256 == 0 and ((M + 255) // 256) * (N // 256) >= _pair_min_tiles()
            prof.end(ev, 2.0 * M * N * K, "gemm_bf16_pair" if pair else "gemm_bf16_single")
    else:
        assert a.dtype == torch.float32 and out.dtype == torch.float32
        assert ln_stats is None and xb_out is None, "the folded LayerNorm exists in the bf16 tcgen05 path only"
        st = _gemm_f32(_ptr(a), c_int(a.stride(0)), _ptr(w), c_int(w.stride(0)), c_int(M), c_int(N), c_int(K),
                       _ptr(bias), c_int(act), _ptr(resid), c_int(ldr), _ptr(rowadd), c_int(period), _ptr(out),
                       c_int(out.stride(0)), _stream())
        _lib.check(st, "mmt_gemm_f32")
    return out


# --------------------------------------------------------------------------------------------- row kernels
_patchify = _lib.fn("mmt_patchify")
_layernorm = _lib.fn("mmt_layernorm")
_groupnorm = _lib.fn("mmt_groupnorm")
_copy_rows = _lib.fn("mmt_copy_rows")
_fusion_prep = _lib.fn("mmt_fusion_prep")
_im2col3x3 = _lib.fn("mmt_im2col3x3")
_corner_decode = _lib.fn("mmt_corner_decode")
_msda = _lib.fn("mmt_msda_fwd")
_msda_bimodal = _lib.fn("mmt_msda_bimodal_fwd")
_mixattn = _lib.fn("mmt_mixattn_fwd")
_ce_scores = _lib.fn("mmt_ce_scores")
_ce_scores_split = _lib.fn("mmt_ce_scores_split")
_ce_topk = _lib.fn("mmt_ce_topk")
_ce_gather = _lib.fn("mmt_ce_gather_tokens")
_ce_recover = _lib.fn("mmt_ce_recover")
_prroi = _lib.fn("mmt_prroi_fwd")


def _is_bf16(t):
    if t.dtype == torch.bfloat16:
        return 1
    assert t.dtype == torch.float32, t.dtype
    return 0


def patchify(img, out, tok_off, tok_per_seq, patch=16):
    """img fp32 NCHW [B,Cin,H,W] -> rows of `out` [.., Cin*P*P] at b*tok_per_seq + tok_off + patch index."""
    _need_cuda(img, out)
    assert img.dtype == torch.float32 and img.is_contiguous() and out.is_contiguous()
    B, Cin, H, W = img.shape
    assert out.shape[1] == Cin * patch * patch
    _ev = _begin()
    _lib.check(_patchify(_ptr(img), _ptr(out), c_int(B), c_int(Cin), c_int(H), c_int(W), c_int(patch),
                         c_int(tok_off), c_int(tok_per_seq), c_int(_is_bf16(out)), _stream()), "mmt_patchify")
    _count(1, "patchify", _ev)
    return out


def layernorm(x, g0, b0, g1=None, b1=None, period=0, eps=1e-6, out_f32=None, out_bf16=None):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    rows, C = x.shape
    for o in (out_f32, out_bf16):
        assert o is None or (o.shape == x.shape and o.is_contiguous())
    assert out_f32 is None or out_f32.dtype == torch.float32
    assert out_bf16 is None or out_bf16.dtype == torch.bfloat16
    _ev = _begin()
    _lib.check(_layernorm(_ptr(x), c_int(rows), c_int(C), c_float(eps), _ptr(g0), _ptr(b0), _ptr(g1), _ptr(b1),
                          c_int(period), _ptr(out_f32), _ptr(out_bf16), _stream()), "mmt_layernorm")
    _count(1, "layernorm", _ev)


def groupnorm(x, B, HW, groups, gamma, beta, eps=1e-5, out_f32=None, out_bf16=None, out_seq_rows=0, out_row_off=0):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] == B * HW
    assert out_f32 is None or out_f32.dtype == torch.float32
    assert out_bf16 is None or out_bf16.dtype == torch.bfloat16
    C = x.shape[1]
    _ev = _begin()
    _lib.check(_groupnorm(_ptr(x), c_int(B), c_int(HW), c_int(C), c_int(groups), c_float(eps), _ptr(gamma),
                          _ptr(beta), _ptr(out_f32), _ptr(out_bf16), c_int(out_seq_rows), c_int(out_row_off),
                          _stream()), "mmt_groupnorm")
    _count(1, "groupnorm", _ev)


def copy_rows(src, seq_stride, row_off, rows_per_seq, nseq, dst):
    _need_cuda(src, dst)
    assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    C = src.shape[-1]
    _ev = _begin()
    _lib.check(_copy_rows(_ptr(src), c_int(seq_stride), c_int(row_off), c_int(rows_per_seq), c_int(nseq), c_int(C),
                          _ptr(dst), c_int(_is_bf16(dst)), _stream()), "mmt_copy_rows")
    _count(1, "copy_rows", _ev)
    return dst


def fusion_prep(src, pos, B, L, out_val=None, out_q=None):
    _need_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    C = src.shape[-1]
    ref = out_val if out_val is not None else out_q
    _ev = _begin()
    _lib.check(_fusion_prep(_ptr(src), _ptr(pos), c_int(B), c_int(L), c_int(C), _ptr(out_val), _ptr(out_q),
                            c_int(_is_bf16(ref)), _stream()), "mmt_fusion_prep")
    _count(1, "fusion_prep", _ev)


def im2col3x3(src1, s1, B, H, W, C, out, src2=None, s2=1):
    """src maps are [B*(H/s)*(W/s), ld] row views (channel slices allowed); out [B*H*W, 9*C] contiguous."""
    _need_cuda(src1, out)
    assert src1.stride(1) == 1 and out.is_contiguous() and out.shape == (B * H * W, 9 * C)
    assert src2 is None or (src2.stride(1) == 1 and src2.dtype == src1.dtype)
    _ev = _begin()
    _lib.check(_im2col3x3(_ptr(src1), c_int(src1.stride(0)), c_int(s1), _ptr(src2),
                          c_int(src2.stride(0) if src2 is not None else 0), c_int(s2), c_int(B), c_int(H), c_int(W),
                          c_int(C), _ptr(out), c_int(_is_bf16(out)), _stream()), "mmt_im2col3x3")
    _count(1, "im2col3x3", _ev)
    return out


_conv3x3 = _lib.fn("mmt_conv3x3_bf16")
_upsample_add = _lib.fn("mmt_upsample_add")


def conv3x3(src, B, H, W, C, w, bias, act, out):
    """Implicit-GEMM Conv2d(k=3, pad=1)+bias+act: src bf16 rows [B*H*W, ld] (channel-slice views allowed, C channels
    used), w bf16 [N, 9*C] packed (ky, kx, c), out [B*H*W, N] bf16/fp32 (column-slice views allowed)."""
    _need_cuda(src, w, out)
    assert src.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and src.stride(1) == 1 and out.stride(1) == 1
    assert src.shape[0] == B * H * W and w.shape[1] == 9 * C and out.shape == (B * H * W, w.shape[0])
    N = w.shape[0]
    prof = PROFILER
    ev = prof.begin() if prof is not None else None
    _lib.check(_conv3x3(_ptr(src), c_int(src.stride(0)), c_int(B), c_int(H), c_int(W), c_int(C), _ptr(w),
                        c_int(w.stride(0)), c_int(N), _ptr(bias), c_int(act), _ptr(out), c_int(out.stride(0)),
                        c_int(1 if out.dtype == torch.float32 else 0), _stream()), "mmt_conv3x3_bf16")
    _count()
    if ev is not None:
        prof.end(ev, 2.0 * B * H * W * N * 9 * C, "gemm_bf16_conv3x3")
    return out


def upsample_add(src1, s1, B, H, W, C, out, src2=None, s2=1):
    _need_cuda(src1, out)
    assert src1.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and out.is_contiguous()
    assert out.shape == (B * H * W, C) and src1.stride(1) == 1 and (src2 is None or src2.stride(1) == 1)
    _ev = _begin()
    _lib.check(_upsample_add(_ptr(src1), c_int(src1.stride(0)), c_int(s1), _ptr(src2),
                             c_int(src2.stride(0) if src2 is not None else 0), c_int(s2), c_int(B), c_int(H), c_int(W),
                             c_int(C), _ptr(out), _stream()), "mmt_upsample_add")
    _count(1, "upsample_add", _ev)
    return out


def corner_decode(x4, w5, b5, a3, a4, B, S, stride_px, img_sz, xyxy, cxcywh, score_maps=None):
    """x4/a3/a4: (tl, br) pairs of row views; w5: (tl, br) fp32 [C4]; b5: (tl, br) floats."""
    _need_cuda(x4[0], xyxy, cxcywh)
    C4 = w5[0].numel()
    if a3 is None:          # plain corner head: no side maps
        a3 = a4 = (None, None)
        lda3 = lda4 = 0
    else:
        assert a3[0].stride(0) == a3[1].stride(0) and a4[0].stride(0) == a4[1].stride(0)
        lda3, lda4 = a3[0].stride(0), a4[0].stride(0)
    assert x4[0].stride(0) == x4[1].stride(0)
    _ev = _begin()
    _lib.check(_corner_decode(_ptr(x4[0]), _ptr(x4[1]), c_int(x4[0].stride(0)), c_int(C4), _ptr(w5[0]), _ptr(w5[1]),
                              c_float(b5[0]), c_float(b5[1]), _ptr(a3[0]), _ptr(a3[1]), c_int(lda3),
                              _ptr(a4[0]), _ptr(a4[1]), c_int(lda4), c_int(B), c_int(S),
                              c_float(stride_px), c_float(img_sz), _ptr(score_maps), _ptr(xyxy), _ptr(cxcywh),
                              c_int(_is_bf16(x4[0])), _stream()), "mmt_corner_decode")
    _count(2, "corner_decode", _ev)


def msda(value, level_hw, sampling_loc, attn_weight, out=None):
    """Reference-layout MSDA forward: value [N,S,M,D], loc [N,Lq,M,L,P,2], attn [N,Lq,M,L,P] -> [N,Lq,M*D]."""
    _need_cuda(value, sampling_loc, attn_weight)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    value = value.contiguous()
    loc = sampling_loc.float().contiguous()
    aw = attn_weight.float().contiguous()
    if out is None:
        out = torch.empty((N, Lq, M * D), device=value.device, dtype=value.dtype)
    hw = (c_int * (2 * L))(*[int(v) for pair in level_hw for v in pair])
    _ev = _begin()
    _lib.check(_msda(_ptr(value), hw, _ptr(loc), _ptr(aw), _ptr(out), c_int(N), c_int(S), c_int(M), c_int(D),
                     c_int(L), c_int(Lq), c_int(P), c_int(_is_bf16(value)), _stream()), "mmt_msda_fwd")
    _count(1, "msda", _ev)
    return out


def msda_bimodal(value, offw, out, B, H, W, M=8, D=64, P=4):
    _need_cuda(value, offw, out)
    assert offw.dtype == torch.float32 and offw.stride(1) == 1 and value.is_contiguous() and out.is_contiguous()
    _ev = _begin()
    _lib.check(_msda_bimodal(_ptr(value), _ptr(offw), c_int(offw.stride(0)), _ptr(out), c_int(B), c_int(H), c_int(W),
                             c_int(M), c_int(D), c_int(P), c_int(_is_bf16(value)), _stream()), "mmt_msda_bimodal_fwd")
    _count(1, "msda_bimodal", _ev)
    return out


def order_tiles(recs):
    """Attention tile records (see mmt_mixattn_fwd) as an int32 array ordered by key count, heavy first (stable): the
    persistent kernel hands items (tile, head) to its CTAs with a fixed stride, which balances only if cost varies
    slowly along the table.  The order of the records is free - each carries its own query and output rows."""
    import numpy as np
    a = np.asarray(recs, dtype=np.int32).reshape(-1, 16)
    # key count only, stable: the query tiles of one sequence keep their neighbourhood, so the CTAs that share a
    # sequence's K / V slices run at the same time and the second reader hits L2 (sorting the 68-row tail tiles away
    # from their 128-row siblings cost +70 % DRAM reads, profiles/r1d_launches_final.md)
    keys = a[:, 7:10].sum(1).astype(np.int64)
    return np.ascontiguousarray(a[np.argsort(-keys, kind="stable")])


def mixattn(qkv0, qkv1, C, heads, tiles, max_keys, out, scale):
    _need_cuda(qkv0, tiles, out)
    assert tiles.dtype == torch.int32 and tiles.is_contiguous() and tiles.shape[1] == 16
    assert qkv0.stride(1) == 1 and out.stride(1) == 1
    rows1 = qkv1.shape[0] if qkv1 is not None else 0
    assert qkv1 is None or qkv1.stride(0) == qkv0.stride(0)
    _ev = _begin()
    _lib.check(_mixattn(_ptr(qkv0), c_int(qkv0.shape[0]), _ptr(qkv1), c_int(rows1), c_int(qkv0.stride(0)), c_int(C),
                        c_int(heads), _ptr(tiles),
                        c_int(tiles.shape[0]), c_int(max_keys), _ptr(out), c_int(out.stride(0)), c_float(scale),
                        c_int(_is_bf16(qkv0)), _stream()), "mmt_mixattn_fwd")
    _count(1, "mixattn", _ev)
    return out


def ce_scores(qkv, C, heads, B, n_tok, Lt, Ls, scale, partial_ws, scores, q_src=None, q_seq_rows=0, k_row_off=None):
    """q_src: buffer holding the template rows (queries) when they do not live in `qkv` (cached-template path):
    q_seq_rows rows per sequence-modality block; k_row_off: first search row inside a block of `qkv` (default Lt)."""
    _need_cuda(qkv, partial_ws, scores, q_src)
    _ev = _begin()
    if q_src is None:
        _lib.check(_ce_scores(_ptr(qkv), c_int(qkv.stride(0)), c_int(C), c_int(heads), c_int(B), c_int(n_tok), c_int(Lt),
                              c_int(Ls), c_float(scale), _ptr(partial_ws), _ptr(scores), c_int(_is_bf16(qkv)), _stream()),
                   "mmt_ce_scores")
    else:
        assert q_src.dtype == qkv.dtype and q_src.stride(0) == qkv.stride(0) and q_src.stride(1) == 1
        _lib.check(_ce_scores_split(_ptr(q_src), c_int(q_seq_rows), _ptr(qkv), c_int(n_tok),
                                    c_int(Lt if k_row_off is None else k_row_off), c_int(qkv.stride(0)), c_int(C),
                                    c_int(heads), c_int(B), c_int(Lt), c_int(Ls), c_float(scale), _ptr(partial_ws),
                                    _ptr(scores), c_int(_is_bf16(qkv)), _stream()), "mmt_ce_scores_split")
    _count(2, "ce_scores", _ev)
    return scores


def ce_topk(scores, B, Ls, keep, gidx_in, gidx_keep, gidx_removed, order):
    _need_cuda(scores, gidx_in, gidx_keep, gidx_removed, order)
    assert order.dtype == torch.int32
    _ev = _begin()
    _lib.check(_ce_topk(_ptr(scores), c_int(B), c_int(Ls), c_int(keep), _ptr(gidx_in), _ptr(gidx_keep),
                        _ptr(gidx_removed), _ptr(order), _stream()), "mmt_ce_topk")
    _count(1, "ce_topk", _ev)


def ce_gather_tokens(x, nseq, n_tok, Lt, order, Ls, keep, x_out):
    _need_cuda(x, order, x_out)
    _ev = _begin()
    _lib.check(_ce_gather(_ptr(x), c_int(nseq), c_int(n_tok), c_int(Lt), _ptr(order), c_int(Ls), c_int(keep),
                          _ptr(x_out), c_int(x.shape[-1]), _stream()), "mmt_ce_gather_tokens")
    _count(1, "ce_gather_tokens", _ev)


def ce_recover(x, nseq, n_tok, Lt, gidx, Lk, Ls0, out):
    _need_cuda(x, gidx, out)
    _ev = _begin()
    _lib.check(_ce_recover(_ptr(x), c_int(nseq), c_int(n_tok), c_int(Lt), _ptr(gidx), c_int(Lk), c_int(Ls0),
                           _ptr(out), c_int(x.shape[-1]), c_int(_is_bf16(out)), _stream()), "mmt_ce_recover")
    _count(1, "ce_recover", _ev)


_layernorm_act = _lib.fn("mmt_layernorm_act")
_patchify2x2 = _lib.fn("mmt_patchify2x2")
_dwconv5x5 = _lib.fn("mmt_dwconv5x5")


def layernorm_act(x, g, b, eps, gelu, out_f32=None, out_bf16=None, seg_rows=0, out_seq_rows=0, out_row_off=0):
    """Channel LayerNorm (+ exact GELU) of fp32 rows; optional re-mapping of the output rows (see mmt_b200.h)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    assert out_f32 is None or (out_f32.dtype == torch.float32 and out_f32.is_contiguous())
    assert out_bf16 is None or (out_bf16.dtype == torch.bfloat16 and out_bf16.is_contiguous())
    _ev = _begin()
    _lib.check(_layernorm_act(_ptr(x), c_int(x.shape[0]), c_int(x.shape[1]), c_float(eps), _ptr(g), _ptr(b),
                              c_int(1 if gelu else 0), _ptr(out_f32), _ptr(out_bf16), c_int(seg_rows),
                              c_int(out_seq_rows), c_int(out_row_off), _stream()), "mmt_layernorm_act")
    _count(1, "layernorm_act", _ev)


def patchify2x2(x, B, H, W, out):
    _need_cuda(x, out)
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    C = x.shape[1]
    assert x.shape[0] == B * H * W and out.shape == (B * (H // 2) * (W // 2), 4 * C)
    _ev = _begin()
    _lib.check(_patchify2x2(_ptr(x), c_int(B), c_int(H), c_int(W), c_int(C), _ptr(out), c_int(_is_bf16(out)),
                            _stream()), "mmt_patchify2x2")
    _count(1, "patchify2x2", _ev)
    return out


def dwconv5x5(x, w, bias, B, H, W, out):
    _need_cuda(x, w, bias, out)
    assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype and x.shape == out.shape
    assert w.dtype == torch.float32 and w.shape == (25, x.shape[1]) and x.shape[0] == B * H * W
    _ev = _begin()
    _lib.check(_dwconv5x5(_ptr(x), _ptr(w), _ptr(bias), c_int(B), c_int(H), c_int(W), c_int(x.shape[1]), _ptr(out),
                          c_int(_is_bf16(x)), _stream()), "mmt_dwconv5x5")
    _count(1, "dwconv5x5", _ev)
    return out


_concat_cols = _lib.fn("mmt_concat_cols")


def concat_cols(a, b, out):
    _need_cuda(a, b, out)
    assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous() and a.shape == b.shape and a.dtype == b.dtype
    assert out.shape == (a.shape[0], 2 * a.shape[1]) and out.dtype == a.dtype
    _ev = _begin()
    _lib.check(_concat_cols(_ptr(a), _ptr(b), c_int(a.shape[0]), c_int(a.shape[1]), _ptr(out), c_int(_is_bf16(a)),
                            _stream()), "mmt_concat_cols")
    _count(1, "concat_cols", _ev)
    return out


_spm_rois = _lib.fn("mmt_spm_rois")


def spm_rois(xyxy, scale, rois):
    _need_cuda(xyxy, rois)
    assert xyxy.dtype == torch.float32 and rois.dtype == torch.float32 and xyxy.is_contiguous() and rois.is_contiguous()
    _ev = _begin()
    _lib.check(_spm_rois(_ptr(xyxy), c_int(xyxy.shape[0]), c_float(scale), _ptr(rois), _stream()), "mmt_spm_rois")
    _count(1, "spm_rois", _ev)
    return rois


def prroi_pool(feat, rois, ph, pw, spatial_scale, channels_last=False, out=None):
    """feat fp32 [N,C,H,W] (or [N,H,W,C] with channels_last) ; rois fp32 [R,5] -> [R,C,ph,pw] (or [R,ph*pw,C])."""
    _need_cuda(feat, rois)
    feat = feat.contiguous()
    rois = rois.contiguous()
    assert feat.dtype == torch.float32 and rois.dtype == torch.float32
    if channels_last:
        N, H, W, C = feat.shape
        shape = (rois.shape[0], ph * pw, C)
    else:
        N, C, H, W = feat.shape
        shape = (rois.shape[0], C, ph, pw)
    if out is None:
        out = torch.empty(shape, device=feat.device, dtype=torch.float32)
    _ev = _begin()
    _lib.check(_prroi(_ptr(feat), _ptr(rois), _ptr(out), c_int(rois.shape[0]), c_int(C), c_int(H), c_int(W), c_int(ph),
                      c_int(pw), c_float(spatial_scale), c_int(1 if channels_last else 0), _stream()), "mmt_prroi_fwd")
    _count(1, "prroi_pool", _ev)
    return out


_frame_crop = _lib.fn("mmt_frame_crop")
_frame_crop_ws = _lib.lib.mmt_frame_crop_workspace_bytes
_frame_crop_ws.restype = ctypes.c_longlong
_FRAME_WS = {}
_track_update = _lib.fn("mmt_track_update")


def frame_crop(frame_ptrs, dims, state, factor, out_sz, n_mod, jet_mask=0, jet_lut=None, active=None, out=None,
               out_u8=None, resize_factor=None):
    """Device-side sample_target + Preprocessor (include/mmt_b200.h: mmt_frame_crop).

    frame_ptrs int64 CUDA [n_mod*B] (device addresses of uint8 HWC frames), dims int32 CUDA [n_mod*B, 3] (H, W, pitch),
    state float64 CUDA [B, 4].  Returns (out fp32 [n_mod, B, 3, S, S] or None, out_u8, resize_factor float64 [B])."""
    _need_cuda(frame_ptrs, dims, state)
    B = state.shape[0]
    assert frame_ptrs.dtype == torch.int64 and frame_ptrs.numel() == n_mod * B and frame_ptrs.is_contiguous()
    assert dims.dtype == torch.int32 and tuple(dims.shape) == (n_mod * B, 3) and dims.is_contiguous()
    assert state.dtype == torch.float64 and tuple(state.shape) == (B, 4) and state.is_contiguous()
    if out is None and out_u8 is None:
        out = torch.empty((n_mod, B, 3, out_sz, out_sz), device=state.device, dtype=torch.float32)
    if out is not None:
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n_mod * B * 3 * out_sz * out_sz
    if out_u8 is not None:
        assert out_u8.dtype == torch.uint8 and out_u8.is_contiguous() and out_u8.numel() == n_mod * B * 3 * out_sz * out_sz
    if resize_factor is None:
        resize_factor = torch.empty(B, device=state.device, dtype=torch.float64)
    assert resize_factor.dtype == torch.float64 and resize_factor.numel() == B
    if active is not None:
        assert active.dtype == torch.uint8 and active.numel() == B and active.is_cuda
    if jet_mask:
        assert jet_lut is not None and jet_lut.dtype == torch.uint8 and jet_lut.numel() == 768 and jet_lut.is_cuda
    need = int(_frame_crop_ws(c_int(B), c_int(n_mod), c_int(out_sz)))
    key = (state.device, torch.cuda.current_stream().cuda_stream)
    ws = _FRAME_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _FRAME_WS[key] = torch.empty(need, dtype=torch.uint8, device=state.device)   # stream-ordered reuse
    _ev = _begin()
    _lib.check(_frame_crop(_ptr(frame_ptrs), _ptr(dims), _ptr(state), _ptr(active), c_int(B), c_int(n_mod),
                           ctypes.c_uint(jet_mask), ctypes.c_double(factor), c_int(out_sz), _ptr(jet_lut), _ptr(out),
                           _ptr(out_u8), _ptr(resize_factor), _ptr(ws), ctypes.c_longlong(ws.numel()), _stream()),
               "mmt_frame_crop")
    _count(2, "frame_crop", _ev)
    return out, out_u8, resize_factor


def track_update(pred_cxcywh, resize_factor, dims, state, search_size, margin=10.0, log=None, active=None):
    """state <- clip_box(map_box_back(pred * search_size / resize_factor)) in place (mmt_track_update)."""
    _need_cuda(pred_cxcywh, resize_factor, dims, state)
    B = state.shape[0]
    assert pred_cxcywh.dtype == torch.float32 and pred_cxcywh.numel() == 4 * B and pred_cxcywh.is_contiguous()
    assert state.dtype == torch.float64 and state.is_contiguous() and resize_factor.dtype == torch.float64
    assert dims.dtype == torch.int32 and dims.is_contiguous() and dims.shape[0] >= B
    if log is not None:
        assert log.dtype == torch.float64 and log.numel() == 4 * B and log.is_contiguous() and log.is_cuda
    _ev = _begin()
    _lib.check(_track_update(_ptr(pred_cxcywh), _ptr(resize_factor), _ptr(dims), _ptr(state), _ptr(log), _ptr(active),
                             c_int(B), c_int(search_size), ctypes.c_double(margin), _stream()), "mmt_track_update")
    _count(1, "track_update", _ev)
    return state


_online_score_update = _lib.fn("mmt_online_score_update")


def online_score_update(logits, max_score, take, decay=1.0, active=None):
    """max_score (float64 [B]) and take (uint8 [B]) updated in place from the SPM logits (mmt_online_score_update)."""
    _need_cuda(logits, max_score, take)
    B = max_score.numel()
    assert logits.dtype == torch.float32 and logits.numel() == B and logits.is_contiguous()
    assert max_score.dtype == torch.float64 and take.dtype == torch.uint8 and take.numel() == B
    _ev = _begin()
    _lib.check(_online_score_update(_ptr(logits), _ptr(max_score), _ptr(take), _ptr(active), c_int(B),
                                    ctypes.c_double(decay), _stream()), "mmt_online_score_update")
    _count(1, "online_score_update", _ev)
    return take


_preprocess_u8 = _lib.fn("mmt_preprocess_u8")


def preprocess_u8(crops_u8, out, per_mod, jet_mask=0, jet_lut=None):
    """uint8 [n, S, S, 3] CUDA crops -> normalised fp32 [n, 3, S, S] (mmt_preprocess_u8)."""
    _need_cuda(crops_u8, out)
    n, S = crops_u8.shape[0], crops_u8.shape[1]
    assert crops_u8.dtype == torch.uint8 and crops_u8.is_contiguous() and tuple(crops_u8.shape) == (n, S, S, 3)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n * 3 * S * S
    if jet_mask:
        assert jet_lut is not None and jet_lut.is_cuda and jet_lut.dtype == torch.uint8 and jet_lut.numel() == 768
    _ev = _begin()
    _lib.check(_preprocess_u8(_ptr(crops_u8), _ptr(out), c_int(n), c_int(S), c_int(per_mod), ctypes.c_uint(jet_mask),
                              _ptr(jet_lut), _stream()), "mmt_preprocess_u8")
    _count(1, "preprocess_u8", _ev)
    return out
