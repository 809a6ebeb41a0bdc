"""GPU: the product kernels against the reference's OWN native kernels, compiled unmodified from /root/reference into
oracle/_ref/ by oracle/build_ref.py (the libraries travel with the repo snapshot; /root/reference is never read here):
  * mmt_prroi_fwd        vs  PrRoIPoolingForwardGpu   (external/PreciseRoIPooling/src/prroi_pooling_gpu_impl.cu:387-402)
  * mmt_msda_fwd         vs  ms_deformable_im2col_cuda (deformable_attention/ops/src/cuda/ms_deform_im2col_cuda.cuh:924)
and, transitively, the numpy / torch oracles of both ops against the same kernels."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _load(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (run `python oracle/build_ref.py` where /root/reference exists)")
    return ctypes.CDLL(path)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("case", [(2, 8, 18, 18, 4, 4, 1.0), (4, 16, 24, 32, 7, 7, 0.5), (1, 1024, 24, 24, 4, 4, 1.0),
                                  (3, 5, 9, 13, 3, 2, 0.25)])
def test_prroi_matches_the_reference_kernel(built_lib, case):
    from mmt_b200 import ops
    from oracle import native_ops_oracle as NO
    lib = _load("libprroi_ref.so")
    lib.PrRoIPoolingForwardGpu.restype = None
    N, C, H, W, ph, pw, scale = case
    g = torch.Generator().manual_seed(N * 100 + C)
    feat = torch.randn(N, C, H, W, generator=g)
    R = 9
    x0 = torch.rand(R, generator=g) * W / scale * 0.8 - 2.0           # partly outside the map on every side
    y0 = torch.rand(R, generator=g) * H / scale * 0.8 - 2.0
    bw = torch.rand(R, generator=g) * W / scale * 0.6
    bh = torch.rand(R, generator=g) * H / scale * 0.6
    bw[0] = 0.0                                                       # empty box -> zeros
    rois = torch.stack([torch.randint(0, N, (R,), generator=g).float(), x0, y0, x0 + bw, y0 + bh], 1).contiguous()
    fc, rc = feat.cuda(), rois.cuda()
    ref = torch.full((R, C, ph, pw), float("nan"), device="cuda")
    lib.PrRoIPoolingForwardGpu(_stream(), _p(fc), _p(rc), _p(ref), C, H, W, ph, pw, ctypes.c_float(scale), R * C * ph * pw)
    out = ops.prroi_pool(fc, rc, ph, pw, scale)
    torch.cuda.synchronize()
    tol = 1e-5 * max(1.0, ref.abs().max().item())
    assert torch.isfinite(ref).all()
    assert (out - ref).abs().max().item() <= tol
    # token layout used by the SPM head == the reference layout transposed
    out_cl = ops.prroi_pool(fc.permute(0, 2, 3, 1).contiguous(), rc, ph, pw, scale, channels_last=True)
    assert (out_cl.view(R, ph, pw, C).permute(0, 3, 1, 2) - ref).abs().max().item() <= tol
    # and the numpy restatement the CPU tests rely on
    if C <= 16:
        assert np.abs(NO.prroi_pool_forward(feat.numpy(), rois.numpy(), ph, pw, scale) - ref.cpu().numpy()).max() <= tol


@pytest.mark.parametrize("cfg", [(2, 8, 64, 648, [(18, 18), (18, 18)], 4, 2), (1, 2, 2, 2, [(6, 4), (3, 2)], 2, 1),
                                 (4, 4, 32, 50, [(5, 7), (9, 3), (2, 2)], 3, 2)])
def test_msda_matches_the_reference_kernel(built_lib, cfg):
    from mmt_b200 import ops
    from oracle import mixformer_oracle as O
    lib = _load("libmsda_ref.so")
    lib.msda_ref_forward.restype = ctypes.c_int
    N, M, D, Lq, shapes, P, step = cfg
    g = torch.Generator().manual_seed(Lq + 7 * M)
    S, L = sum(h * w for h, w in shapes), len(shapes)
    value = torch.randn(N, S, M, D, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.4 - 0.2                  # includes out-of-map samples
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    ss = torch.tensor(shapes, dtype=torch.int64)
    lsi = torch.cat([ss.new_zeros(1), (ss[:, 0] * ss[:, 1]).cumsum(0)[:-1]])
    vc, lc, ac = value.cuda().contiguous(), loc.cuda().contiguous(), attn.cuda().contiguous()
    ref = torch.zeros(N, Lq, M * D, device="cuda")
    ssc, lsic = ss.cuda(), lsi.cuda()                 # keep the device tables alive across the asynchronous launch
    st = lib.msda_ref_forward(_p(vc), _p(ssc), _p(lsic), _p(lc), _p(ac), _p(ref), N, S, M, D, L, Lq, P, step, _stream())
    assert st == 0
    out = ops.msda(vc, shapes, lc, ac)
    torch.cuda.synchronize()
    tol = 1e-6 + 1e-5 * ref.abs().max().item()
    assert (out - ref).abs().max().item() <= tol
    assert (O.msda_core(value, shapes, loc, attn) - ref.cpu()).abs().max().item() <= tol      # the torch oracle too
