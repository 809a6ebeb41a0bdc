"""CPU: the reference's own known-answer test of PrRoIPool, restated against the oracle
(external/PreciseRoIPooling/pytorch/tests/test_prroi_pooling2d.py:21-35): PrRoIPool(7, 7, scale 0.5) over rois
[0,0,0,14,14] and [1,14,14,28,28] of a rand(4,16,24,32) map equals avg_pool2d(k=2, s=1) windows of the map."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import native_ops_oracle as NO


def test_prroi_oracle_matches_avg_pool_kat():
    torch.manual_seed(0)
    features = torch.rand(4, 16, 24, 32)
    rois = np.array([[0, 0, 0, 14, 14], [1, 14, 14, 28, 28]], dtype=np.float32)
    out = NO.prroi_pool_forward(features.numpy(), rois, 7, 7, 0.5)
    golden = F.avg_pool2d(features, kernel_size=2, stride=1)
    assert np.allclose(out[0], golden[0, :, :7, :7].numpy(), atol=1e-5)
    assert np.allclose(out[1], golden[1, :, 7:14, 7:14].numpy(), atol=1e-5)


def test_prroi_oracle_edge_cases():
    f = np.random.default_rng(1).standard_normal((1, 3, 6, 5)).astype(np.float32)
    rois = np.array([[0, 1.0, 1.0, 1.0, 4.0],        # zero width -> zeros
                     [0, -3.0, -2.0, 2.5, 3.5],      # partly outside the map: outside reads are zeros
                     [0, 0.0, 0.0, 4.0, 5.0]], dtype=np.float32)
    out = NO.prroi_pool_forward(f, rois, 2, 2, 1.0)
    assert np.all(out[0] == 0)
    assert np.isfinite(out).all()
    # whole-map integral of a bilinear surface over [0,W-1]x[0,H-1] equals the trapezoid rule
    ones = np.ones((1, 1, 6, 5), dtype=np.float32)
    full = NO.prroi_pool_forward(ones, np.array([[0, 0, 0, 4, 5]], dtype=np.float32), 1, 1, 1.0)
    assert abs(float(full[0, 0, 0, 0]) - 1.0) < 1e-6
