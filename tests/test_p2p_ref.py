"""CPU: the point-to-plane restatement (tests/p2p_ref.py) against closed forms."""
import numpy as np
import torch

import p2p_ref


def _K():
    return torch.tensor([[100.0, 0, 39.5], [0, 100.0, 29.5], [0, 0, 1]])


def test_normals_of_a_slanted_plane():
    K = _K()
    H, W = 60, 80
    u = torch.arange(W, dtype=torch.float32)[None].expand(H, W)
    # plane z = 2 / (1 - 0.3 x/z)  <=>  z - 0.3 x = 2 : normal (-0.3, 0, 1) / |.|
    x_over_z = (u - K[0][2]) / K[0][0]
    depth = 2.0 / (1.0 - 0.3 * x_over_z)
    n = p2p_ref.depth_to_normals(depth, K)
    want = torch.tensor([-0.3, 0.0, 1.0]) / np.sqrt(1.09)
    inner = n[2:-2, 2:-2].reshape(-1, 3)
    assert torch.allclose(inner, want.expand_as(inner), atol=2e-4)          # sign: dx x dy points along +z


def test_shifted_plane_gives_the_shift():
    K = _K()
    depth0 = torch.full((60, 80), 2.0)
    depth1 = torch.full((60, 80), 2.004)
    I = torch.eye(4)
    m, d = p2p_ref.point2plane_dist(depth0, depth1, K, I, I, frustum=True, method="max")
    paired = d[~torch.isnan(d)]
    assert len(paired) > 0.8 * 60 * 80
    assert torch.allclose(paired, torch.full_like(paired, 0.004), atol=1e-5)
    assert abs(float(m) - 0.004) < 1e-5
    s, _ = p2p_ref.point2plane_dist(depth0, depth1, K, I, I, method="sum")
    assert abs(float(s) / (len(paired) * 0.004 ** 2) - 1.0) < 2e-3
    far, d2 = p2p_ref.point2plane_dist(depth0, torch.full((60, 80), 2.5), K, I, I, method="sum")
    assert torch.isnan(d2).all() and float(far) == 0.0                     # nothing within 2 cm
