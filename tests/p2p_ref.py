"""Test infrastructure: a CPU restatement of the reference's compute_point2plane_dist (src/vtgaussian_slam.py:1070-1155)
with its two absent third-party pieces restated from their published algorithms --
  * kornia.geometry.depth_to_normals (kornia 0.7): unproject every pixel at integer coordinates (depth_to_3d, no
    half-pixel offset), Sobel spatial_gradient (normalised by 8, replicated borders), cross(dx, dy), F.normalize;
  * open3d.pipelines.registration.evaluate_registration: per source point the nearest target point within the
    threshold (KD-tree hybrid search, max_nn = 1) -- here scipy's cKDTree.
Only tests/ use this; the product path is vtgaussian_slam_b200.keyframes.point2plane_dist (CUDA).  Parity unpinned
against kornia / Open3D themselves (neither is installed)."""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.spatial import cKDTree


def depth_to_normals(depth, K):
    """depth[H,W], K 3x3 -> normals[H,W,3] in the camera frame."""
    H, W = depth.shape
    u = torch.arange(W, dtype=torch.float32)[None].expand(H, W)
    v = torch.arange(H, dtype=torch.float32)[:, None].expand(H, W)
    xyz = torch.stack(((u - K[0][2]) / K[0][0] * depth, (v - K[1][2]) / K[1][1] * depth, depth))      # [3,H,W]
    sx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0
    k = torch.stack((sx, sx.T))[:, None]                                                                # [2,1,3,3]
    g = F.conv2d(F.pad(xyz[:, None], (1, 1, 1, 1), mode="replicate"), k)                                # [3,2,H,W]
    n = torch.cross(g[:, 0], g[:, 1], dim=0)
    return F.normalize(n, dim=0, p=2).permute(1, 2, 0)


def _points(depth, K, w2c):
    H, W = depth.shape
    xx = ((torch.arange(W, dtype=torch.float32) - K[0][2] + 0.5) / K[0][0]).repeat(H)
    yy = ((torch.arange(H, dtype=torch.float32) - K[1][2] + 0.5) / K[1][1]).repeat_interleave(W)
    z = depth.reshape(-1)
    cam = torch.stack((xx * z, yy * z, z, torch.ones_like(z)), -1)
    return (torch.inverse(w2c) @ cam.T).T[:, :3]


def _frustum(w2c, K, pts, H, W):
    cam = (w2c @ torch.cat((pts, torch.ones_like(pts[:, :1])), 1).T).T[:, :3]
    uv = (K @ cam.T).T
    z = uv[:, 2] + 1e-8
    u, v = uv[:, 0] / z, uv[:, 1] / z
    return (u < W) & (u > 0) & (v < H) & (v > 0) & (z > 0)


def point2plane_dist(depth0, depth1, K, w2c0, w2c1, frustum=True, method="sum", threshold=0.02):
    """-> (metric, per-source signed distances with NaN for unpaired points, over ALL pixels of frame 1)."""
    depth0, depth1, K = depth0.reshape(depth0.shape[-2:]).float(), depth1.reshape(depth1.shape[-2:]).float(), K.float()
    H, W = depth0.shape
    R0 = torch.inverse(w2c0)[:3, :3]
    n0 = depth_to_normals(depth0, K).reshape(-1, 3) @ R0.T
    p0, p1 = _points(depth0, K, w2c0), _points(depth1, K, w2c1)
    ok0, ok1 = depth0.reshape(-1) > 0, depth1.reshape(-1) > 0
    if frustum:
        ok0 &= _frustum(w2c1, K, p0, H, W)
        ok1 &= _frustum(w2c0, K, p1, H, W)
    out = torch.full((H * W,), float("nan"))
    i0, i1 = torch.where(ok0)[0], torch.where(ok1)[0]
    if len(i0) and len(i1):
        d, j = cKDTree(p0[i0].double().numpy()).query(p1[i1].double().numpy(), k=1, distance_upper_bound=threshold)
        hit = np.isfinite(d)
        src, tgt = i1[torch.as_tensor(hit)], i0[torch.as_tensor(j[hit])]
        out[src] = (n0[tgt] * (p1[src] - p0[tgt])).sum(1)
    d = out[~torch.isnan(out)]
    if method == "sum":
        m = (d ** 2).sum()
    elif method == "max":
        m = d.abs().max()
    else:
        m = d.abs().topk(100)[0].mean()
    return m, out
