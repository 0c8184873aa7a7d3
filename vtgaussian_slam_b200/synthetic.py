"""Seedable synthetic Replica/TUM/ScanNet++-shaped RGB-D frames and the view-tied
Gaussians the reference would build from them (no dataset needed).

The Gaussian construction follows the reference's own:
  get_pointcloud            src/vtgaussian_slam.py:76-128  (one Gaussian per valid pixel at
                            ((x-cx+0.5)/fx*z, (y-cy+0.5)/fy*z, z), z = depth*1.005,
                            isotropic scale z/((fx+fy)/2))
  initialize_params         src/vtgaussian_slam.py:132-177 (identity quaternions,
                            logit_opacity = 0, log_scale = log(scale), one cam pose per frame)
  edge densification        src/vtgaussian_slam.py:212-252,1025-1044 (extra Gaussians on
                            edge pixels of the 2x grid; Canny+dilate there, a gradient
                            quantile here)
  setup_camera              utils/recon_helpers.py:4-27
numpy only.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: (W, H, fx, fy, cx, cy)   reference configs/data/*.yaml
    "replica": (1200, 680, 600.0, 600.0, 599.5, 339.5),
    "tum_fr1": (640, 480, 517.3, 516.5, 318.6, 255.3),
    "scannetpp": (1752, 1168, 1752 * 0.5 / np.tan(np.deg2rad(35.0)), 1752 * 0.5 / np.tan(np.deg2rad(35.0)), 875.5, 583.5),
}


def intrinsics(shape="replica", width=None, height=None):
    """3x3 K for a named shape, optionally rescaled to width x height."""
    W, H, fx, fy, cx, cy = SHAPES[shape]
    if width is None:
        width, height = W, H
    sx, sy = width / W, height / H
    K = np.array([[fx * sx, 0, (cx + 0.5) * sx - 0.5], [0, fy * sy, (cy + 0.5) * sy - 0.5], [0, 0, 1]], np.float64)
    return int(width), int(height), K


def setup_camera(w, h, k, w2c, near=0.01, far=100.0, bg=(0.0, 0.0, 0.0)):
    """numpy mirror of setup_camera (reference utils/recon_helpers.py:4-27): returns the
    11 GaussianRasterizationSettings fields with matrices in the row-vector convention."""
    fx, fy, cx, cy = k[0][0], k[1][1], k[0][2], k[1][2]
    w2c = np.asarray(w2c, np.float32)
    cam_center = np.linalg.inv(w2c.astype(np.float64))[:3, 3].astype(np.float32)
    view = w2c.T.copy()
    opengl_proj = np.array([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                            [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                            [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                            [0.0, 0.0, 1.0, 0.0]], np.float32).T
    full_proj = (view @ opengl_proj).astype(np.float32)
    return dict(image_height=int(h), image_width=int(w), tanfovx=float(w / (2 * fx)), tanfovy=float(h / (2 * fy)),
                bg=np.asarray(bg, np.float32), scale_modifier=1.0, viewmatrix=view.reshape(1, 4, 4),
                projmatrix=full_proj.reshape(1, 4, 4), sh_degree=0, campos=cam_center, prefiltered=False)


def _room_depth(W, H, K, boxes=True):
    """Analytic box room seen from inside (camera at the origin looking down +z) plus a few
    fronto-parallel slabs: returns depth[H,W] (metres, z-depth) and hit points[H,W,3]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # ray through pixel (x, y) in the rasteriser's convention: setup_camera's projection maps the
    # camera-frame direction ((x - cx + 0.5)/fx, (y - cy + 0.5)/fy, 1) onto pixel centre (x, y)
    x = (np.arange(W, dtype=np.float64) - cx + 0.5) / fx
    y = (np.arange(H, dtype=np.float64) - cy + 0.5) / fy
    dx, dy = np.meshgrid(x, y)
    big = 1e9
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.full((H, W), 4.2)                                   # back wall z = 4.2
        t = np.minimum(t, np.where(dx > 0, 3.1 / dx, big))          # right wall x = 3.1
        t = np.minimum(t, np.where(dx < 0, -2.7 / dx, big))         # left wall x = -2.7
        t = np.minimum(t, np.where(dy > 0, 1.3 / dy, big))          # floor y = 1.3
        t = np.minimum(t, np.where(dy < 0, -1.5 / dy, big))         # ceiling y = -1.5
    if boxes:
        for (x0, x1, y0, y1, z0) in [(-1.6, -0.4, 0.1, 1.3, 2.4), (0.5, 1.7, 0.4, 1.3, 3.0), (-0.3, 0.35, -0.2, 0.5, 1.7)]:
            inside = (dx * z0 > x0) & (dx * z0 < x1) & (dy * z0 > y0) & (dy * z0 < y1)
            t = np.where(inside & (z0 < t), z0, t)
    pts = np.stack([dx * t, dy * t, t], -1)
    return t, pts


def _texture(pts):
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]
    r = 0.55 + 0.25 * np.sin(1.7 * x + 0.9 * z) * np.cos(1.1 * y)
    g = 0.50 + 0.25 * np.sin(2.3 * y + 0.4 * x + 1.0)
    b = 0.45 + 0.25 * np.cos(1.3 * z + 0.8 * x * y)
    chk = ((np.floor(x * 2.0) + np.floor(y * 2.0) + np.floor(z * 2.0)) % 2) * 0.12
    return np.clip(np.stack([r + chk, g - chk, b + 0.5 * chk], 0), 0.0, 1.0)


def make_frame(shape="replica", width=None, height=None, seed=0):
    """-> dict(W,H,K, im[3,H,W] float32 in [0,1], depth[1,H,W] float32 metres)."""
    W, H, K = intrinsics(shape, width, height)
    depth, pts = _room_depth(W, H, K)
    rng = np.random.default_rng(seed)
    im = _texture(pts) + rng.normal(0, 0.004, (3, H, W))
    return dict(W=W, H=H, K=K, im=np.clip(im, 0, 1).astype(np.float32), depth=depth[None].astype(np.float32))


def view_tied_gaussians(frame, n_target=None, n_edge=0, opacity="fresh", seed=2, color_noise=0.02):
    """The reference's per-pixel view-tied Gaussians for `frame`.
    n_target: None -> one per pixel; else the pixel grid is sub-sampled isotropically to about
              n_target Gaussians and sigma scaled by the same factor (coverage stays ~1).
    n_edge:   extra Gaussians on the n_edge strongest edge pixels of the 2x grid (sigma 0.5 px).
    opacity:  'fresh' (logit 0, as a new section) or 'trained' (logit ~ U(0,4)).
    Returns a dict shaped like the reference's `params` (numpy float32)."""
    W, H, K = frame["W"], frame["H"], frame["K"]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    rng = np.random.default_rng(seed)
    if n_target is None or n_target >= W * H:
        f = 1.0
        Wv, Hv = W, H
    else:
        f = np.sqrt(W * H / float(n_target))
        Wv, Hv = int(round(W / f)), int(round(H / f))
    # virtual pixel grid (Wv x Hv) with intrinsics scaled by 1/f
    fxv, fyv = fx / f, fy / f
    cxv, cyv = (cx + 0.5) / f - 0.5, (cy + 0.5) / f - 0.5
    Kv = np.array([[fxv, 0, cxv], [0, fyv, cyv], [0, 0, 1]])
    depth_v, pts_v = _room_depth(Wv, Hv, Kv)
    z = (depth_v * 1.005).reshape(-1)
    xg, yg = np.meshgrid(np.arange(Wv, dtype=np.float64), np.arange(Hv, dtype=np.float64))
    xx = ((xg - cxv + 0.5) / fxv).reshape(-1)
    yy = ((yg - cyv + 0.5) / fyv).reshape(-1)
    means = np.stack([xx * z, yy * z, z], -1)
    scale = z / ((fxv + fyv) / 2.0)
    cols = _texture(pts_v).reshape(3, -1).T
    if n_edge > 0:
        W2, H2 = 2 * W, 2 * H
        K2 = np.array([[2 * fx, 0, 2 * cx + 0.5], [0, 2 * fy, 2 * cy + 0.5], [0, 0, 1]])
        d2, p2 = _room_depth(W2, H2, K2)
        tex2 = _texture(p2)
        gy, gx = np.gradient(d2)
        strength = np.abs(gx) + np.abs(gy)
        for ch in range(3):
            cy_, cx_ = np.gradient(tex2[ch])
            strength = strength + 0.5 * (np.abs(cx_) + np.abs(cy_))
        # 3x3 dilation of the strength map (max filter), as the reference dilates its Canny mask
        pad = np.pad(strength, 1, mode="edge")
        dil = np.max(np.stack([pad[i:i + H2, j:j + W2] for i in range(3) for j in range(3)], 0), 0)
        flat = dil.reshape(-1) + rng.uniform(0, 1e-9, dil.size)
        idx = np.argpartition(-flat, n_edge - 1)[:n_edge]
        idx.sort()
        ye, xe = np.divmod(idx, W2)
        ze = d2.reshape(-1)[idx] * 1.005
        xe_ = (xe - K2[0, 2] + 0.5) / K2[0, 0]
        ye_ = (ye - K2[1, 2] + 0.5) / K2[1, 1]
        means = np.concatenate([means, np.stack([xe_ * ze, ye_ * ze, ze], -1)], 0)
        scale = np.concatenate([scale, ze / ((K2[0, 0] + K2[1, 1]) / 2.0)], 0)
        cols = np.concatenate([cols, tex2.reshape(3, -1).T[idx]], 0)
    N = means.shape[0]
    cols = cols + rng.normal(0, color_noise, cols.shape)
    if opacity == "fresh":
        logit = np.zeros((N, 1))
    else:
        logit = rng.uniform(0.0, 4.0, (N, 1))
    rots = np.tile(np.array([1.0, 0, 0, 0]), (N, 1))
    return dict(means3D=means.astype(np.float32), rgb_colors=cols.astype(np.float32),
                unnorm_rotations=rots.astype(np.float32), logit_opacities=logit.astype(np.float32),
                log_scales=np.log(scale)[:, None].astype(np.float32))


def perturbed_pose(seed=1, trans_sigma=0.01, rot_deg=0.5):
    """Small camera perturbation (tracking start): -> (cam_unnorm_rot[4], cam_trans[3])."""
    rng = np.random.default_rng(seed)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rot_deg)
    q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])
    t = rng.normal(0, trans_sigma, 3)
    return q.astype(np.float32), t.astype(np.float32)


def random_scene(n, width, height, seed=0, anisotropic=True, opacity_range=(0.05, 1.0), scale_px=(0.5, 6.0)):
    """A generic random scene for parity tests (NOT view-tied): Gaussians scattered in the
    frustum with random anisotropic scales/rotations, some behind the camera / off-screen.
    Returns (K, arrays dict) with activated render variables."""
    rng = np.random.default_rng(seed)
    f = 0.9 * width
    K = np.array([[f, 0, (width - 1) / 2.0], [0, f, (height - 1) / 2.0], [0, 0, 1.0]])
    z = rng.uniform(0.1, 6.0, n)
    x = (rng.uniform(-0.15, 1.15, n) * width - K[0, 2]) / f * z
    y = (rng.uniform(-0.15, 1.15, n) * height - K[1, 2]) / f * z
    means = np.stack([x, y, z], -1)
    px = rng.uniform(scale_px[0], scale_px[1], (n, 1))
    if anisotropic:
        ratio = rng.uniform(0.3, 1.0, (n, 3))
    else:
        ratio = np.ones((n, 3))
    scales = px * ratio * z[:, None] / f
    q = rng.normal(size=(n, 4)) if anisotropic else np.tile([1.0, 0, 0, 0], (n, 1))
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    op = rng.uniform(opacity_range[0], opacity_range[1], n)
    col = rng.uniform(0, 1, (n, 3))
    return K, dict(means3D=means.astype(np.float32), scales=scales.astype(np.float32), rotations=q.astype(np.float32),
                   opacities=op.astype(np.float32), colors=col.astype(np.float32))
