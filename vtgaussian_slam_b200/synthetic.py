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


ROOM = dict(x=(-2.7, 3.1), y=(-1.5, 1.3), z=(-1.6, 4.2))                   # axis-aligned box room (metres, world frame)
SLABS = [(-1.6, -0.4, 0.1, 1.3, 2.4), (0.5, 1.7, 0.4, 1.3, 3.0), (-0.3, 0.35, -0.2, 0.5, 1.7)]   # (x0, x1, y0, y1, z) rectangles


def _room_depth(W, H, K, boxes=True, c2w=None):
    """Analytic box room seen from inside plus a few slabs parallel to the world xy-plane, ray-cast from the camera
    `c2w` (4x4 camera-to-world, default: at the origin looking down +z): returns the z-depth[H,W] in the camera
    frame (metres) and the WORLD hit points[H,W,3]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # ray through pixel (x, y) in the rasteriser's convention: setup_camera's projection maps the
    # camera-frame direction ((x - cx + 0.5)/fx, (y - cy + 0.5)/fy, 1) onto pixel centre (x, y)
    x = (np.arange(W, dtype=np.float64) - cx + 0.5) / fx
    y = (np.arange(H, dtype=np.float64) - cy + 0.5) / fy
    dx, dy = np.meshgrid(x, y)
    if c2w is None:
        o = np.zeros(3)
        d = (dx, dy, np.ones_like(dx))
    else:
        c2w = np.asarray(c2w, np.float64)
        o, R = c2w[:3, 3], c2w[:3, :3]
        d = tuple(R[i, 0] * dx + R[i, 1] * dy + R[i, 2] for i in range(3))
    big = 1e9
    # the camera-frame direction has z = 1, so the ray parameter t IS the z-depth
    t = np.full((H, W), big)
    with np.errstate(divide="ignore", invalid="ignore"):
        for axis, (lo, hi) in zip(range(3), (ROOM["x"], ROOM["y"], ROOM["z"])):
            t = np.minimum(t, np.where(d[axis] > 0, (hi - o[axis]) / d[axis], big))
            t = np.minimum(t, np.where(d[axis] < 0, (lo - o[axis]) / d[axis], big))
        if boxes:
            for (x0, x1, y0, y1, z0) in SLABS:
                ts = np.where(np.abs(d[2]) > 1e-12, (z0 - o[2]) / d[2], big)
                hx, hy = o[0] + d[0] * ts, o[1] + d[1] * ts
                hit = (ts > 0) & (hx > x0) & (hx < x1) & (hy > y0) & (hy < y1) & (ts < t)
                t = np.where(hit, ts, t)
    pts = np.stack([o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t], -1)
    return t, pts


def _texture(pts):
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]
    r = 0.55 + 0.25 * np.sin(1.7 * x + 0.9 * z) * np.cos(1.1 * y)
    g = 0.50 + 0.25 * np.sin(2.3 * y + 0.4 * x + 1.0)
    b = 0.45 + 0.25 * np.cos(1.3 * z + 0.8 * x * y)
    chk = ((np.floor(x * 2.0) + np.floor(y * 2.0) + np.floor(z * 2.0)) % 2) * 0.12
    return np.clip(np.stack([r + chk, g - chk, b + 0.5 * chk], 0), 0.0, 1.0)


def make_frame(shape="replica", width=None, height=None, seed=0, c2w=None):
    """-> dict(W,H,K, im[3,H,W] float32 in [0,1], depth[1,H,W] float32 metres[, c2w]) seen from camera pose c2w."""
    W, H, K = intrinsics(shape, width, height)
    depth, pts = _room_depth(W, H, K, c2w=c2w)
    rng = np.random.default_rng(seed)
    im = _texture(pts) + rng.normal(0, 0.004, (3, H, W))
    fr = dict(W=W, H=H, K=K, im=np.clip(im, 0, 1).astype(np.float32), depth=depth[None].astype(np.float32))
    if c2w is not None:
        fr["c2w"] = np.asarray(c2w, np.float64)
    return fr


def trajectory(num_frames, step_m=0.01, step_deg=0.3, seed=3):
    """A smooth hand-held-like camera path inside the room: c2w[num_frames,4,4], frame 0 = identity (dataset poses
    are relative to the first frame, reference datasets/gradslam_datasets/basedataset.py:288-292).  Consecutive
    frames are ~step_m apart and ~step_deg rotated; the path is a sum of slow sinusoids, so constant-velocity
    propagation is a good but not exact initial guess."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 6)
    s = np.arange(num_frames, dtype=np.float64)
    span = max(num_frames - 1, 1)
    # amplitudes chosen so that the mean per-frame step is ~step_m / ~step_deg regardless of the length
    A_t = step_m * span / 4.0
    A_r = np.deg2rad(step_deg) * span / 4.0
    A_t, A_r = min(A_t, 0.6), min(A_r, np.deg2rad(20.0))
    w = 2 * np.pi / max(span, 8) * np.array([1.0, 0.7, 1.3, 0.9, 1.1, 0.6])
    if A_t < step_m * span / 4.0:                  # long sequences: keep the per-frame step by cycling faster
        w = w * (step_m * span / 4.0) / A_t
    tr = np.stack([A_t * (np.sin(w[0] * s + ph[0]) - np.sin(ph[0])),
                   0.4 * A_t * (np.sin(w[1] * s + ph[1]) - np.sin(ph[1])),
                   0.7 * A_t * (np.sin(w[2] * s + ph[2]) - np.sin(ph[2]))], -1)
    ang = np.stack([0.5 * A_r * (np.sin(w[3] * s + ph[3]) - np.sin(ph[3])),
                    A_r * (np.sin(w[4] * s + ph[4]) - np.sin(ph[4])),
                    0.3 * A_r * (np.sin(w[5] * s + ph[5]) - np.sin(ph[5]))], -1)
    out = np.tile(np.eye(4), (num_frames, 1, 1))
    for i in range(num_frames):
        ax, ay, az = ang[i]
        Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        out[i, :3, :3] = Rz @ Ry @ Rx
        out[i, :3, 3] = tr[i]
    return out


def make_sequence(shape="tum_fr1", num_frames=8, width=None, height=None, step_m=0.01, step_deg=0.3, seed=0):
    """-> list of frames (make_frame dicts carrying their ground-truth 'c2w') along `trajectory`."""
    poses = trajectory(num_frames, step_m, step_deg, seed=seed + 3)
    return [make_frame(shape, width, height, seed=seed + i, c2w=poses[i]) for i in range(num_frames)]


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


def section_gaussians(frame, c2w=None, opacity="trained", seed=2, color_noise=0.02, factor=1.005):
    """View-tied Gaussians of one posed frame in the WORLD frame, built from the frame's own depth image the way
    the reference does at a section start (get_pointcloud + initialize_params, src/vtgaussian_slam.py:76-177):
    numpy twin of slam_loop.section_from_frame."""
    W, H, K = frame["W"], frame["H"], frame["K"]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    c2w = np.eye(4) if c2w is None else np.asarray(c2w, np.float64)
    rng = np.random.default_rng(seed)
    z = frame["depth"].reshape(-1).astype(np.float64) * factor
    xx = np.tile((np.arange(W) - cx + 0.5) / fx, H)
    yy = np.repeat((np.arange(H) - cy + 0.5) / fy, W)
    pts = np.stack([xx * z, yy * z, z], -1) @ c2w[:3, :3].T + c2w[:3, 3]
    keep = z > 0
    n = int(keep.sum())
    cols = frame["im"].reshape(3, -1).T[keep] + rng.normal(0, color_noise, (n, 3))
    logit = np.zeros((n, 1)) if opacity == "fresh" else rng.uniform(0.0, 4.0, (n, 1))
    rots = np.tile(np.array([1.0, 0, 0, 0]), (n, 1))
    scale = z[keep] / ((fx + fy) / 2.0)
    return dict(means3D=pts[keep].astype(np.float32), rgb_colors=cols.astype(np.float32), unnorm_rotations=rots.astype(np.float32),
                logit_opacities=logit.astype(np.float32), log_scales=np.log(scale)[:, None].astype(np.float32))


def multi_section_scene(shape="scannetpp", sections=4, spacing_m=0.35, width=None, height=None, seed=0, opacity="trained"):
    """BASELINE config 5 shape: `sections` overlapping view-tied sections created from poses ~spacing_m apart along
    `trajectory`, concatenated (the reference concatenates the selected sections' params before rendering,
    src/vtgaussian_slam.py:2734).  Returns (frames, poses[c2w], params) -- all sections in the world frame."""
    poses = trajectory(sections, step_m=spacing_m, step_deg=4.0, seed=seed + 3)
    frames = [make_frame(shape, width, height, seed=seed + i, c2w=poses[i]) for i in range(sections)]
    parts = [section_gaussians(frames[i], poses[i], opacity=opacity, seed=seed + 10 + i) for i in range(sections)]
    params = {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}
    return frames, poses, params


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
