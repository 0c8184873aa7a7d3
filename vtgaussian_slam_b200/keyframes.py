"""Overlap-based keyframe selection (utils/keyframe_selection.py:10-120), batched and
device-agnostic: the current frame's sampled pixels are back-projected once and tested against ALL keyframes in one
batched projection on whatever device the inputs live on -- the reference loops over keyframes in Python and moves
every keyframe's tensors to the GPU and back on each call (:68-71, :108-112).

Semantics kept exactly (including two quirks):
  * pixels are sampled WITH replacement from the valid-depth pixels (`torch.randint`), and `get_pointcloud` then drops
    every point whose coordinates, rounded to 1e-4 and made absolute, coincide with another point's or with the origin
    (:29-37) -- so repeated samples remove each other;
  * a point counts as inside a keyframe when its projection is more than `edge` pixels from every border and its
    depth (+1e-5) is positive; keyframes are ranked by that fraction (stable, descending), those with fraction 0 are
    dropped, the first k ids are returned.
"""
from __future__ import annotations

import torch


def backproject_samples(depth, intrinsics, w2c, sampled_indices):
    """World points of the sampled (row, col) pixels of depth[1,H,W] (reference get_pointcloud of this module: no
    half-pixel offset here, unlike the section builder), minus coincident points."""
    fx, fy, cx, cy = intrinsics[0][0], intrinsics[1][1], intrinsics[0][2], intrinsics[1][2]
    rows, cols = sampled_indices[:, 0], sampled_indices[:, 1]
    z = depth[0, rows, cols]
    cam = torch.stack(((cols - cx) / fx * z, (rows - cy) / fy * z, z), dim=-1)
    hom = torch.cat([cam, torch.ones_like(cam[:, :1])], dim=1)
    pts = (torch.inverse(w2c) @ hom.T).T[:, :3]
    keyed = torch.cat([torch.abs(torch.round(pts, decimals=4)), torch.zeros((1, 3), dtype=pts.dtype, device=pts.device)], dim=0)
    _, inverse, counts = keyed.unique(dim=0, return_inverse=True, return_counts=True)
    duplicated = torch.isin(inverse, torch.where(counts.gt(1))[0])[:pts.shape[0]]
    return pts[~duplicated]


def overlap_fractions(pts, intrinsics, keyframe_w2c, width, height, edge=20):
    """Fraction of `pts` that projects inside each keyframe: keyframe_w2c[Kf,4,4] -> [Kf]."""
    hom = torch.cat([pts, torch.ones_like(pts[:, :1])], dim=1)                       # [n, 4]
    cam = (keyframe_w2c @ hom.T).transpose(1, 2)[..., :3]                             # [Kf, n, 3]
    proj = cam @ intrinsics.T                                                         # [Kf, n, 3]
    zed = proj[..., 2:] + 1e-5
    uv = proj[..., :2] / zed
    inside = (uv[..., 0] < width - edge) & (uv[..., 0] > edge) & (uv[..., 1] < height - edge) & (uv[..., 1] > edge) & (zed[..., 0] > 0)
    return inside.sum(dim=1) / max(pts.shape[0], 1) if pts.shape[0] else torch.full((keyframe_w2c.shape[0],), float("nan"), device=pts.device)


def keyframe_selection_overlap(gt_depth, w2c, intrinsics, keyframe_list, k, pixels=1600, edge_value=20, save_percent=False,
                               generator=None):
    """Signature and results of the reference's keyframe_selection_overlap; `keyframe_list[i]['est_w2c']` is the only
    field read.  `generator` (optional) seeds the pixel sampling; by default torch's global generator is used, like
    the reference."""
    width, height = gt_depth.shape[2], gt_depth.shape[1]
    valid = torch.stack(torch.where(gt_depth[0] > 0), dim=1)
    pick = torch.randint(valid.shape[0], (pixels,), generator=generator)
    pts = backproject_samples(gt_depth, intrinsics, w2c, valid[pick.to(valid.device)])
    if not keyframe_list:
        return []
    stack = torch.stack([kf["est_w2c"].to(pts.device) for kf in keyframe_list])
    frac = overlap_fractions(pts, intrinsics[:3, :3].to(pts.device), stack, width, height, edge_value)
    order = torch.sort(frac, descending=True, stable=True).indices.tolist()
    ranked = [{"id": i, "percent_inside": frac[i]} for i in order]
    if save_percent:
        return ranked
    return [r["id"] for r in ranked if r["percent_inside"] > 0.0][:k]


# ---- overlap-visibility mask of the tracking loss (reference src/vtgaussian_slam.py:376-404, :536-583) --------------
def frame_points(gt_depth, intrinsics, w2c):
    """World points of EVERY pixel of gt_depth[1,H,W] in row-major order (the reference's get_pointcloud_forvismask,
    :538-556: `depth >= 0` keeps all pixels; pixel centres without the half-pixel offset)."""
    H, W = gt_depth.shape[1], gt_depth.shape[2]
    fx, fy, cx, cy = intrinsics[0][0], intrinsics[1][1], intrinsics[0][2], intrinsics[1][2]
    rows = torch.arange(H, device=gt_depth.device).repeat_interleave(W)
    cols = torch.arange(W, device=gt_depth.device).repeat(H)
    z = gt_depth[0].reshape(-1)
    cam = torch.stack(((cols - cx) / fx * z, (rows - cy) / fy * z, z), dim=-1)
    hom = torch.cat([cam, torch.ones_like(cam[:, :1])], dim=1)
    return (torch.inverse(w2c) @ hom.T).T[:, :3]


def get_vis_mask(overlap_w2c, pts, intrinsics, overlap_gtdepth, vis_mask_thres, height, width):
    """mask[H,W]: pixel i of the current frame (its world point pts[i]) is seen by the overlapping keyframe when
    the keyframe's depth, bilinearly sampled (zeros outside, align_corners) at the projection, agrees with the
    projected depth to within vis_mask_thres x the smaller of the two (reference get_vis_mask)."""
    hom = torch.cat([pts, torch.ones_like(pts[:, :1])], dim=1)
    cam = (overlap_w2c @ hom.T).T[:, :3]
    proj = (intrinsics @ cam.T).T
    z = proj[:, 2] + 1e-5
    u, v = proj[:, 0] / z, proj[:, 1] / z
    grid = torch.stack((u / (width - 1) * 2.0 - 1.0, v / (height - 1) * 2.0 - 1.0), dim=-1).reshape(1, 1, -1, 2)
    seen = torch.nn.functional.grid_sample(overlap_gtdepth.to(grid.device).reshape(1, 1, height, width), grid,
                                           padding_mode="zeros", align_corners=True).reshape(-1)
    return (torch.abs(seen - z) < vis_mask_thres * torch.minimum(seen, z)).reshape(overlap_gtdepth.shape[1:])


def tracking_vis_mask(gt_depth, intrinsics, curr_w2c, overlaps, vis_mask_thres=0.05):
    """OR of get_vis_mask over the overlapping keyframes [(w2c, gt_depth), ...] (one for TUM, first / mid / last for
    ScanNet(++), reference :563-574) -> bool[1,H,W], ready for slam_ops.get_loss(vis_mask=...) /
    FusedRenderer.tracking_loss(pixel_mask=...)."""
    H, W = gt_depth.shape[1], gt_depth.shape[2]
    pts = frame_points(gt_depth, intrinsics, curr_w2c)
    out = torch.zeros((H, W), dtype=torch.bool, device=gt_depth.device)
    for w2c, depth in overlaps:
        out |= get_vis_mask(w2c, pts, intrinsics, depth, vis_mask_thres, H, W)
    return out[None]


def overlap_fractions_visible(pts, intrinsics, keyframe_w2c, keyframe_depth, width, height, edge=20, kf_depth_thresh=0.01,
                              chunk=8):
    """Like overlap_fractions, with the keyframe's own depth map deciding visibility (reference
    keyframe_selection_overlap_visbased, utils/keyframe_selection.py:121-229): a point counts when it projects
    inside the margins AND the keyframe's depth sampled there agrees with its projected depth to within
    kf_depth_thresh x the smaller of the two.  keyframe_depth[Kf,1,H,W]; keyframes are processed `chunk` at a time."""
    hom = torch.cat([pts, torch.ones_like(pts[:, :1])], dim=1)
    out = []
    for a in range(0, keyframe_w2c.shape[0], chunk):
        w2c, dep = keyframe_w2c[a:a + chunk], keyframe_depth[a:a + chunk].to(pts.device)
        cam = (w2c @ hom.T).transpose(1, 2)[..., :3]
        proj = cam @ intrinsics.T
        z = proj[..., 2] + 1e-5
        u, v = proj[..., 0] / z, proj[..., 1] / z
        inside = (u < width - edge) & (u > edge) & (v < height - edge) & (v > edge) & (z > 0)
        grid = torch.stack((u / (width - 1) * 2.0 - 1.0, v / (height - 1) * 2.0 - 1.0), dim=-1)[:, None]        # [c, 1, n, 2]
        seen = torch.nn.functional.grid_sample(dep.reshape(-1, 1, height, width), grid, padding_mode="zeros",
                                               align_corners=True)[:, 0, 0]
        visible = torch.abs(seen - z) < kf_depth_thresh * torch.minimum(seen, z)
        out.append((inside & visible).sum(dim=1) / max(pts.shape[0], 1))
    return torch.cat(out) if out else torch.zeros(0, device=pts.device)


def keyframe_selection_overlap_visbased(gt_depth, w2c, intrinsics, keyframe_list, k, pixels=1600, edge_value=20,
                                        save_percent=False, kf_depth_thresh=0.01, earliest_thres=0.5):
    """Signature and results of the reference's function of the same name (the variant its main loop imports): every
    valid-depth pixel is used (`pixels` is ignored, as upstream), keyframes need 'est_w2c' and 'depth';
    -> (ids of the k best-overlapping keyframes, [the LOWEST-ranked keyframe whose overlap still exceeds
    earliest_thres] or, if none does, the first list again)."""
    width, height = gt_depth.shape[2], gt_depth.shape[1]
    valid = torch.stack(torch.where(gt_depth[0] > 0), dim=1)
    pts = backproject_samples(gt_depth, intrinsics, w2c, valid)
    if not keyframe_list:
        return [] if save_percent else ([], [])
    stack = torch.stack([kf["est_w2c"].to(pts.device) for kf in keyframe_list])
    depths = torch.stack([kf["depth"].reshape(1, height, width) for kf in keyframe_list])
    frac = overlap_fractions_visible(pts, intrinsics[:3, :3].to(pts.device), stack, depths, width, height, edge_value, kf_depth_thresh)
    order = torch.sort(frac, descending=True, stable=True).indices.tolist()
    ranked = [{"id": i, "percent_inside": frac[i]} for i in order]
    if save_percent:
        return ranked
    selected = [r["id"] for r in ranked if r["percent_inside"] > 0.0][:k]
    earliest = [r["id"] for r in ranked if r["percent_inside"] > earliest_thres][-1:]
    return selected, (earliest if earliest else selected)


# ---- section ("base frame") choice of the main loop (reference utils/keyframe_selection.py:568-703, called at
# src/vtgaussian_slam.py:1549-1553 to pick the sections a new frame is tracked against) ---------------------------------
def quantize_selected_time_idx(selected_time_idx, num_frames_each_base_frame):
    """Keyframe ids -> ids of the sections they belong to, without duplicates (reference :568-577; it returns the set's
    iteration order, every caller sorts it)."""
    return list({int(i / num_frames_each_base_frame) for i in selected_time_idx})


def keyframe_selection_overlap_visbased_earliest_dynamic_new_topkbase(gt_depth, w2c, intrinsics, keyframe_list, k, config, pixels=1600,
                                                                      edge_value=20, kf_depth_thresh=0.01, earliest_thres=0.5,
                                                                      lower_earliest_thres_percent=0.8, topk_base=3):
    """Signature and results of the reference's function of the same name: the visibility-based overlap of every keyframe
    with the current frame (all valid-depth pixels; `k` and `pixels` are ignored, as upstream), a threshold that starts at
    `earliest_thres` and is lowered by `lower_earliest_thres_percent` until at least three sections hold a keyframe above
    it (or the list is short, or the threshold falls under 1 %), and then the EARLIEST `topk_base` such sections (the
    earliest one alone when topk_base is None).  config needs 'baseframe_every' and 'overlap_every'."""
    width, height = gt_depth.shape[2], gt_depth.shape[1]
    valid = torch.stack(torch.where(gt_depth[0] > 0), dim=1)
    pts = backproject_samples(gt_depth, intrinsics, w2c, valid)
    stack = torch.stack([kf["est_w2c"].to(pts.device) for kf in keyframe_list])
    depths = torch.stack([kf["depth"].reshape(1, height, width) for kf in keyframe_list])
    frac = overlap_fractions_visible(pts, intrinsics[:3, :3].to(pts.device), stack, depths, width, height, edge_value, kf_depth_thresh).tolist()
    per_section = int(config["baseframe_every"] / config["overlap_every"])
    thres, first = earliest_thres, True
    while True:
        thres = thres if first else lower_earliest_thres_percent * thres
        first = False
        above = [i for i, f in enumerate(frac) if f > thres]
        sections = sorted(quantize_selected_time_idx(above, per_section))
        if len(sections) >= 3 or (len(frac) <= 3 * per_section and len(sections) > 0) or thres < 0.01:
            break
    if not above:
        above = [len(frac) - 1]                          # nothing overlaps: the latest keyframe
    above = sorted(above)
    if topk_base is None:
        return sorted(quantize_selected_time_idx(above[:1], per_section))
    sections = sorted(quantize_selected_time_idx(above, per_section))
    return sections[:min(topk_base, len(sections))]


# ---- point-to-plane metric on the device (reference compute_point2plane_dist, src/vtgaussian_slam.py:1070-1155) --------
def _rows12(m):
    import ctypes as C
    a = torch.as_tensor(m).detach().to(torch.float64).cpu().reshape(4, 4)[:3].reshape(-1).tolist()
    return (C.c_float * 12)(*a)


def point2plane_dist(latest_depth, curr_depth, intrinsics, latest_w2c, curr_w2c, frustum=True, latest_varmask=None,
                     curr_varmask=None, method="sum", threshold=0.02, return_pairs=False):
    """The reference's `choose_metric` of a base frame's tracking iterations: the current frame's points (curr_depth at
    curr_w2c) are paired with their nearest neighbours within `threshold` among the latest (overlapping) frame's points
    (latest_depth at latest_w2c), both restricted to the other view's frustum, and the point-to-plane distances along the
    latest frame's depth normals are reduced ('sum' of squares, 'max', or the mean of the 100 largest: `p2p_method`).
    Everything runs in three kernels of libvtgs_cuda.so on the depth maps' device (CUDA only); the reference goes through
    kornia, numpy and an Open3D KD-tree on the CPU.  depth maps: [1,H,W] or [H,W] float32; -> 0-dim tensor."""
    import ctypes as C
    from . import _lib
    from .rasterizer import _ptr, _require_cuda, _stream_ptr
    d0 = latest_depth.reshape(latest_depth.shape[-2:]).contiguous().float()
    d1 = curr_depth.reshape(curr_depth.shape[-2:]).contiguous().float()
    _require_cuda("latest_depth", d0)
    _require_cuda("curr_depth", d1)
    H, W = d0.shape
    dev = d0.device
    K = torch.as_tensor(intrinsics).detach().cpu()
    intr = (C.c_float * 4)(float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
    w0, w1 = torch.as_tensor(latest_w2c).detach().double().cpu(), torch.as_tensor(curr_w2c).detach().double().cpu()
    c0, c1 = torch.linalg.inv(w0), torch.linalg.inv(w1)
    P = H * W
    f32 = dict(dtype=torch.float32, device=dev)
    pts0, nrm0, pts1 = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32), torch.empty((P, 3), **f32)
    ok0, ok1 = torch.empty(P, dtype=torch.uint8, device=dev), torch.empty(P, dtype=torch.uint8, device=dev)
    u8 = lambda m: None if m is None else m.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
    m0, m1 = u8(latest_varmask), u8(curr_varmask)
    table = torch.empty(1 << max(16, (2 * P - 1).bit_length()), dtype=torch.int32, device=dev)
    nxt = torch.empty(P, dtype=torch.int32, device=dev)
    dist = torch.empty(P, **f32)
    idx = torch.empty(P, dtype=torch.int32, device=dev) if return_pairs else None
    L, s = _lib.lib(), _stream_ptr(dev)
    with torch.cuda.device(dev):
        _lib.check(L.vtgs_p2p_prepare(W, H, C.byref(intr), C.byref(_rows12(c0)), C.byref(_rows12(w1)) if frustum else None,
                                      _ptr(d0), _ptr(m0), _ptr(pts0), _ptr(nrm0), _ptr(ok0), s))
        _lib.check(L.vtgs_p2p_prepare(W, H, C.byref(intr), C.byref(_rows12(c1)), C.byref(_rows12(w0)) if frustum else None,
                                      _ptr(d1), _ptr(m1), _ptr(pts1), None, _ptr(ok1), s))
        _lib.check(L.vtgs_p2p_match(P, _ptr(pts0), _ptr(nrm0), _ptr(ok0), P, _ptr(pts1), _ptr(ok1), float(threshold),
                                    _ptr(table), table.numel(), _ptr(nxt), _ptr(dist), _ptr(idx), s))
    d = torch.nan_to_num(dist, nan=0.0)
    if method == "sum":
        out = (d * d).sum()
    elif method == "max":
        out = d.abs().max()
    elif method == "max100":
        out = d.abs()[~torch.isnan(dist)].topk(100)[0].mean()
    else:
        raise ValueError(f"unknown p2p_method {method!r}")
    return (out, dist, idx) if return_pairs else out
