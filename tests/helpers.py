"""Shared test helpers (test infrastructure)."""
import numpy as np

import oracle
from vtgaussian_slam_b200 import synthetic


def oracle_camera(width, height, K, w2c=None, bg=(0, 0, 0), sigma_mult=3.0, tile_rows=(0, 0)):
    w2c = np.eye(4) if w2c is None else w2c
    s = synthetic.setup_camera(width, height, K, w2c, bg=bg)
    cam = oracle.make_camera(width, height, s["tanfovx"], s["tanfovy"], s["viewmatrix"], s["projmatrix"],
                             bg=bg, radius_sigma_mult=sigma_mult, tile_rows=tile_rows)
    return cam, s


def cam_dict(s):
    """settings dict -> the dict tests/torch_ref.render expects."""
    return dict(W=s["image_width"], H=s["image_height"], tanfovx=s["tanfovx"], tanfovy=s["tanfovy"],
                view=s["viewmatrix"].reshape(-1), proj=s["projmatrix"].reshape(-1), bg=[float(b) for b in s["bg"]],
                scale_modifier=s["scale_modifier"])


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), floor)
    return np.abs(a - b).max() / scale


def oracle_chain_grads(p, q, t, g, depth_row=(0.0, 0.0, 1.0, 0.0)):
    """Gradients w.r.t. the reference's parameters from the oracle's rasteriser-input gradients `g` (Oracle.backward of
    the six-plane pass): the reference's own autograd chain through get_depth_and_silhouette, the activations and
    transform_to_frame (slam_ops host mirrors, pinned by tests/golden), evaluated in fp64 on the CPU.
    -> dict: cam_unnorm_rots[4], cam_trans[3], means3D, rgb_colors, unnorm_rotations, logit_opacities, log_scales."""
    import torch
    from vtgaussian_slam_b200 import slam_ops
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    P["cam_unnorm_rots"] = torch.tensor(q, dtype=torch.float64).reshape(1, 4, 1).requires_grad_(True)
    P["cam_trans"] = torch.tensor(t, dtype=torch.float64).reshape(1, 3, 1).requires_grad_(True)
    tg = slam_ops.transform_to_frame(P, 0, gaussians_grad=True, camera_grad=True)
    rv = slam_ops.transformed_params2rendervar(P, tg)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[2] = torch.tensor(depth_row, dtype=torch.float64)
    dsc = slam_ops.get_depth_and_silhouette(tg["means3D"], w2c)
    outs = [rv["means3D"], rv["scales"], rv["rotations"], rv["opacities"][:, 0], rv["colors_precomp"], dsc]
    gouts = [g["means3D"], g["scales"], g["rotations"], g["opacities"], g["colors"][:, :3], g["colors"][:, 3:]]
    torch.autograd.backward(outs, [torch.tensor(x, dtype=torch.float64) for x in gouts])
    out = {k: P[k].grad.numpy() for k in p}
    out["cam_unnorm_rots"] = P["cam_unnorm_rots"].grad.numpy().reshape(4)
    out["cam_trans"] = P["cam_trans"].grad.numpy().reshape(3)
    return out


def tracking_dL(img6, gt_rgb, gt_depth, w_im, w_depth, sil_thres):
    """dL/dplanes [6,H,W] and the loss of the reference's tracking loss (get_loss :513-605, tum-style fixed threshold,
    use_sil_for_loss) on numpy planes -- the comparand of vtgs_loss."""
    gd = gt_depth.reshape(img6.shape[1:])
    mask = (gd > 0) & (img6[4] > sil_thres) & ~np.isnan(img6[3]) & ~np.isnan(img6[5] - img6[3] ** 2)
    e_im = img6[:3] - gt_rgb
    e_d = img6[3] - gd
    dL = np.zeros_like(img6)
    dL[:3] = w_im * np.sign(e_im) * mask
    dL[3] = w_depth * np.sign(e_d) * mask
    loss = w_im * np.abs(e_im)[:, mask].astype(np.float64).sum() + w_depth * np.abs(e_d)[mask].astype(np.float64).sum()
    return dL.astype(np.float32), float(loss)
