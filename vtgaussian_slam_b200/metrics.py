"""Result metrics of the reference's evaluation (the part that needs no renderer):

  * ate_after_alignment   utils/eval_helpers.py:29-83 -- Horn's closed-form rigid alignment of the estimated camera
                          centres onto the ground truth, then the MEAN translational error (the reference reports
                          the mean, not the RMSE, as "ATE RMSE")
  * psnr                  utils/slam_external.py:49-51 -- per-channel-row PSNR of images in [0, 1]
  * depth_l1              utils/eval_helpers.py (eval loop) -- mean |d_render - d_gt| over valid ground-truth depth
numpy / torch only, device-agnostic.
"""
from __future__ import annotations

import numpy as np
import torch


def align_horn(model, data):
    """Rigid (R, t) minimising sum |R model_i + t - data_i|^2 for 3xN point sets; also the per-point error."""
    model, data = np.asarray(model, np.float64), np.asarray(data, np.float64)
    mc, dc = model.mean(1, keepdims=True), data.mean(1, keepdims=True)
    W = (model - mc) @ (data - dc).T                     # sum of outer products
    U, _, Vh = np.linalg.svd(W.T)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1.0
    R = U @ S @ Vh
    t = dc - R @ mc
    err = np.sqrt(((R @ model + t - data) ** 2).sum(0))
    return R, t, err


def ate_after_alignment(gt_c2w, est_c2w):
    """Mean translational error of the camera centres after Horn alignment (reference evaluate_ate)."""
    g = np.stack([np.asarray(torch.as_tensor(m).detach().cpu(), np.float64)[:3, 3] for m in gt_c2w]).T
    e = np.stack([np.asarray(torch.as_tensor(m).detach().cpu(), np.float64)[:3, 3] for m in est_c2w]).T
    return float(align_horn(g, e)[2].mean())


def psnr(img1, img2):
    """20 log10(1 / sqrt(mse)) with the mse taken per leading-dimension row (reference calc_psnr) -> [C, 1]."""
    mse = ((img1 - img2) ** 2).reshape(img1.shape[0], -1).mean(1, keepdim=True)
    return 20 * torch.log10(1.0 / torch.sqrt(mse))


def depth_l1(rendered_depth, gt_depth):
    """Mean absolute depth error over pixels with valid ground truth (gt > 0)."""
    valid = gt_depth > 0
    return (torch.abs(rendered_depth - gt_depth) * valid).sum() / valid.sum().clamp_min(1)
