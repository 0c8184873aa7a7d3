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


def _gauss_window(size=11, sigma=1.5, dtype=torch.float32, device="cpu"):
    c = torch.arange(size, dtype=dtype, device=device) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _valid_gauss(x, win):
    """Separable 'valid' (unpadded) Gaussian filtering of x[B,C,H,W] per channel."""
    C = x.shape[1]
    k = win.to(x)
    x = torch.nn.functional.conv2d(x, k.reshape(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)
    return torch.nn.functional.conv2d(x, k.reshape(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)


def ms_ssim(X, Y, data_range=1.0, size_average=True, win_size=11, win_sigma=1.5,
            weights=(0.0448, 0.2856, 0.3001, 0.2363, 0.1333), K=(0.01, 0.03)):
    """Multi-scale SSIM as the reference's evaluation calls it (`pytorch_msssim.ms_ssim(im[None], gt[None],
    data_range=1.0, size_average=True)`, utils/eval_helpers.py:453).  pytorch_msssim (requirements.txt, version not
    pinned) is not installed here; this restates its published algorithm: five scales, unpadded separable 11-tap Gaussian
    (sigma 1.5), contrast-structure terms of the first four scales and the full SSIM of the last, each clamped at 0,
    raised to the standard weights and multiplied; 2x2 average pooling (odd sizes padded) between scales.
    X, Y: [B,C,H,W] with min(H, W) > 160."""
    if min(X.shape[-2:]) <= (win_size - 1) * 2 ** 4:
        raise ValueError("image too small for five scales: the smaller side must exceed 160 pixels")
    win = _gauss_window(win_size, win_sigma, X.dtype, X.device)
    w = torch.tensor(weights, dtype=X.dtype, device=X.device)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mcs = []
    for lvl in range(len(weights)):
        mu1, mu2 = _valid_gauss(X, win), _valid_gauss(Y, win)
        s1 = _valid_gauss(X * X, win) - mu1 * mu1
        s2 = _valid_gauss(Y * Y, win) - mu2 * mu2
        s12 = _valid_gauss(X * Y, win) - mu1 * mu2
        cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
        ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs_map
        if lvl < len(weights) - 1:
            mcs.append(torch.relu(cs_map.flatten(2).mean(-1)))
            pad = [s % 2 for s in X.shape[2:]]
            X = torch.nn.functional.avg_pool2d(X, kernel_size=2, padding=pad)
            Y = torch.nn.functional.avg_pool2d(Y, kernel_size=2, padding=pad)
    last = torch.relu(ssim_map.flatten(2).mean(-1))
    val = torch.prod(torch.stack(mcs + [last], dim=0) ** w.view(-1, 1, 1), dim=0)        # [B,C]
    return val.mean() if size_average else val.mean(1)
