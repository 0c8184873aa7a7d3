"""CPU: the MS-SSIM restatement (pytorch_msssim is not installed; see metrics.ms_ssim) -- structural properties."""
import pytest
import torch

from vtgaussian_slam_b200 import metrics


def test_ms_ssim_properties():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 176, 200, generator=g)
    assert abs(float(metrics.ms_ssim(x, x)) - 1.0) < 1e-6
    noisy = (x + 0.05 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    noisier = (x + 0.2 * torch.randn(x.shape, generator=g)).clamp(0, 1)
    a, b = float(metrics.ms_ssim(x, noisy)), float(metrics.ms_ssim(x, noisier))
    assert 0.0 < b < a < 1.0
    assert abs(float(metrics.ms_ssim(noisy, x)) - a) < 1e-6                      # symmetric
    per_image = metrics.ms_ssim(x, noisy, size_average=False)
    assert per_image.shape == (2,) and abs(float(per_image.mean()) - a) < 1e-6
    # a constant image against itself: all variances vanish, every term is c / c = 1
    c = torch.full((1, 3, 170, 170), 0.25)
    assert abs(float(metrics.ms_ssim(c, c)) - 1.0) < 1e-6
    with pytest.raises(ValueError):
        metrics.ms_ssim(x[..., :160, :], x[..., :160, :])


def test_first_scale_is_the_plain_unpadded_ssim():
    g = torch.Generator().manual_seed(1)
    x, y = torch.rand(1, 1, 40, 48, generator=g), torch.rand(1, 1, 40, 48, generator=g)
    # one scale with weight 1: the mean of the SSIM map over the valid (unpadded) window positions, computed densely here
    got = float(metrics.ms_ssim(torch.nn.functional.interpolate(x, scale_factor=8), torch.nn.functional.interpolate(y, scale_factor=8),
                                weights=(1.0,)))
    X, Y = torch.nn.functional.interpolate(x, scale_factor=8)[0, 0], torch.nn.functional.interpolate(y, scale_factor=8)[0, 0]
    w = metrics._gauss_window()
    w2 = w[:, None] * w[None]
    pa, pb = X.unfold(0, 11, 1).unfold(1, 11, 1), Y.unfold(0, 11, 1).unfold(1, 11, 1)
    mu1, mu2 = (pa * w2).sum((-1, -2)), (pb * w2).sum((-1, -2))
    s1, s2 = (pa * pa * w2).sum((-1, -2)) - mu1 ** 2, (pb * pb * w2).sum((-1, -2)) - mu2 ** 2
    s12 = (pa * pb * w2).sum((-1, -2)) - mu1 * mu2
    ssim = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 ** 2 + mu2 ** 2 + 1e-4) * (s1 + s2 + 9e-4))
    assert abs(got - float(ssim.mean())) < 1e-5
