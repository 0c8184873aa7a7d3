"""CPU: vtgaussian_slam_b200.metrics against the reference's align / evaluate_ate / calc_psnr
(golden vectors: tests/golden/make_metrics_golden.py)."""
import os

import numpy as np
import torch

from vtgaussian_slam_b200 import metrics

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.npz"))


def test_horn_alignment_and_ate_match_the_reference():
    for k in range(3):
        gt, est = G[f"ate{k}.gt"], G[f"ate{k}.est"]
        R, t, err = metrics.align_horn(gt[:, :3, 3].T, est[:, :3, 3].T)
        assert np.allclose(R, G[f"ate{k}.R"], atol=1e-9) and np.allclose(t, G[f"ate{k}.t"], atol=1e-9)
        assert np.allclose(err, G[f"ate{k}.err"], atol=1e-9)
        assert abs(np.linalg.det(R) - 1.0) < 1e-9                                  # a rotation, never a reflection
        v = metrics.ate_after_alignment([torch.tensor(m) for m in gt], [torch.tensor(m) for m in est])
        assert abs(v - float(G[f"ate{k}.value"])) < 1e-7                         # (the reference centres the float32 points in float32)
    assert metrics.ate_after_alignment(list(G["ate1.gt"]), list(G["ate1.gt"])) < 1e-7


def test_psnr_matches_the_reference_and_depth_l1_ignores_invalid_pixels():
    a, b = torch.tensor(G["psnr.a"]), torch.tensor(G["psnr.b"])
    assert np.allclose(metrics.psnr(a, b).numpy(), G["psnr.value"], rtol=0, atol=0)
    d = torch.tensor([[1.0, 2.0], [0.0, 4.0]])
    r = torch.tensor([[1.5, 2.0], [9.0, 3.0]])
    assert abs(metrics.depth_l1(r, d).item() - (0.5 + 0.0 + 1.0) / 3) < 1e-7
