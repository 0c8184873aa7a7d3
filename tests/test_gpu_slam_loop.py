"""-m gpu: the compact tracking + mapping loop (BASELINE config 4 shape, reduced) recovers a synthetic trajectory."""
import numpy as np
import pytest
import torch

from vtgaussian_slam_b200 import synthetic
from vtgaussian_slam_b200.slam_loop import LoopConfig, ViewTiedSLAM, ate_rmse, section_from_frame

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_section_from_frame_matches_the_numpy_construction():
    fr = synthetic.make_frame("tum_fr1", 160, 120, seed=0)
    ref = synthetic.view_tied_gaussians(fr, color_noise=0.0)
    p = section_from_frame(torch.as_tensor(fr["im"]).to(DEV), torch.as_tensor(fr["depth"]).to(DEV), fr["K"], np.eye(4), DEV)
    assert p["means3D"].shape == ref["means3D"].shape
    # the numpy generator ray-casts the analytic scene in fp64, this one reads the fp32 depth image
    assert np.allclose(p["means3D"].cpu().numpy(), ref["means3D"], rtol=2e-6, atol=2e-6)
    assert np.allclose(p["log_scales"].cpu().numpy(), ref["log_scales"], atol=2e-6)
    assert float(p["logit_opacities"].abs().max()) == 0.0 and float((p["unnorm_rotations"][:, 0] - 1).abs().max()) == 0.0


@pytest.mark.parametrize("use_graph", [True, False])
def test_loop_tracks_a_synthetic_sequence(use_graph):
    n = 13
    W, H, K = synthetic.intrinsics("tum_fr1", 320, 240)
    poses = synthetic.trajectory(n, step_m=0.01, step_deg=0.3)
    cfg = LoopConfig(track_iters=60, map_iters=10, baseframe_every=6, map_every=3, use_graph=use_graph)
    slam = ViewTiedSLAM(W, H, K, cfg, device=DEV)
    for i in range(n):
        slam.process(synthetic.make_frame("tum_fr1", 320, 240, seed=i, c2w=poses[i]))
    est = np.stack([np.linalg.inv(m) for m in slam.w2c])
    err = ate_rmse(est, poses)
    still = ate_rmse(np.tile(np.eye(4), (n, 1, 1)), poses)
    assert len(slam.sections) == 3 and slam.stats["track_iters"] == (n - 1) * 61
    assert err < 0.2 * still and err < 0.01, (err, still)
