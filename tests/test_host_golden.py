"""CPU: the host-side mirror (vtgaussian_slam_b200.slam_ops) and the oracle's spec'd front end
against golden vectors produced by the REFERENCE's own functions
(tests/golden/make_golden.py imports /root/reference/utils/*.py)."""
import os

import numpy as np
import pytest
import torch

import oracle
from vtgaussian_slam_b200 import slam_ops, synthetic

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_golden.npz"))


def _params(tag):
    return {k: torch.tensor(G[f"{tag}.params.{k}"]) for k in
            ("means3D", "rgb_colors", "unnorm_rotations", "logit_opacities", "log_scales", "cam_unnorm_rots", "cam_trans")}


def test_build_rotation():
    R = slam_ops.build_rotation(torch.tensor(G["build_rotation.q"]))
    np.testing.assert_allclose(R.numpy(), G["build_rotation.R"], atol=1e-6)


@pytest.mark.parametrize("tag", ["iso", "aniso"])
def test_transform_and_rendervars(tag):
    p = _params(tag)
    tg = slam_ops.transform_to_frame(p, 3, gaussians_grad=True, camera_grad=True)
    np.testing.assert_allclose(tg["means3D"].numpy(), G[f"{tag}.ttf.means3D"], atol=2e-6)
    np.testing.assert_allclose(tg["unnorm_rotations"].numpy(), G[f"{tag}.ttf.unnorm_rotations"], atol=2e-6)
    rv = slam_ops.transformed_params2rendervar(p, tg)
    for k in ("means3D", "colors_precomp", "rotations", "opacities", "scales", "means2D"):
        np.testing.assert_allclose(rv[k].detach().numpy(), G[f"{tag}.rendervar.{k}"], atol=2e-6, err_msg=k)
    assert rv["means2D"].requires_grad and not rv["means2D"].is_leaf
    w2c = torch.tensor(G[f"{tag}.w2c"])
    np.testing.assert_allclose(slam_ops.get_depth_and_silhouette(tg["means3D"], w2c).numpy(), G[f"{tag}.depth_sil"], atol=2e-6)
    dv = slam_ops.transformed_params2depthplussilhouette(p, w2c, tg)
    np.testing.assert_allclose(dv["colors_precomp"].numpy(), G[f"{tag}.dsvar.colors_precomp"], atol=2e-6)


@pytest.mark.parametrize("tag", ["iso", "aniso"])
def test_oracle_frontend_matches_reference_builders(tag):
    """oracle.frontend (spec'd exp / sigmoid / normalise, shared with the fused CUDA kernel)."""
    p = {k: G[f"{tag}.params.{k}"] for k in ("means3D", "rgb_colors", "unnorm_rotations", "logit_opacities", "log_scales")}
    q = G[f"{tag}.params.cam_unnorm_rots"][0, :, 3]
    t = G[f"{tag}.params.cam_trans"][0, :, 3]
    w2c = G[f"{tag}.w2c"]
    m, s, r, o, c6 = oracle.frontend(p["means3D"], p["rgb_colors"], p["unnorm_rotations"], p["logit_opacities"],
                                     p["log_scales"], q, t, depth_row=w2c[2])
    np.testing.assert_allclose(m, G[f"{tag}.rendervar.means3D"], atol=3e-6)
    np.testing.assert_allclose(s, G[f"{tag}.rendervar.scales"], rtol=3e-7 * 4)
    np.testing.assert_allclose(o, G[f"{tag}.rendervar.opacities"][:, 0], rtol=1e-6)
    np.testing.assert_allclose(r, G[f"{tag}.rendervar.rotations"], atol=1e-6)
    np.testing.assert_allclose(c6[:, :3], G[f"{tag}.rendervar.colors_precomp"])
    np.testing.assert_allclose(c6[:, 3:], G[f"{tag}.dsvar.colors_precomp"], rtol=2e-6, atol=2e-6)


def test_losses_and_ssim():
    a, b = torch.tensor(G["ssim.a"]), torch.tensor(G["ssim.b"])
    assert abs(slam_ops.calc_ssim(a, b).item() - float(G["ssim.value"])) < 1e-6
    assert abs(slam_ops.calc_ssim(a, a * 0.9 + 0.05).item() - float(G["ssim.close"])) < 1e-6
    assert abs(slam_ops.l1_loss_v1(a, b).item() - float(G["l1.value"])) < 1e-7
    assert abs(slam_ops.l1_loss_v1_mask(a, b, torch.tensor(G["l1.mask"])).item() - float(G["l1.masked"])) < 1e-7
    qm = slam_ops.quat_mult(torch.tensor(G["quat_mult.q1"]), torch.tensor(G["quat_mult.q2"]))
    np.testing.assert_allclose(qm.numpy(), G["quat_mult.out"], atol=1e-6)


def test_setup_camera():
    cam = slam_ops.setup_camera(1200, 680, G["cam.K"], G["cam.w2c"], device="cpu")
    np.testing.assert_allclose(cam.viewmatrix.numpy(), G["cam.viewmatrix"], atol=1e-7)
    np.testing.assert_allclose(cam.projmatrix.numpy(), G["cam.projmatrix"], atol=1e-7)
    np.testing.assert_allclose(cam.campos.numpy(), G["cam.campos"], atol=1e-7)
    sc = G["cam.scalars"]
    assert (cam.image_height, cam.image_width, cam.sh_degree) == (int(sc[0]), int(sc[1]), int(sc[5]))
    assert abs(cam.tanfovx - sc[2]) < 1e-12 and abs(cam.tanfovy - sc[3]) < 1e-12 and cam.prefiltered is False
    # the numpy twin used by the synthetic scenes agrees too
    s = synthetic.setup_camera(1200, 680, G["cam.K"], G["cam.w2c"])
    np.testing.assert_allclose(s["viewmatrix"], G["cam.viewmatrix"], atol=1e-7)
    np.testing.assert_allclose(s["projmatrix"], G["cam.projmatrix"], atol=1e-6)


def test_initialize_optimizer_groups():
    p = {k: torch.nn.Parameter(v) for k, v in _params("iso").items()}
    lrs = dict(means3D=0.0, rgb_colors=0.0025, unnorm_rotations=0.0, logit_opacities=0.05, log_scales=0.005,
               cam_unnorm_rots=1e-8, cam_trans=1e-7)
    opt = slam_ops.initialize_optimizer(p, lrs, tracking=False)
    assert [g["name"] for g in opt.param_groups] == list(p.keys())
    assert opt.defaults["eps"] == 1e-15 and opt.param_groups[1]["lr"] == 0.0025
    assert slam_ops.initialize_optimizer(p, lrs, tracking=True).defaults["eps"] == 1e-8


def test_tracking_loss_host_logic_on_oracle_render():
    """get_loss's mask / loss arithmetic (pure torch part) on a CPU-oracle render: sums, not means,
    masked by depth>0 & silhouette>thr (reference :513-605)."""
    fr = synthetic.make_frame("replica", 120, 68, seed=0)
    p = synthetic.view_tied_gaussians(fr)
    m, s, r, o, c6 = oracle.frontend(p["means3D"], p["rgb_colors"], p["unnorm_rotations"], p["logit_opacities"],
                                     p["log_scales"], [1, 0, 0, 0], [0.01, 0, 0])
    from helpers import oracle_camera
    cam, _ = oracle_camera(fr["W"], fr["H"], fr["K"])
    out = oracle.Oracle().forward(cam, m, s, r, o, c6)
    img = torch.tensor(out["color"])
    data = dict(im=torch.tensor(fr["im"]), depth=torch.tensor(fr["depth"]))
    loss, wl = slam_ops._masks_and_losses(img[:3], img[3:], data, dict(im=0.5, depth=0.025), True, 0.99, True, False, True,
                                          None, "tum", None, None, None, None, None)
    mask = (data["depth"] > 0) & (img[4:5] > 0.99)
    want_d = (data["depth"] - img[3:4]).abs()[mask].sum()
    want_i = (data["im"] - img[:3]).abs()[mask.expand(3, -1, -1)].sum()
    assert torch.allclose(wl["depth"], 0.025 * want_d) and torch.allclose(wl["im"], 0.5 * want_i)
    assert torch.allclose(loss, 0.5 * want_i + 0.025 * want_d)
