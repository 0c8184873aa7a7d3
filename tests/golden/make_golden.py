"""Generates tests/golden/host_golden.npz by IMPORTING THE REFERENCE's own pure-PyTorch host
functions (this container only; /root/reference does not exist on the GPU box) with .cuda()
neutralised, and recording their outputs on seeded inputs:

    utils/slam_external.py : build_rotation, calc_ssim
    utils/slam_helpers.py  : transform_to_frame, transformed_params2rendervar,
                             get_depth_and_silhouette, transformed_params2depthplussilhouette,
                             l1_loss_v1, l1_loss_v1_mask, quat_mult
    utils/recon_helpers.py : setup_camera   (its `diff_gaussian_rasterization` import resolves
                                             to this repo's drop-in shim)

These pin the host-side restatements (vtgaussian_slam_b200.slam_ops, oracle.frontend) to the
reference.  The rasteriser itself cannot be pinned this way (absent pip dependency).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

# ---- neutralise the reference's hard-coded CUDA placement -------------------------------------
torch.Tensor.cuda = lambda self, *a, **k: self
for _name in ("zeros", "ones", "tensor", "zeros_like", "ones_like", "eye", "arange"):
    _orig = getattr(torch, _name)

    def _wrap(*a, _orig=_orig, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k.pop("device")
        return _orig(*a, **k)
    setattr(torch, _name, _wrap)

from utils import slam_external as RE      # noqa: E402
from utils import slam_helpers as RH       # noqa: E402
from utils import recon_helpers as RR      # noqa: E402


def main():
    g = torch.Generator().manual_seed(1234)
    out = {}
    N, T = 257, 5
    q = torch.randn(N, 4, generator=g)
    out["build_rotation.q"] = q.numpy()
    out["build_rotation.R"] = RE.build_rotation(q).numpy()

    for tag, sdim in (("iso", 1), ("aniso", 3)):
        params = {
            "means3D": torch.randn(N, 3, generator=g) * 2 + torch.tensor([0.0, 0.0, 3.0]),
            "rgb_colors": torch.rand(N, 3, generator=g),
            "unnorm_rotations": torch.randn(N, 4, generator=g),
            "logit_opacities": torch.randn(N, 1, generator=g) * 2,
            "log_scales": torch.randn(N, sdim, generator=g) * 0.5 - 3.0,
            "cam_unnorm_rots": torch.randn(1, 4, T, generator=g) * 0.1 + torch.tensor([1.0, 0, 0, 0]).reshape(1, 4, 1),
            "cam_trans": torch.randn(1, 3, T, generator=g) * 0.1,
        }
        for k, v in params.items():
            out[f"{tag}.params.{k}"] = v.numpy()
        t_idx = 3
        tg = RH.transform_to_frame(params, t_idx, gaussians_grad=True, camera_grad=True)
        out[f"{tag}.ttf.means3D"] = tg["means3D"].numpy()
        out[f"{tag}.ttf.unnorm_rotations"] = tg["unnorm_rotations"].numpy()
        rv = RH.transformed_params2rendervar(params, tg)
        for k in ("means3D", "colors_precomp", "rotations", "opacities", "scales", "means2D"):
            out[f"{tag}.rendervar.{k}"] = rv[k].detach().numpy()
        w2c = torch.eye(4)
        w2c[:3, :3] = RE.build_rotation(torch.tensor([[0.98, 0.05, -0.1, 0.02]]))[0]
        w2c[:3, 3] = torch.tensor([0.1, -0.2, 0.05])
        out[f"{tag}.w2c"] = w2c.numpy()
        ds = RH.get_depth_and_silhouette(tg["means3D"], w2c)
        out[f"{tag}.depth_sil"] = ds.numpy()
        dv = RH.transformed_params2depthplussilhouette(params, w2c, tg)
        out[f"{tag}.dsvar.colors_precomp"] = dv["colors_precomp"].numpy()

    a, b = torch.rand(3, 40, 56, generator=g), torch.rand(3, 40, 56, generator=g)
    out["ssim.a"], out["ssim.b"] = a.numpy(), b.numpy()
    out["ssim.value"] = np.array(RE.calc_ssim(a, b).item())
    out["ssim.close"] = np.array(RE.calc_ssim(a, a * 0.9 + 0.05).item())
    out["l1.value"] = np.array(RH.l1_loss_v1(a, b).item())
    m = (torch.rand(3, 40, 56, generator=g) > 0.5).float()
    out["l1.mask"] = m.numpy()
    out["l1.masked"] = np.array(RH.l1_loss_v1_mask(a, b, m).item())
    q1, q2 = torch.randn(9, 4, generator=g), torch.randn(9, 4, generator=g)
    out["quat_mult.q1"], out["quat_mult.q2"] = q1.numpy(), q2.numpy()
    out["quat_mult.out"] = RH.quat_mult(q1, q2).numpy()

    K = np.array([[600.0, 0, 599.5], [0, 600.0, 339.5], [0, 0, 1]])
    w2c = np.eye(4)
    w2c[:3, 3] = [0.1, 0.0, -0.2]
    cam = RR.setup_camera(1200, 680, K, w2c)
    out["cam.K"], out["cam.w2c"] = K, w2c
    out["cam.viewmatrix"] = cam.viewmatrix.numpy()
    out["cam.projmatrix"] = cam.projmatrix.numpy()
    out["cam.campos"] = cam.campos.numpy()
    out["cam.scalars"] = np.array([cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, cam.scale_modifier, cam.sh_degree])
    out["cam.bg"] = cam.bg.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
