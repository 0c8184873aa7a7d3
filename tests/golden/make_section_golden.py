"""Golden vectors for the section builders of vtgaussian_slam_b200.slam_loop from the reference's own functions.

    python tests/golden/make_section_golden.py          # build container only (needs /root/reference)

src/vtgaussian_slam.py cannot be imported here (rasteriser, open3d, ... absent): `get_pointcloud` (:76-128) and
`geometric_edge_mask` (:1022-1041) are compiled from their own source text and run unchanged, with Tensor.cuda made
the identity.  The sequence of calls is that of initialize_params_base_timestep (:285-345)."""
import ast
import os
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

if __name__ == "__main__":
    from vtgaussian_slam_b200 import synthetic
    torch.Tensor.cuda = lambda self, *a, **k: self
    tree = ast.parse(open("/root/reference/src/vtgaussian_slam.py").read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_pointcloud", "geometric_edge_mask")]
    ns = {"torch": torch, "np": np, "cv2": cv2}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "vtgaussian_slam.py", "exec"), ns)
    get_pointcloud, geometric_edge_mask = ns["get_pointcloud"], ns["geometric_edge_mask"]

    pose = synthetic.trajectory(9, 0.2, 7.0, seed=6)[6]
    W, H, K = synthetic.intrinsics("tum_fr1", 96, 72)
    W2, H2, K2 = synthetic.intrinsics("tum_fr1", 192, 144)
    fr = synthetic.make_frame("tum_fr1", 96, 72, seed=3, c2w=pose)
    fr2 = synthetic.make_frame("tum_fr1", 192, 144, seed=3, c2w=pose)
    fr["depth"][0, :6, :9] = 0.0
    fr2["depth"][0, 100:, 150:] = 0.0
    color, depth = torch.tensor(fr["im"]), torch.tensor(fr["depth"])
    color2, depth2 = torch.tensor(fr2["im"]), torch.tensor(fr2["depth"])
    w2c = torch.tensor(np.linalg.inv(pose), dtype=torch.float32)
    intr, intr2 = torch.tensor(K.astype(np.float32)), torch.tensor(K2.astype(np.float32))
    color0_np = (color.permute(1, 2, 0) * 255).numpy()                      # what dataset[0] returns: HWC, 0..255, float32
    mask_variation = geometric_edge_mask(color0_np, dilate=True, RGB=True)
    p0, d0 = get_pointcloud(color, depth, intr, w2c, mask=(depth > 0).reshape(-1), compute_mean_sq_dist=True, mean_sq_dist_method="projective")
    mv = cv2.resize(mask_variation, (color2.shape[2], color2.shape[1]), interpolation=cv2.INTER_NEAREST).astype(np.bool_)
    m2 = (depth2 > 0).reshape(-1) & torch.tensor(mv.reshape(-1))
    p1, d1 = get_pointcloud(color2, depth2, intr2, w2c, mask=m2, compute_mean_sq_dist=True, mean_sq_dist_method="projective")
    cld, msd = torch.cat((p0, p1), 0), torch.cat((d0, d1), 0)
    np.savez_compressed(os.path.join(HERE, "section_golden.npz"), pose=pose, im=fr["im"], depth=fr["depth"], im2=fr2["im"], depth2=fr2["depth"],
                        K=K, K2=K2, edge_mask=mask_variation, means3D=cld[:, :3].numpy(), rgb=cld[:, 3:6].numpy(),
                        log_scales=torch.log(torch.sqrt(msd)).numpy(), n_base=np.int64(p0.shape[0]))
    print(cld.shape, int(p0.shape[0]), float(mask_variation.mean()) / 255)
