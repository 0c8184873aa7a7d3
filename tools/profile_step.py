"""Eager tracking iterations of the bench workload for ncu (kernels launched directly, no graph).
    python tools/profile_step.py [--small] [--iters 3]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vtgaussian_slam_b200.fused import TrackingSolver  # noqa: E402
from vtgaussian_slam_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--band", type=int, nargs=2, default=(0, 0), help="tile-row band (as one rank of a sharded run would render)")
a = ap.parse_args()
wl = bench.build_workload("small" if a.small else "c2")
dev = torch.device("cuda:0")
fr, s = wl["frame"], wl["settings"]
settings = GaussianRasterizationSettings(
    image_height=fr["H"], image_width=fr["W"], tanfovx=s["tanfovx"], tanfovy=s["tanfovy"], bg=torch.tensor(s["bg"], device=dev),
    scale_modifier=1.0, viewmatrix=torch.tensor(s["viewmatrix"], device=dev), projmatrix=torch.tensor(s["projmatrix"], device=dev),
    sh_degree=0, campos=torch.tensor(s["campos"], device=dev), prefiltered=False)
params = {k: torch.tensor(v, device=dev) for k, v in wl["params"].items()}
solver = TrackingSolver(settings, params, device=dev, w_im=0.5, w_depth=0.025, sil_thres=0.99, use_graph=False, tile_rows=tuple(a.band))
solver.set_frame(torch.tensor(fr["im"]), torch.tensor(fr["depth"]), wl["q"], wl["t"])
for _ in range(a.iters):
    solver.step()
torch.cuda.synchronize()
print("loss", solver.loss_terms()[0].item())
