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
