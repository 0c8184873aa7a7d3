"""Generates tests/golden/get_loss_golden.npz by EXECUTING THE REFERENCE's own `get_loss`
(/root/reference/src/vtgaussian_slam.py:407-689, together with its `get_vis_mask`, :376-404) --
this container only; /root/reference does not exist on the GPU box.

src/vtgaussian_slam.py cannot be imported (rasteriser, open3d, matplotlib ... are absent), so the
two function definitions are cut out of its source text with `ast` and exec'd, unmodified, in a
namespace that holds what they reference:

    transform_to_frame, transformed_params2rendervar, transformed_params2depthplussilhouette,
    l1_loss_v1, l1_loss_v1_mask            imported from the reference's utils/slam_helpers.py
    calc_ssim                              imported from the reference's utils/slam_external.py
    Renderer                               a CPU rasteriser with autograd whose arithmetic is this
                                           repository's oracle (test infrastructure; the real
                                           rasteriser is an absent pip dependency, so the render
                                           itself stays "parity unpinned" -- what these vectors pin
                                           is everything get_loss does AROUND the two renders:
                                           masks, threshold ladder, median, visibility masks,
                                           loss terms, weights, seen / max_2D_radius, and the
                                           autograd chain down to the parameters and the pose)

`.cuda()` is neutralised (no GPU here).  Cases: tracking x {replica iteration 0 (threshold
ladder), replica iteration > 0, tum (far-depth filter + one overlap keyframe), scannetpp (outlier
median + three overlap keyframes)} and mapping x {plain, do_ba, additional_mask}.

    python tests/golden/make_get_loss_golden.py
"""
import ast
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

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

import oracle                               # noqa: E402
from vtgaussian_slam_b200 import synthetic  # noqa: E402
from vtgaussian_slam_b200.slam_loop import quat_from_matrix  # noqa: E402


from oracle_renderer import OracleRenderer as Renderer  # noqa: E402  (tests/oracle_renderer.py)


def reference_functions():
    """get_vis_mask and get_loss exactly as written in the reference's source file."""
    path = os.path.join(REF, "src", "vtgaussian_slam.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = dict(torch=torch, F=F, time=time, os=os, np=np, plt=None, Renderer=Renderer,
              transform_to_frame=RH.transform_to_frame, transformed_params2rendervar=RH.transformed_params2rendervar,
              transformed_params2depthplussilhouette=RH.transformed_params2depthplussilhouette,
              l1_loss_v1=RH.l1_loss_v1, l1_loss_v1_mask=RH.l1_loss_v1_mask, calc_ssim=RE.calc_ssim)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_vis_mask", "get_loss"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["get_loss"], ns["get_vis_mask"]


W, H = 112, 80


def build_scene(seed=0):
    """A small view-tied section (one Gaussian per pixel + edge-densified ones, 'trained' opacities) seen from a
    slightly perturbed pose, with three earlier keyframes for the visibility masks."""
    poses = synthetic.trajectory(4, step_m=0.04, step_deg=1.5, seed=5)
    frames = [synthetic.make_frame("replica", W, H, seed=seed + k, c2w=poses[k]) for k in range(4)]
    fr = frames[3]
    p = synthetic.section_gaussians(fr, c2w=poses[3], opacity="trained", seed=2)
    rng = np.random.default_rng(7)
    p["rgb_colors"] = np.clip(p["rgb_colors"] + rng.normal(0, 0.05, p["rgb_colors"].shape), 0, 1).astype(np.float32)
    w2c = np.linalg.inv(poses[3])
    # tracking start: the true pose with a small perturbation (un-normalised quaternion, ~0.4 deg / ~1 cm)
    q = (1.3 * quat_from_matrix(w2c[:3, :3]) + rng.normal(0, 0.004, 4)).astype(np.float32)
    t = (w2c[:3, 3] + rng.normal(0, 0.01, 3)).astype(np.float32)
    gt_depth = fr["depth"].copy()
    gt_depth[0, :6, :9] = 0.0                      # invalid-depth pixels (depth > 0 mask)
    gt_depth[0, 40:44, 60:70] *= 1.8               # outliers for the median mask
    return dict(frames=frames, poses=poses, params=p, q=q, t=t, im=fr["im"], depth=gt_depth, K=fr["K"])


def run_case(get_loss, sc, name, out, **kw):
    s = synthetic.setup_camera(W, H, sc["K"], np.eye(4))
    cam = oracle.make_camera(W, H, s["tanfovx"], s["tanfovy"], s["viewmatrix"], s["projmatrix"])
    T = 3
    params = {k: torch.tensor(v).clone().requires_grad_(True) for k, v in sc["params"].items()}
    cq = np.tile(np.array([1, 0, 0, 0], np.float32)[None, :, None], (1, 1, T)).copy()
    ct = np.zeros((1, 3, T), np.float32)
    cq[0, :, 2], ct[0, :, 2] = sc["q"], sc["t"]
    params["cam_unnorm_rots"] = torch.tensor(cq).requires_grad_(True)
    params["cam_trans"] = torch.tensor(ct).requires_grad_(True)
    N = params["means3D"].shape[0]
    variables = dict(max_2D_radius=torch.zeros(N), means2D_gradient_accum=torch.zeros(N), denom=torch.zeros(N))
    curr_data = dict(cam=cam, im=torch.tensor(sc["im"]), depth=torch.tensor(sc["depth"]), id=2,
                     intrinsics=torch.tensor(sc["K"], dtype=torch.float32), w2c=torch.eye(4))
    ret = get_loss(params, curr_data, variables, 2, **kw)
    loss, variables, wl = ret[0], ret[1], ret[2]
    loss.backward()
    out[f"{name}.loss"] = np.float32(loss.item())
    for k, v in wl.items():
        out[f"{name}.wl.{k}"] = np.float32(v.item())
    for k in ("cam_unnorm_rots", "cam_trans", "means3D", "rgb_colors", "unnorm_rotations", "logit_opacities", "log_scales"):
        g = params[k].grad
        out[f"{name}.grad.{k}"] = (torch.zeros_like(params[k]) if g is None else g).numpy()
        out[f"{name}.hasgrad.{k}"] = np.bool_(g is not None)
    out[f"{name}.seen"] = variables["seen"].numpy()
    out[f"{name}.max_2D_radius"] = variables["max_2D_radius"].numpy()
    m2g = variables["means2D"].grad
    out[f"{name}.means2D_grad"] = (torch.zeros(N, 3) if m2g is None else m2g).numpy()
    if len(ret) > 3:
        out[f"{name}.mse_ls"] = np.asarray(ret[3], np.float64)
        out[f"{name}.sil_thres_ls"] = np.asarray(ret[4], np.float64)
    print(f"{name:28s} loss {loss.item():.6f}  " + " ".join(f"{k}={v.item():.5f}" for k, v in wl.items()))


def main():
    get_loss, _ = reference_functions()
    sc = build_scene()
    out = {"W": np.int32(W), "H": np.int32(H), "K": np.asarray(sc["K"], np.float64), "q": sc["q"], "t": sc["t"],
           "im": sc["im"], "depth": sc["depth"]}
    for k, v in sc["params"].items():
        out[f"params.{k}"] = v
    # overlap keyframes (w2c + depth) and the current frame's w2c for the visibility masks
    w2c = [np.linalg.inv(p).astype(np.float32) for p in sc["poses"]]
    out["overlap_w2c"] = np.stack(w2c[:3])
    out["overlap_depth"] = np.stack([f["depth"] for f in sc["frames"][:3]])
    out["curr_w2c"] = w2c[3]
    ov = dict(curr_w2c=torch.tensor(w2c[3]), overlap_w2c=torch.tensor(w2c[0]), overlap_gtdepth=torch.tensor(sc["frames"][0]["depth"]),
              overlap_mid_w2c=torch.tensor(w2c[1]), overlap_mid_gtdepth=torch.tensor(sc["frames"][1]["depth"]),
              overlap_last_w2c=torch.tensor(w2c[2]), overlap_last_gtdepth=torch.tensor(sc["frames"][2]["depth"]))
    lw_track = dict(im=0.5, depth=1.0)
    common = dict(loss_weights=lw_track, use_sil_for_loss=True, sil_thres=0.99, use_l1=True)
    # tracking
    run_case(get_loss, sc, "track_replica_it0", out, **common, ignore_outlier_depth_loss=False, tracking=True, tracking_iteration=0,
             dataset_name="replica", presence_sil_mask_mse_ls=[], sil_thres_ls=[])
    run_case(get_loss, sc, "track_replica_it5", out, **common, ignore_outlier_depth_loss=False, tracking=True, tracking_iteration=5,
             dataset_name="replica", presence_sil_mask_mse_ls=[0.01], sil_thres_ls=[0.995])
    run_case(get_loss, sc, "track_tum", out, **common, ignore_outlier_depth_loss=False, tracking=True, tracking_iteration=1,
             dataset_name="tum", far_depth_filter_thres=3.0, vis_mask_thres=0.05, curr_w2c=ov["curr_w2c"], overlap_w2c=ov["overlap_w2c"],
             overlap_gtdepth=ov["overlap_gtdepth"])
    run_case(get_loss, sc, "track_tum_nosil", out, loss_weights=lw_track, use_sil_for_loss=False, sil_thres=0.99, use_l1=True,
             ignore_outlier_depth_loss=False, tracking=True, tracking_iteration=1, dataset_name="tum")
    run_case(get_loss, sc, "track_scannetpp", out, **common, ignore_outlier_depth_loss=True, tracking=True, tracking_iteration=1,
             dataset_name="scannetpp", far_depth_filter_thres=3.0, vis_mask_thres=0.05, **ov)
    # mapping
    lw_map = dict(im=0.5, depth=1.0)
    mcommon = dict(loss_weights=lw_map, use_sil_for_loss=False, sil_thres=0.5, use_l1=True, ignore_outlier_depth_loss=False, mapping=True)
    run_case(get_loss, sc, "map_plain", out, **mcommon, dataset_name="replica")
    run_case(get_loss, sc, "map_ba", out, **mcommon, do_ba=True, dataset_name="tum")
    am = torch.zeros(1, H, W)
    am[0, 20:50, 30:80] = 1.0
    out["additional_mask"] = am.numpy()
    run_case(get_loss, sc, "map_addmask", out, **mcommon, additional_mask=am, dataset_name="replica")
    path = os.path.join(HERE, "get_loss_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
