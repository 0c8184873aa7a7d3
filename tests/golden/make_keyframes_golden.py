"""Golden vectors for vtgaussian_slam_b200.keyframes from the reference's utils/keyframe_selection.py.

    python tests/golden/make_keyframes_golden.py        # build container only (needs /root/reference)

The reference module imports on CPU but calls `.cuda()` at run time: Tensor.cuda is made the identity and
torch.cuda.empty_cache a no-op for the duration of the calls (nothing else is changed)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def scene(seed):
    from vtgaussian_slam_b200 import synthetic
    W, H, K = synthetic.intrinsics("tum_fr1", 160, 120)
    poses = synthetic.trajectory(12, step_m=0.12, step_deg=6.0, seed=seed)
    fr = synthetic.make_frame("tum_fr1", 160, 120, seed=seed, c2w=poses[5])
    depth = fr["depth"].copy()
    depth[0, :10, :30] = 0.0                                  # some invalid depth
    return depth, K.astype(np.float32), poses


if __name__ == "__main__":
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    from utils.keyframe_selection import get_pointcloud, keyframe_selection_overlap
    out = {}
    for case, seed in enumerate((0, 1, 2)):
        depth, K, poses = scene(seed)
        gt_depth = torch.tensor(depth)
        intr = torch.tensor(K)
        w2c = torch.tensor(np.linalg.inv(poses[5]), dtype=torch.float32)
        kfs = [dict(id=i, est_w2c=torch.tensor(np.linalg.inv(poses[i]), dtype=torch.float32)) for i in range(12) if i != 5]
        torch.manual_seed(100 + seed)
        ranked = keyframe_selection_overlap(gt_depth, w2c, intr, [dict(k) for k in kfs], 4, pixels=400, edge_value=8, save_percent=True)
        torch.manual_seed(100 + seed)
        chosen = keyframe_selection_overlap(gt_depth, w2c, intr, [dict(k) for k in kfs], 4, pixels=400, edge_value=8)
        out[f"c{case}.depth"], out[f"c{case}.K"], out[f"c{case}.poses"] = depth, K, poses
        out[f"c{case}.ranked_ids"] = np.array([r["id"] for r in ranked])
        out[f"c{case}.ranked_frac"] = np.array([float(r["percent_inside"]) for r in ranked], np.float32)
        out[f"c{case}.chosen"] = np.array(chosen)
        # get_pointcloud on explicit samples, with repeats and an invalid-depth pixel
        idx = torch.tensor([[40, 50], [40, 50], [5, 5], [100, 20], [60, 150], [100, 20], [77, 33]])
        out[f"c{case}.samples"] = idx.numpy()
        out[f"c{case}.pts"] = get_pointcloud(gt_depth, intr, w2c, idx).numpy()
    np.savez_compressed(os.path.join(HERE, "keyframes_golden.npz"), **out)
    for c in range(3):
        print(out[f"c{c}.chosen"], out[f"c{c}.ranked_frac"][:5], out[f"c{c}.pts"].shape)


def vis_mask_goldens():
    """get_vis_mask (src/vtgaussian_slam.py:376-404) and the nested get_pointcloud_forvismask (:538-556): that module
    cannot be imported (rasteriser, open3d ... absent), so the two function definitions are compiled from its source."""
    import ast
    import torch.nn.functional as F
    from vtgaussian_slam_b200 import synthetic
    tree = ast.parse(open("/root/reference/src/vtgaussian_slam.py").read())
    want = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("get_vis_mask", "get_pointcloud_forvismask") and node.name not in want:
            want[node.name] = node
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=list(want.values()), type_ignores=[]), "vtgaussian_slam.py", "exec"), ns)
    out = {}
    W, H, K = synthetic.intrinsics("tum_fr1", 96, 72)
    poses = synthetic.trajectory(8, step_m=0.25, step_deg=9.0, seed=4)
    cur = synthetic.make_frame("tum_fr1", 96, 72, seed=0, c2w=poses[2])
    gt_depth = torch.tensor(cur["depth"])
    gt_depth[0, 60:, 80:] = 0.0
    intr = torch.tensor(K.astype(np.float32))
    curr_w2c = torch.tensor(np.linalg.inv(poses[2]), dtype=torch.float32)
    idx = torch.stack(torch.where(gt_depth[0] >= 0), dim=1)
    pts = ns["get_pointcloud_forvismask"](gt_depth, intr, curr_w2c, idx)
    out["vis.depth"], out["vis.K"], out["vis.poses"], out["vis.pts"] = gt_depth.numpy(), intr.numpy(), poses, pts.numpy()
    for j, k in enumerate((0, 4, 7)):
        other = synthetic.make_frame("tum_fr1", 96, 72, seed=1 + k, c2w=poses[k])
        od = torch.tensor(other["depth"])
        out[f"vis.other{j}"] = od.numpy()
        out[f"vis.mask{j}"] = ns["get_vis_mask"](torch.tensor(np.linalg.inv(poses[k]), dtype=torch.float32), pts.clone(), intr, od, 0.05, 72, 96).numpy()
    return out


def visbased_goldens():
    """keyframe_selection_overlap_visbased (utils/keyframe_selection.py:121-229), the variant the main loop uses."""
    import contextlib
    import io
    from utils.keyframe_selection import keyframe_selection_overlap_visbased
    from vtgaussian_slam_b200 import synthetic
    out = {}
    W, H, K = synthetic.intrinsics("tum_fr1", 80, 60)
    poses = synthetic.trajectory(10, step_m=0.15, step_deg=7.0, seed=12)
    cur = synthetic.make_frame("tum_fr1", 80, 60, seed=0, c2w=poses[4])
    gt_depth = torch.tensor(cur["depth"])
    gt_depth[0, 50:, :20] = 0.0
    intr = torch.tensor(K.astype(np.float32))
    w2c = torch.tensor(np.linalg.inv(poses[4]), dtype=torch.float32)
    ids = [i for i in range(10) if i != 4]
    depths = [synthetic.make_frame("tum_fr1", 80, 60, seed=30 + i, c2w=poses[i])["depth"] for i in ids]
    mk = lambda: [dict(est_w2c=torch.tensor(np.linalg.inv(poses[i]), dtype=torch.float32), depth=torch.tensor(d)) for i, d in zip(ids, depths)]
    with contextlib.redirect_stdout(io.StringIO()):                # the reference prints its ranking
        ranked = keyframe_selection_overlap_visbased(gt_depth, w2c, intr, mk(), 3, edge_value=6, save_percent=True, kf_depth_thresh=0.02)
        sel, early = keyframe_selection_overlap_visbased(gt_depth, w2c, intr, mk(), 3, edge_value=6, kf_depth_thresh=0.02, earliest_thres=0.3)
        sel2, early2 = keyframe_selection_overlap_visbased(gt_depth, w2c, intr, mk(), 3, edge_value=6, kf_depth_thresh=0.02, earliest_thres=0.99)
    out["vb.depth"], out["vb.K"], out["vb.poses"], out["vb.kf_depths"] = gt_depth.numpy(), intr.numpy(), poses, np.stack(depths)
    out["vb.ranked_ids"] = np.array([r["id"] for r in ranked])
    out["vb.ranked_frac"] = np.array([float(r["percent_inside"]) for r in ranked], np.float32)
    out["vb.sel"], out["vb.early"], out["vb.sel2"], out["vb.early2"] = np.array(sel), np.array(early), np.array(sel2), np.array(early2)
    return out


def topkbase_goldens():
    """keyframe_selection_overlap_visbased_earliest_dynamic_new_topkbase (utils/keyframe_selection.py:581-703), the
    section choice of the main loop (src/vtgaussian_slam.py:1549-1553), on a 24-keyframe orbit: several thresholds,
    topk_base 3 / 2 / None, a short list and a list nothing overlaps with."""
    import contextlib
    import io
    from utils.keyframe_selection import keyframe_selection_overlap_visbased_earliest_dynamic_new_topkbase as ref_fn
    from vtgaussian_slam_b200 import synthetic
    out = {}
    W, H, K = synthetic.intrinsics("tum_fr1", 80, 60)
    poses = synthetic.trajectory(25, step_m=0.10, step_deg=5.0, seed=21)
    cur = synthetic.make_frame("tum_fr1", 80, 60, seed=0, c2w=poses[24])
    gt_depth = torch.tensor(cur["depth"])
    gt_depth[0, :8, 60:] = 0.0
    intr = torch.tensor(K.astype(np.float32))
    w2c = torch.tensor(np.linalg.inv(poses[24]), dtype=torch.float32)
    depths = [synthetic.make_frame("tum_fr1", 80, 60, seed=50 + i, c2w=poses[i])["depth"] for i in range(24)]
    mk = lambda n: [dict(est_w2c=torch.tensor(np.linalg.inv(poses[i]), dtype=torch.float32), depth=torch.tensor(depths[i])) for i in range(n)]
    cfg = dict(baseframe_every=8, overlap_every=2)
    cases = [dict(n=24, earliest_thres=0.5, topk_base=3), dict(n=24, earliest_thres=0.9, topk_base=3),
             dict(n=24, earliest_thres=0.3, topk_base=2), dict(n=24, earliest_thres=0.5, topk_base=None),
             dict(n=6, earliest_thres=0.5, topk_base=3), dict(n=24, earliest_thres=0.999, topk_base=3, lower=0.5),
             dict(n=4, earliest_thres=0.5, topk_base=3, far=True)]
    out["tk.depth"], out["tk.K"], out["tk.poses"], out["tk.kf_depths"] = gt_depth.numpy(), intr.numpy(), poses, np.stack(depths)
    for j, c in enumerate(cases):
        kfs = mk(c["n"])
        if c.get("far"):                                   # keyframes that look away: nothing overlaps
            for kf in kfs:
                kf["est_w2c"] = torch.tensor(np.diag([-1.0, 1.0, -1.0, 1.0]), dtype=torch.float32) @ kf["est_w2c"]
        with contextlib.redirect_stdout(io.StringIO()):
            got = ref_fn(gt_depth, w2c, intr, kfs, 3, cfg, edge_value=6, kf_depth_thresh=0.02, earliest_thres=c["earliest_thres"],
                         lower_earliest_thres_percent=c.get("lower", 0.8), topk_base=c["topk_base"])
        out[f"tk.case{j}"] = np.array(sorted(got))
        out[f"tk.cfg{j}"] = np.array([c["n"], c["earliest_thres"], -1 if c["topk_base"] is None else c["topk_base"], c.get("lower", 0.8), float(bool(c.get("far")))])
    out["tk.ncases"] = np.array(len(cases))
    return out


if __name__ == "__main__":
    g = dict(np.load(os.path.join(HERE, "keyframes_golden.npz")))
    g.update(vis_mask_goldens())
    g.update(visbased_goldens())
    g.update(topkbase_goldens())
    print({k: v.tolist() for k, v in g.items() if k.startswith("tk.case")})
    print({k: v for k, v in g.items() if k in ("vb.sel", "vb.early", "vb.early2", "vb.ranked_frac")})
    np.savez_compressed(os.path.join(HERE, "keyframes_golden.npz"), **g)
    print({k: float(v.mean()) for k, v in g.items() if k.startswith("vis.mask")})
