"""CPU: vtgaussian_slam_b200.keyframes against the reference's keyframe_selection_overlap / get_pointcloud
(golden vectors: tests/golden/make_keyframes_golden.py)."""
import os

import numpy as np
import torch

from vtgaussian_slam_b200 import keyframes

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keyframes_golden.npz"))


def _case(c):
    depth, K, poses = torch.tensor(G[f"c{c}.depth"]), torch.tensor(G[f"c{c}.K"]), G[f"c{c}.poses"]
    w2c = torch.tensor(np.linalg.inv(poses[5]), dtype=torch.float32)
    kfs = [dict(id=i, est_w2c=torch.tensor(np.linalg.inv(poses[i]), dtype=torch.float32)) for i in range(12) if i != 5]
    return depth, K, w2c, kfs


def test_selection_and_ranking_match_the_reference():
    for c, seed in enumerate((0, 1, 2)):
        depth, K, w2c, kfs = _case(c)
        torch.manual_seed(100 + seed)
        ranked = keyframes.keyframe_selection_overlap(depth, w2c, K, kfs, 4, pixels=400, edge_value=8, save_percent=True)
        assert [r["id"] for r in ranked] == list(G[f"c{c}.ranked_ids"])                 # including the order of ties
        assert np.allclose([float(r["percent_inside"]) for r in ranked], G[f"c{c}.ranked_frac"], atol=1e-7)
        torch.manual_seed(100 + seed)
        assert keyframes.keyframe_selection_overlap(depth, w2c, K, kfs, 4, pixels=400, edge_value=8) == list(G[f"c{c}.chosen"])
    assert keyframes.keyframe_selection_overlap(depth, w2c, K, [], 4) == []


def test_backprojection_drops_coincident_points_like_the_reference():
    for c in range(3):
        depth, K, w2c, _ = _case(c)
        pts = keyframes.backproject_samples(depth, K, w2c, torch.tensor(G[f"c{c}.samples"]))
        assert pts.shape == G[f"c{c}.pts"].shape                     # repeated samples removed each other, so did the zero-depth pixel
        assert np.allclose(pts.numpy(), G[f"c{c}.pts"], atol=1e-6)


def test_a_generator_makes_the_sampling_reproducible_without_touching_global_state():
    depth, K, w2c, kfs = _case(0)
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    a = keyframes.keyframe_selection_overlap(depth, w2c, K, kfs, 3, pixels=300, edge_value=8, generator=g1)
    b = keyframes.keyframe_selection_overlap(depth, w2c, K, kfs, 3, pixels=300, edge_value=8, generator=g2)
    assert a == b and len(a) == 3
    far = [dict(est_w2c=torch.tensor(np.diag([-1.0, 1.0, -1.0, 1.0]), dtype=torch.float32))]      # looks the other way
    assert keyframes.keyframe_selection_overlap(depth, w2c, K, far, 3, pixels=300, edge_value=8) == []


def test_visibility_mask_matches_the_reference():
    depth, K, poses = torch.tensor(G["vis.depth"]), torch.tensor(G["vis.K"]), G["vis.poses"]
    curr_w2c = torch.tensor(np.linalg.inv(poses[2]), dtype=torch.float32)
    pts = keyframes.frame_points(depth, K, curr_w2c)
    assert np.allclose(pts.numpy(), G["vis.pts"], atol=1e-6)
    overlaps = []
    for j, k in enumerate((0, 4, 7)):
        w2c = torch.tensor(np.linalg.inv(poses[k]), dtype=torch.float32)
        od = torch.tensor(G[f"vis.other{j}"])
        m = keyframes.get_vis_mask(w2c, pts, K, od, 0.05, 72, 96)
        ref = G[f"vis.mask{j}"]
        assert m.shape == ref.shape and m.dtype == torch.bool
        assert (m.numpy() != ref).mean() < 2e-4          # knife-edge pixels of the float comparison at most
        overlaps.append((w2c, od))
    both = keyframes.tracking_vis_mask(depth, K, curr_w2c, overlaps, 0.05)
    assert both.shape == (1, 72, 96)
    assert ((both[0].numpy()) != (G["vis.mask0"] | G["vis.mask1"] | G["vis.mask2"])).mean() < 5e-4
    assert 0.3 < float(both.float().mean()) < 1.0


def test_visibility_based_selection_matches_the_reference():
    depth, K, poses = torch.tensor(G["vb.depth"]), torch.tensor(G["vb.K"]), G["vb.poses"]
    w2c = torch.tensor(np.linalg.inv(poses[4]), dtype=torch.float32)
    ids = [i for i in range(10) if i != 4]
    kfs = [dict(est_w2c=torch.tensor(np.linalg.inv(poses[i]), dtype=torch.float32), depth=torch.tensor(d)) for i, d in zip(ids, G["vb.kf_depths"])]
    ranked = keyframes.keyframe_selection_overlap_visbased(depth, w2c, K, kfs, 3, edge_value=6, save_percent=True, kf_depth_thresh=0.02)
    assert [r["id"] for r in ranked] == list(G["vb.ranked_ids"])
    assert np.allclose([float(r["percent_inside"]) for r in ranked], G["vb.ranked_frac"], atol=3e-4)      # knife-edge pixels of the depth test
    sel, early = keyframes.keyframe_selection_overlap_visbased(depth, w2c, K, kfs, 3, edge_value=6, kf_depth_thresh=0.02, earliest_thres=0.3)
    assert sel == list(G["vb.sel"]) and early == list(G["vb.early"])
    sel2, early2 = keyframes.keyframe_selection_overlap_visbased(depth, w2c, K, kfs, 3, edge_value=6, kf_depth_thresh=0.02, earliest_thres=0.99)
    assert sel2 == list(G["vb.sel2"]) and early2 == list(G["vb.early2"]) == sel2                          # nothing above the threshold
    # chunking over keyframes does not change the result
    pts = keyframes.backproject_samples(depth, K, w2c, torch.stack(torch.where(depth[0] > 0), dim=1))
    stack = torch.stack([k["est_w2c"] for k in kfs])
    deps = torch.stack([k["depth"] for k in kfs])
    a = keyframes.overlap_fractions_visible(pts, K, stack, deps, 80, 60, 6, 0.02, chunk=2)
    b = keyframes.overlap_fractions_visible(pts, K, stack, deps, 80, 60, 6, 0.02, chunk=64)
    assert torch.equal(a, b)


def test_section_choice_of_the_main_loop_matches_the_reference():
    """keyframe_selection_overlap_visbased_earliest_dynamic_new_topkbase: dynamic threshold + earliest top-k sections."""
    depth, K, poses = torch.tensor(G["tk.depth"]), torch.tensor(G["tk.K"]), G["tk.poses"]
    w2c = torch.tensor(np.linalg.inv(poses[24]), dtype=torch.float32)
    cfg = dict(baseframe_every=8, overlap_every=2)
    for j in range(int(G["tk.ncases"])):
        n, thres, topk, lower, far = G[f"tk.cfg{j}"]
        kfs = [dict(est_w2c=torch.tensor(np.linalg.inv(poses[i]), dtype=torch.float32), depth=torch.tensor(G["tk.kf_depths"][i])) for i in range(int(n))]
        if far:
            for kf in kfs:
                kf["est_w2c"] = torch.tensor(np.diag([-1.0, 1.0, -1.0, 1.0]), dtype=torch.float32) @ kf["est_w2c"]
        got = keyframes.keyframe_selection_overlap_visbased_earliest_dynamic_new_topkbase(
            depth, w2c, K, kfs, 3, cfg, edge_value=6, kf_depth_thresh=0.02, earliest_thres=float(thres),
            lower_earliest_thres_percent=float(lower), topk_base=None if topk < 0 else int(topk))
        assert got == list(G[f"tk.case{j}"]), (j, got, G[f"tk.case{j}"])
    assert sorted(keyframes.quantize_selected_time_idx([0, 1, 5, 9, 9, 11], 4)) == [0, 1, 2]
