"""-m gpu: the caller-side kernels of SURVEY 8(f) N3 / N4 -- device point-to-plane metric against the CPU restatement of
the reference's compute_point2plane_dist (tests/p2p_ref.py), evaluation metrics against the reference's formulas written
with torch ops, and the rendered-sequence evaluation on a short SLAM run."""
import os
import tempfile

import numpy as np
import pytest
import torch

import p2p_ref
from vtgaussian_slam_b200 import keyframes, metrics, synthetic
from vtgaussian_slam_b200.evaluation import FrameEvaluator, eval_sequence
from vtgaussian_slam_b200.slam_loop import (LoopConfig, ViewTiedSLAM, export_params_ls, import_params_ls, section_from_frame)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _two_frames(w=160, h=120, step_m=0.05, step_deg=2.0):
    W, H, K = synthetic.intrinsics("tum_fr1", w, h)
    poses = synthetic.trajectory(2, step_m=step_m, step_deg=step_deg, seed=3)
    f0 = synthetic.make_frame("tum_fr1", w, h, seed=0, c2w=poses[0])
    f1 = synthetic.make_frame("tum_fr1", w, h, seed=1, c2w=poses[1])
    d0, d1 = torch.tensor(f0["depth"]), torch.tensor(f1["depth"])
    d0[0, :7, :9] = 0.0
    d1[0, -5:, :] = 0.0
    Kt = torch.tensor(K, dtype=torch.float32)
    w0, w1 = (torch.tensor(np.linalg.inv(p), dtype=torch.float32) for p in poses)
    return d0, d1, Kt, w0, w1


@pytest.mark.parametrize("frustum", [True, False])
def test_point2plane_matches_the_cpu_restatement(frustum):
    d0, d1, K, w0, w1 = _two_frames()
    # a slightly wrong pose for the current frame, as inside a tracking loop
    w1 = w1.clone()
    w1[:3, 3] += torch.tensor([0.004, -0.002, 0.003])
    ref_sum, ref_d = p2p_ref.point2plane_dist(d0, d1, K, w0, w1, frustum=frustum, method="sum")
    got_sum, got_d, idx = keyframes.point2plane_dist(d0.to(DEV), d1.to(DEV), K, w0, w1, frustum=frustum, method="sum", return_pairs=True)
    got_d = got_d.cpu()
    paired_ref, paired_got = ~torch.isnan(ref_d), ~torch.isnan(got_d)
    assert paired_ref.sum() > 0.5 * d1.numel()
    # the same source points find a partner (points exactly at the 2 cm radius may differ), with the same distance
    assert (paired_ref != paired_got).float().mean() < 1e-3
    both = paired_ref & paired_got
    close = (ref_d[both] - got_d[both]).abs() < 1e-5
    assert close.float().mean() > 0.999          # fp32 ties between two equally near neighbours
    assert abs(float(got_sum) / float(ref_sum) - 1.0) < 2e-3
    for method in ("max", "max100"):
        a, _ = p2p_ref.point2plane_dist(d0, d1, K, w0, w1, frustum=frustum, method=method)
        b = keyframes.point2plane_dist(d0.to(DEV), d1.to(DEV), K, w0, w1, frustum=frustum, method=method)
        assert abs(float(a) - float(b)) < 2e-5 + 1e-3 * abs(float(a))
    assert int((idx >= 0).sum()) == int(paired_got.sum())


def test_point2plane_prefers_the_true_pose_and_is_repeatable():
    d0, d1, K, w0, w1 = _two_frames()
    good = keyframes.point2plane_dist(d0.to(DEV), d1.to(DEV), K, w0, w1)
    off = w1.clone()
    off[:3, 3] += 0.006
    bad = keyframes.point2plane_dist(d0.to(DEV), d1.to(DEV), K, w0, off)
    again = keyframes.point2plane_dist(d0.to(DEV), d1.to(DEV), K, w0, off)
    assert float(good) < 0.2 * float(bad)
    assert float(bad) == float(again)              # lists are built with atomics; the nearest-neighbour choice is not order dependent
    with pytest.raises(Exception):
        keyframes.point2plane_dist(d0, d1, K, w0, w1)          # CPU tensors: no fallback


def _scene(w=320, h=240):
    W, H, K = synthetic.intrinsics("tum_fr1", w, h)
    poses = synthetic.trajectory(2, step_m=0.02, step_deg=0.5, seed=5)
    f0 = synthetic.make_frame("tum_fr1", w, h, seed=0, c2w=poses[0])
    f1 = synthetic.make_frame("tum_fr1", w, h, seed=1, c2w=poses[1])
    s = synthetic.setup_camera(W, H, K, np.eye(4))
    from gpu_helpers import settings_from
    return W, H, K, poses, f0, f1, settings_from(s, torch.device(DEV))


@pytest.mark.parametrize("use_presence", [False, True])
def test_eval_metrics_match_the_reference_formulas(use_presence):
    W, H, K, poses, f0, f1, settings = _scene()
    params = section_from_frame(torch.as_tensor(f0["im"]).to(DEV), torch.as_tensor(f0["depth"]).to(DEV), K, poses[0], DEV)
    ev = FrameEvaluator(settings, DEV)
    img = ev.render(params, np.linalg.inv(poses[1]))
    gt_rgb = torch.as_tensor(f1["im"]).to(DEV)
    gt_depth = torch.as_tensor(f1["depth"]).to(DEV).clone()
    gt_depth[0, :20, :30] = 0.0
    sil_thres = 0.5
    m = ev.frame_metrics(img, gt_rgb, gt_depth, sil_thres, use_presence=use_presence)
    # utils/eval_helpers.py:431-477 with torch ops
    valid = gt_depth > 0
    presence = img[4] > sil_thres
    w = valid & presence if use_presence else valid
    psnr = metrics.psnr(img[:3] * w, gt_rgb * w).mean()
    diff = torch.abs((img[3:4] * valid - gt_depth) * (presence if use_presence else 1.0)) * valid
    l1 = diff.sum() / valid.sum()
    assert abs(m["psnr"] - float(psnr)) < 1e-3
    assert abs(m["depth_l1"] - float(l1)) < 1e-6 + 1e-5 * float(l1)
    assert m["depth_rmse"] == m["depth_l1"] and m["valid"] == float(valid.sum())
    want_ssim = metrics.ms_ssim((img[:3] * w)[None].cpu(), (gt_rgb * w)[None].cpu())
    assert abs(m["ms_ssim"] - float(want_ssim)) < 1e-4
    assert 15.0 < m["psnr"] < 60.0 and 0.3 < m["ms_ssim"] <= 1.0


def test_rendered_sequence_evaluation_of_a_short_run():
    n = 9
    W, H, K = synthetic.intrinsics("tum_fr1", 320, 240)
    poses = synthetic.trajectory(n, step_m=0.01, step_deg=0.3)
    frames = [synthetic.make_frame("tum_fr1", 320, 240, seed=i, c2w=poses[i]) for i in range(n)]
    cfg = LoopConfig(track_iters=40, map_iters=10, baseframe_every=4, map_every=2)
    slam = ViewTiedSLAM(W, H, K, cfg, device=DEV)
    for fr in frames:
        slam.process(fr)
    with tempfile.TemporaryDirectory() as d:
        path = export_params_ls(os.path.join(d, "params_ls.npy"), slam.store, slam.w2c)
        store, w2c = import_params_ls(path, DEV)
    res = eval_sequence(frames, store, w2c, slam.settings, baseframe_every=4, sil_thres=0.5, eval_every=2, gt_c2w=list(poses))
    assert res["frame"] == [0, 2, 4, 6, 8] and len(res["psnr"]) == 5
    assert res["avg_psnr"] > 22.0 and res["avg_depth_l1"] < 0.03 and res["avg_ms_ssim"] > 0.8, res
    assert res["ate_rmse"] < 0.01 and res["lpips"] is None
    # frames rendered from a non-adjacent pair of sections (the reference's baseframe_corr_list) go through gather()
    res2 = eval_sequence(frames, store, w2c, slam.settings, baseframe_every=4, sil_thres=0.5, eval_every=4,
                         baseframe_corr_list=[[0, 4], [0, 8]], gt_c2w=list(poses))
    assert res2["frame"] == [0, 4, 8] and res2["avg_psnr"] > 20.0


@pytest.mark.parametrize("size", [(48, 64), (33, 50), (96, 128), (30, 44)])
def test_device_frame_conversion_matches_the_cpu_loader(size):
    """vtgs_frame_convert on the raw decoded bytes == the reference-pinned CPU path (cv2.resize on float64, tests/test_frames.py)."""
    from vtgaussian_slam_b200 import frames
    fix = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_fixture")
    cam = dict(image_height=48, image_width=64, fx=57.7, fy=57.9, cx=31.9, cy=23.8, png_depth_scale=1000.0)
    src = frames.ScannetSource(cam, os.path.join(fix, "scannet"), "scene0000_00", desired_height=size[0], desired_width=size[1])
    for i in (0, 5, 10):
        cpu = src[i]
        raw = src.decode_raw(i)
        im, depth = src.convert_on_device(raw[0], raw[1], DEV)
        assert im.shape == cpu["im"].shape and depth.shape == cpu["depth"].shape
        assert torch.equal(depth.cpu(), cpu["depth"])                              # nearest + fp64 division: bit-exact
        d = (im.cpu() - cpu["im"]).abs()
        assert float(d.max()) <= 6e-8 and float((d > 0).float().mean()) < 1e-3    # fp64 bilinear, last float32 bit at most
    got = list(src.prefetch(DEV, ahead=2, device_convert=True))
    assert [f["index"] for f in got] == list(range(len(src)))
    assert got[4]["im"].is_cuda and torch.allclose(got[4]["im"].cpu(), src[4]["im"], atol=6e-8) and torch.equal(got[4]["depth"].cpu(), src[4]["depth"])
    plain = list(src.prefetch(DEV, ahead=2))                                        # CPU conversion + pinned upload
    assert torch.equal(plain[7]["im"].cpu(), src[7]["im"])
