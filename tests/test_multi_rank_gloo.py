"""CPU, world_size 2 over gloo: the host-side logic of the two sharding modes.

* tracking (tile-row bands): every rank evaluates ONLY its band (oracle with tile_row_begin/end),
  the 16-float message (pose-gradient partials + loss terms) is all-reduced, and the result must
  equal the single-rank full-image evaluation -- the property the GPU path relies on
  (fused.TrackingSolver(process_group=...), bench.py --gpus N).
* mapping (keyframe shards): per-keyframe parameter gradients all-reduced == sum over keyframes.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tracking_terms(tile_rows, q, t):
    """Loss terms and dL/d(cam-frame means) sums of one band through the oracle."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import oracle_camera
    from vtgaussian_slam_b200 import synthetic
    fr = synthetic.make_frame("replica", 160, 96, seed=0)
    p = synthetic.view_tied_gaussians(fr, opacity="trained")
    cam, _ = oracle_camera(fr["W"], fr["H"], fr["K"], tile_rows=tile_rows)
    m, s, r, o, c6 = oracle.frontend(p["means3D"], p["rgb_colors"], p["unnorm_rotations"], p["logit_opacities"], p["log_scales"], q, t)
    orc = oracle.Oracle()
    out = orc.forward(cam, m, s, r, o, c6)
    img, gd = out["color"], fr["depth"][0]
    mask = (gd > 0) & (img[4] > 0.99)
    if tile_rows[1] > tile_rows[0]:
        band = np.zeros_like(mask)
        band[tile_rows[0] * 16:tile_rows[1] * 16] = True
        mask &= band
    dL = np.zeros((6,) + gd.shape, np.float32)
    dL[:3] = 0.5 * np.sign(img[:3] - fr["im"]) * mask
    dL[3] = 0.025 * np.sign(img[3] - gd) * mask
    loss_im = 0.5 * np.abs(img[:3] - fr["im"])[:, mask].sum()
    loss_d = 0.025 * np.abs(img[3] - gd)[mask].sum()
    g = orc.backward(dL)
    gm = g["means3D"].astype(np.float64)
    gm[:, 2] += g["colors"][:, 3]                      # depth channel chains into z (depth_row = 0,0,1,0)
    msg = np.zeros(16)
    msg[0:3] = gm.sum(0)                               # dL/dt
    msg[3:12] = (gm[:, :, None] * p["means3D"][:, None, :].astype(np.float64)).sum(0).reshape(-1)   # dL/dR
    msg[12], msg[13], msg[14] = loss_im + loss_d, loss_im, loss_d
    msg[15] = mask.sum()
    return msg, (fr["H"] + 15) // 16


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    q, t = np.array([0.9999, 0.005, -0.003, 0.002], np.float32), np.array([0.01, -0.004, 0.006], np.float32)
    _, gy = _tracking_terms((0, 1), q, t)
    band = bench.balanced_bands(np.ones(gy), world)[rank]
    msg, _ = _tracking_terms(band, q, t)
    tmsg = torch.tensor(msg)
    dist.all_reduce(tmsg)                              # the one collective of a tracking iteration
    # mapping: keyframe k -> rank k % world; gradient all-reduce == sum over keyframes
    kf_grads = [torch.full((5,), float(k + 1)) for k in range(4)]
    mine = sum(kf_grads[k] for k in range(4) if k % world == rank)
    dist.all_reduce(mine)
    if rank == 0:
        ret["msg"] = tmsg.numpy().copy()
        ret["map"] = mine.numpy().copy()
        ret["band0"] = band
    dist.barrier()
    dist.destroy_process_group()


def test_band_partition_covers_all_rows():
    sys.path.insert(0, ROOT)
    import bench
    for gy in (43, 30, 73, 7):
        for world in (1, 2, 4, 8):
            bands = bench.balanced_bands(np.ones(gy), world)          # uniform work: an even split
            assert bands[0][0] == 0 and bands[-1][1] == gy
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_bands_properties():
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(0)
    for rows in (43, 30, 73, 8):
        for world in (1, 2, 4, 8):
            w = rng.uniform(0.2, 3.0, rows)
            bands = bench.balanced_bands(w, world)
            assert bands[0][0] == 0 and bands[-1][1] == rows and len(bands) == world
            assert all(b > a for a, b in bands) and all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            loads = [w[a:b].sum() for a, b in bands]
            if rows >= 4 * world:
                assert max(loads) <= w.sum() / world + w.max() + 1e-9        # within one row of the ideal share
    assert bench.balanced_bands(np.zeros(8), 8) == [(i, i + 1) for i in range(8)]


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_equals_single_rank():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    q, t = np.array([0.9999, 0.005, -0.003, 0.002], np.float32), np.array([0.01, -0.004, 0.006], np.float32)
    full, _ = _tracking_terms((0, 0), q, t)
    got = ret["msg"]
    assert got[15] == full[15]                                         # mask counts add up exactly
    np.testing.assert_allclose(got[12:15], full[12:15], rtol=1e-6)     # loss terms
    scale = np.abs(full[:12]).max()
    assert np.abs(got[:12] - full[:12]).max() <= 1e-5 * scale          # pose-gradient partial sums
    np.testing.assert_allclose(ret["map"], np.full(5, 10.0))


# ---- keyframe-sharded mapping step (vtgs_sharded_adam): the algorithm over gloo ---------------------------------------
def _adam_slice(p, g, m, v, lr, step, eps=1e-15, b1=0.9, b2=0.999):
    """vtgs_adam's update (csrc/fused.cu adam_kernel) in float32 numpy."""
    f = np.float32
    m[:] = m + f(1 - b1) * (g - m)
    v[:] = f(b2) * v + f(1 - b2) * g * g
    step_size = (lr / (1.0 - b1 ** step)).astype(np.float32)
    bc2s = f(np.sqrt(1.0 - b2 ** step))
    p[:] = p - step_size * (m / (np.sqrt(v) / bc2s + f(eps)))


def _sharded_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from vtgaussian_slam_b200.fused import MappingSolver
    sizes, lrs = [3 * 1001, 1001, 1001], [0.0025, 0.05, 0.005]          # rgb, logit opacity, log scale of 1001 Gaussians
    n = sum(sizes)
    n_pad = (n + 3) // 4 * 4
    lr_el = np.zeros(n_pad, np.float32)
    off = 0
    for c, lr in zip(sizes, lrs):
        lr_el[off:off + c] = lr
        off += c
    lr_el[off:] = lrs[-1]
    rng = np.random.default_rng(7)
    p = np.zeros(n_pad, np.float32)
    p[:n] = rng.normal(size=n).astype(np.float32)
    b, e = MappingSolver.sharded_slices(n_pad, world)[rank]
    m, v = np.zeros(e - b, np.float32), np.zeros(e - b, np.float32)       # this rank's slice of the moments only
    for step in range(1, 5):
        g_all = [np.zeros(n_pad, np.float32) for _ in range(world)]       # every rank's keyframe gradients (same seeds everywhere)
        for r in range(world):
            g_all[r][:n] = np.random.default_rng(100 * step + r).normal(size=n).astype(np.float32)
        mine = torch.tensor(g_all[rank])
        gathered = [torch.zeros(n_pad) for _ in range(world)]
        dist.all_gather(gathered, mine)                                   # stands in for the peer loads over NVLink
        gs = gathered[0][b:e].numpy().copy()
        for r in range(1, world):                                         # rank order: a fixed order
            gs = gs + gathered[r][b:e].numpy()
        sl = p[b:e].copy()
        _adam_slice(sl, gs, m, v, lr_el[b:e], step)
        per = max(x[1] - x[0] for x in MappingSolver.sharded_slices(n_pad, world))
        padded = torch.zeros(per)
        padded[:e - b] = torch.tensor(sl)
        outs = [torch.zeros(per) for _ in range(world)]
        dist.all_gather(outs, padded)                                     # stands in for the peer stores
        for r, (rb, re_) in enumerate(MappingSolver.sharded_slices(n_pad, world)):
            p[rb:re_] = outs[r][:re_ - rb].numpy()
    ret[rank] = p.copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_step_algorithm_equals_replicated_adam(world):
    """Reduce-scatter in rank order + Adam on the owned slice (sharded moments) + all-gather == torch.optim.Adam applied by
    every rank to the all-reduced gradients, and the ranks end bitwise equal.  (The CUDA kernel is checked against vtgs_adam
    in tests/test_gpu_fused.py and across 2 GPUs by tools/check_multi_gpu.py.)"""
    sys.path.insert(0, ROOT)
    from vtgaussian_slam_b200.fused import MappingSolver
    sizes, lrs = [3 * 1001, 1001, 1001], [0.0025, 0.05, 0.005]
    n = sum(sizes)
    n_pad = (n + 3) // 4 * 4
    sl = MappingSolver.sharded_slices(n_pad, world)
    assert sl[0][0] == 0 and sl[-1][1] == n_pad and all(a[1] == b[0] for a, b in zip(sl, sl[1:])) and all(b % 4 == 0 for b, _ in sl)
    assert MappingSolver.sharded_slices(8, 8) == [(0, 4), (4, 8)] + [(8, 8)] * 6          # more ranks than float4s
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, 29600 + world, ret), nprocs=world, join=True)
    for r in range(1, world):
        assert np.array_equal(ret[0], ret[r])
    # single process: torch.optim.Adam per tensor on the summed gradients
    rng = np.random.default_rng(7)
    p0 = rng.normal(size=n).astype(np.float32)
    offs = np.cumsum([0] + sizes)
    ts = [torch.nn.Parameter(torch.tensor(p0[offs[i]:offs[i + 1]].copy())) for i in range(3)]
    opt = torch.optim.Adam([{"params": [t], "lr": lr} for t, lr in zip(ts, lrs)], lr=0.0, eps=1e-15)
    for step in range(1, 5):
        g = sum(np.random.default_rng(100 * step + r).normal(size=n).astype(np.float32) for r in range(world))
        for i, t in enumerate(ts):
            t.grad = torch.tensor(g[offs[i]:offs[i + 1]].copy())
        opt.step()
    ref = np.concatenate([t.detach().numpy() for t in ts])
    assert np.abs(ret[0][:n] - ref).max() <= 5e-6
    assert not ret[0][n:].any()

