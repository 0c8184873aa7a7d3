"""-m gpu parity tests of the drop-in rasteriser (vtgs_forward / vtgs_backward through the
C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): radii, sort keys, tile ranges, per-pixel contributor counts
bit-exact; image / depth within 1e-4 absolute; gradients within 1e-3 relative."""
import numpy as np
import pytest
import torch

import oracle
from helpers import oracle_camera, rel_err
from vtgaussian_slam_b200 import synthetic

pytestmark = pytest.mark.gpu

IMG_ATOL = 1e-4
GRAD_RTOL = 1e-3


def _compare_forward(ref, got, N):
    assert got["R"] == ref["R"]
    assert np.array_equal(got["radii"], ref["radii"])
    assert np.array_equal(got["tiles_touched"], ref["tiles_touched"])
    assert np.array_equal(got["ranges"], ref["ranges"])
    assert np.array_equal(got["point_list"], ref["point_list"])
    assert np.array_equal(got["keys"], ref["keys"])
    assert np.array_equal(got["n_contrib"], ref["n_contrib"])
    vis = ref["radii"] > 0
    if N:
        assert np.array_equal(got["means2D"][vis], ref["means2D"][vis])
        assert np.array_equal(got["depths"][vis], ref["depths"][vis])
        assert np.array_equal(got["conic_opacity"][vis], ref["conic_opacity"][vis])
    assert np.abs(got["color"] - ref["color"]).max() <= IMG_ATOL
    assert np.abs(got["depth"] - ref["depth"]).max() <= IMG_ATOL * max(1.0, np.abs(ref["depth"]).max())
    assert np.abs(got["final_T"] - ref["final_T"]).max() <= IMG_ATOL


def _run_case(W, H, sc, K, bg=(0, 0, 0), grads=True, seed=0, sigma_mult=3.0, tile_rows=(0, 0)):
    from gpu_helpers import cuda_forward_all
    from vtgaussian_slam_b200 import rasterizer
    rasterizer.set_radius_sigma_mult(sigma_mult)
    try:
        cam_o, s = oracle_camera(W, H, K, bg=bg, sigma_mult=sigma_mult, tile_rows=tile_rows)
        o = oracle.Oracle()
        ref = o.forward(cam_o, sc["means3D"], sc["scales"], sc["rotations"], sc["opacities"], sc["colors"])
        got, cam, ws = cuda_forward_all(s, sc, tile_rows=tile_rows)
        N = sc["means3D"].shape[0]
        _compare_forward(ref, got, N)
        bit_exact = np.array_equal(got["color"], ref["color"]) and np.array_equal(got["final_T"], ref["final_T"])
        if grads and N:
            rng = np.random.default_rng(seed)
            dL = rng.normal(size=(3, H, W)).astype(np.float32)
            g_ref = o.backward(dL)
            g = rasterizer.rasterize_backward(cam, ws, torch.tensor(dL, device="cuda:0"))
            names = ["means3D", "means2D", "colors", "opacities", "scales", "rotations"]
            # dL/dq of an isotropic splat is analytically zero (Sigma = s^2 I is rotation invariant): both
            # sides then hold cancellation noise, so rotations are compared on the scale s * |dL/ds|
            rot_floor = float(np.abs(g_ref["scales"]).max() * np.abs(sc["scales"]).max())
            for name, t in zip(names, g):
                e = rel_err(t.cpu().numpy().reshape(g_ref[name].shape), g_ref[name],
                            floor=max(rot_floor, 1e-6) if name == "rotations" else 1e-6)
                assert e <= GRAD_RTOL, (name, e)
        return ref, got, bit_exact
    finally:
        rasterizer.set_radius_sigma_mult(3.0)


@pytest.mark.parametrize("seed,aniso,n,W,H,bg", [
    (0, True, 2000, 64, 48, (0, 0, 0)),
    (1, False, 3000, 100, 70, (0.2, 0.4, 0.1)),       # ragged edge tiles, non-zero background
    (2, True, 20000, 320, 200, (0, 0, 0)),
    (3, True, 5000, 33, 17, (0, 0, 0)),                # > 256 entries per tile: several staged batches
])
def test_random_scene_parity(seed, aniso, n, W, H, bg):
    K, sc = synthetic.random_scene(n, W, H, seed=seed, anisotropic=aniso)
    _, _, bit_exact = _run_case(W, H, sc, K, bg=bg, seed=seed)
    assert bit_exact, "forward planes are expected to be bit-identical to the oracle"


@pytest.mark.parametrize("seed", list(range(10, 22)))
def test_seed_sweep_small_scenes(seed):
    """A dozen more random scenes of varying shape (odd image sizes, mixed iso/anisotropic, scales from sub-pixel to a
    third of the image, opaque to nearly transparent): the forward must stay bit-identical, gradients within 1e-3."""
    rng = np.random.default_rng(seed)
    W, H = int(rng.integers(17, 150)), int(rng.integers(17, 110))
    n = int(rng.integers(50, 4000))
    lo = float(rng.choice([0.2, 0.5, 2.0]))
    hi = lo * float(rng.choice([2.0, 8.0, 30.0]))
    K, sc = synthetic.random_scene(n, W, H, seed=seed, anisotropic=bool(seed % 2), scale_px=(lo, min(hi, W / 3)),
                                   opacity_range=(0.01, 1.0) if seed % 3 else (0.5, 1.0))
    bg = (0, 0, 0) if seed % 4 else (0.3, 0.1, 0.7)
    _, _, bit_exact = _run_case(W, H, sc, K, bg=bg, seed=seed)
    assert bit_exact


def test_long_tile_lists_take_the_global_sort_path():
    # one 16x16 tile with > 4096 entries: the per-tile sort leaves shared memory
    W, H = 16, 16
    K, sc = synthetic.random_scene(9000, W, H, seed=4, anisotropic=False, scale_px=(0.3, 1.0), opacity_range=(0.02, 0.2))
    ref, got, _ = _run_case(W, H, sc, K, seed=4)
    assert (ref["ranges"][:, 1] - ref["ranges"][:, 0]).max() > 4096


@pytest.mark.parametrize("n,quantum", [(1500, 2e-3), (7000, 1e-3), (3000, 0.5)])
def test_short_and_long_equal_depth_runs(n, quantum):
    # per-tile radix sort on the depth bits + repair of equal-depth runs by Gaussian index: short runs (local
    # repair) in the shared-memory (n <= 2048) and the streaming (n > 2048) variants, and runs far beyond the
    # repair limit (re-sorted by the 64-bit network)
    W, H = 16, 16
    K, sc = synthetic.random_scene(n, W, H, seed=6, anisotropic=False, scale_px=(0.3, 1.0), opacity_range=(0.02, 0.2))
    z_old = sc["means3D"][:, 2].copy()
    z_new = np.maximum(np.round(z_old / quantum) * quantum, quantum).astype(np.float32)
    sc["means3D"] *= (z_new / z_old)[:, None]                       # same pixel, quantised depth
    vis = z_new > 0.2
    assert len(np.unique(z_new[vis])) < vis.sum()                   # there are exact ties
    ref, got, bit_exact = _run_case(W, H, sc, K, seed=6)
    assert bit_exact


def test_saturation_and_equal_depth_ties():
    W, H = 48, 48
    K, sc = synthetic.random_scene(6000, W, H, seed=5, anisotropic=False, opacity_range=(0.6, 1.0), scale_px=(2.0, 6.0))
    sc["means3D"][:, 2] = np.round(sc["means3D"][:, 2] * 4) / 4 + 0.25      # many exactly equal depths
    ref, got, _ = _run_case(W, H, sc, K, seed=5)
    lens = ref["ranges"][:, 1] - ref["ranges"][:, 0]
    assert ref["n_contrib"].max() < lens.max()            # early termination happened
    assert (ref["final_T"] < 2e-4).any()


def test_view_tied_frame_parity():
    fr = synthetic.make_frame("replica", 300, 170, seed=0)
    p = synthetic.view_tied_gaussians(fr, n_edge=8000, opacity="trained")
    m, s, r, o, c6 = oracle.frontend(p["means3D"], p["rgb_colors"], p["unnorm_rotations"], p["logit_opacities"],
                                     p["log_scales"], [1, 0, 0, 0], [0, 0, 0])
    sc = dict(means3D=m, scales=s, rotations=r, opacities=o, colors=c6[:, :3].copy())
    _run_case(fr["W"], fr["H"], sc, fr["K"], seed=6)
    sc["colors"] = c6[:, 3:].copy()                        # the depth / silhouette / depth^2 pass
    _run_case(fr["W"], fr["H"], sc, fr["K"], seed=7)


def test_radius_multiplier_and_tile_band():
    W, H = 128, 96
    K, sc = synthetic.random_scene(4000, W, H, seed=8)
    _run_case(W, H, sc, K, sigma_mult=2.0, seed=8)
    _run_case(W, H, sc, K, tile_rows=(2, 5), seed=9)


def test_empty_and_all_culled():
    W, H = 40, 24
    K, sc = synthetic.random_scene(10, W, H, seed=1)
    empty = {k: v[:0] for k, v in sc.items()}
    ref, got, _ = _run_case(W, H, empty, K, bg=(0.1, 0.2, 0.3), grads=False)
    assert got["R"] == 0 and np.allclose(got["color"][:, 0, 0], [0.1, 0.2, 0.3])
    sc["means3D"][:, 2] = -1.0
    _run_case(W, H, sc, K)


def test_autograd_module_matches_oracle_and_rejects_cpu():
    from gpu_helpers import settings_from
    from diff_gaussian_rasterization import GaussianRasterizer
    W, H = 96, 64
    K, sc = synthetic.random_scene(2500, W, H, seed=11)
    cam_o, s = oracle_camera(W, H, K)
    dev = torch.device("cuda:0")
    t = {k: torch.tensor(v, device=dev, requires_grad=True) for k, v in sc.items()}
    m2d = torch.zeros_like(t["means3D"], requires_grad=True)
    op = t["opacities"][:, None]
    rend = GaussianRasterizer(raster_settings=settings_from(s, dev))
    color, radii, depth = rend(means3D=t["means3D"], means2D=m2d, opacities=op, colors_precomp=t["colors"],
                               scales=t["scales"], rotations=t["rotations"])
    assert color.shape == (3, H, W) and depth.shape == (1, H, W) and radii.dtype == torch.int32
    assert not depth.requires_grad and color.requires_grad
    dL = torch.randn(3, H, W, device=dev)
    (color * dL).sum().backward()
    o = oracle.Oracle()
    ref = o.forward(cam_o, sc["means3D"], sc["scales"], sc["rotations"], sc["opacities"], sc["colors"])
    g = o.backward(dL.cpu().numpy())
    assert np.array_equal(radii.cpu().numpy(), ref["radii"])
    assert rel_err(m2d.grad.cpu().numpy(), g["means2D"]) <= GRAD_RTOL
    assert rel_err(t["means3D"].grad.cpu().numpy(), g["means3D"]) <= GRAD_RTOL
    assert rel_err(t["opacities"].grad.cpu().numpy(), g["opacities"]) <= GRAD_RTOL
    vis = rend.markVisible(t["means3D"]).cpu().numpy()
    assert np.array_equal(vis, oracle.mark_visible(cam_o, sc["means3D"]))
    with pytest.raises(Exception):
        rend(means3D=t["means3D"], means2D=m2d, opacities=op, scales=t["scales"], rotations=t["rotations"])
    with pytest.raises(RuntimeError):
        rend(means3D=t["means3D"].cpu(), means2D=m2d.cpu(), opacities=op.cpu(), colors_precomp=t["colors"].cpu(),
             scales=t["scales"].cpu(), rotations=t["rotations"].cpu())


def test_giant_offscreen_and_degenerate_splats():
    """Robustness of the binning: splats larger than the image (every tile, all 8 regions), splats far off screen
    (saturating float->int casts in getRect), needle-thin and tiny splats, zero / >1 opacities, huge depth spread."""
    W, H = 150, 90
    K, sc = synthetic.random_scene(600, W, H, seed=21, anisotropic=True, scale_px=(0.2, 3.0))
    f = K[0, 0]
    sc["scales"][:20] *= 200.0                                   # giants: radius >> image
    sc["scales"][20:40, 0] *= 1e-4                               # needles
    sc["scales"][40:60] *= 1e-3                                  # sub-pixel
    sc["means3D"][60:80, 0] = 1e6                                # far off screen (x), in front of the camera
    sc["means3D"][80:90, 2] = 1e4                                # very far
    sc["means3D"][90:100, 2] = 0.2000001                         # just beyond the near cull
    sc["opacities"][100:110] = 0.0
    sc["opacities"][110:120] = 1.5                               # > 1 (clamped to 0.99 by the blend)
    sc["opacities"][120:130] = 1.0 / 255.0                       # exactly the alpha threshold
    ref, got, bit_exact = _run_case(W, H, sc, K, seed=21)
    assert bit_exact
    gx, gy = (W + 15) // 16, (H + 15) // 16
    assert ref["tiles_touched"].max() == gx * gy                 # at least one giant covers every tile


def test_gaussian_ids_beyond_24_bits():
    """N >= 2^24: the sort key carries the full 32-bit Gaussian index (the reference's ids are 32-bit) and the region
    masks are re-derived from the records.  2^24 culled Gaussians in front of a small scene must leave every stage
    output of that scene unchanged (ids shifted by 2^24), forward and backward, in both entry points."""
    from gpu_helpers import cuda_forward_all
    from vtgaussian_slam_b200 import rasterizer
    from vtgaussian_slam_b200.fused import FusedRenderer
    from gpu_helpers import settings_from
    W, H, n, PAD = 160, 96, 3000, 1 << 24
    K, sc = synthetic.random_scene(n, W, H, seed=11)
    _, s = oracle_camera(W, H, K)
    ref, cam, ws = cuda_forward_all(s, sc)
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    g_ref = [t.clone() for t in rasterizer.rasterize_backward(cam, ws, dL)]
    pad = dict(means3D=np.tile(np.array([[0.0, 0.0, -5.0]], np.float32), (PAD, 1)),          # behind the camera: culled
               scales=np.full((PAD, 3), 0.01, np.float32), rotations=np.tile(np.array([[1.0, 0, 0, 0]], np.float32), (PAD, 1)),
               opacities=np.full((PAD,), 0.5, np.float32), colors=np.zeros((PAD, 3), np.float32))
    big = {k: np.concatenate([pad[k], sc[k]]) for k in sc}
    del pad
    got, cam2, ws2 = cuda_forward_all(s, big)
    assert got["R"] == ref["R"] and np.array_equal(got["ranges"], ref["ranges"])
    assert np.array_equal(got["point_list"].astype(np.int64), ref["point_list"].astype(np.int64) + PAD)
    assert np.array_equal(got["keys"], ref["keys"])
    assert not got["radii"][:PAD].any() and np.array_equal(got["radii"][PAD:], ref["radii"])
    for k in ("color", "depth", "final_T", "n_contrib"):
        assert np.array_equal(got[k], ref[k]), k
    g = rasterizer.rasterize_backward(cam2, ws2, dL)
    for a, b in zip(g, g_ref):
        a = a.reshape(PAD + n, -1)
        assert not a[:PAD].any()
        assert rel_err(a[PAD:].cpu().numpy(), b.reshape(n, -1).cpu().numpy()) <= 1e-5
    del g, ws2, got, big
    torch.cuda.empty_cache()
    # fused six-plane path (isotropic parameters)
    fr = synthetic.make_frame("replica", 160, 96, seed=0)
    p = synthetic.view_tied_gaussians(fr, n_edge=500, opacity="trained")
    settings = settings_from(synthetic.setup_camera(fr["W"], fr["H"], fr["K"], np.eye(4)), torch.device("cuda:0"))
    q, t = synthetic.perturbed_pose(seed=1, trans_sigma=0.01, rot_deg=0.5)
    q, t = torch.tensor(q).cuda(), torch.tensor(t).cuda()
    small = {k: torch.tensor(v).cuda() for k, v in p.items()}
    n2 = small["means3D"].shape[0]
    r0 = FusedRenderer(settings, n2)
    img0 = r0.forward(small, q, t)[0].clone()
    far = {k: torch.zeros((PAD,) + v.shape[1:], device="cuda") for k, v in small.items()}
    far["means3D"][:, 2] = -5.0
    far["unnorm_rotations"][:, 0] = 1.0
    bigp = {k: torch.cat([far[k], small[k]]).contiguous() for k in small}
    del far
    r1 = FusedRenderer(settings, PAD + n2)
    img1, radii1 = r1.forward(bigp, q, t)
    assert torch.equal(img1, img0) and not radii1[:PAD].any()
    gt_rgb, gt_d = torch.tensor(fr["im"]).cuda(), torch.tensor(fr["depth"]).cuda()
    grads = []
    for r, prm in ((r0, small), (r1, bigp)):
        r.forward(prm, q, t)
        r.tracking_loss(gt_rgb, gt_d)
        dq, dt = torch.zeros(4).cuda(), torch.zeros(3).cuda()
        r.backward(prm, q, t, pose_grads=(dq, dt))
        grads.append(torch.cat([dq, dt]).cpu().numpy())
    assert rel_err(grads[1], grads[0]) <= 1e-5
