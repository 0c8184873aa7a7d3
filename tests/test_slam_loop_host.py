"""CPU: host-side pieces of the compact SLAM loop (pose algebra, posed synthetic frames)."""
import numpy as np

from vtgaussian_slam_b200 import synthetic
from vtgaussian_slam_b200.slam_loop import matrix_from_quat, propagate_pose, quat_from_matrix, ate_rmse


def _rand_rot(rng):
    q = rng.normal(size=4)
    return matrix_from_quat(q, np.zeros(3))[:3, :3]


def test_quaternion_matrix_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(200):
        R = _rand_rot(rng)
        q = quat_from_matrix(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-12 and q[0] >= 0
        assert np.allclose(matrix_from_quat(q, np.zeros(3))[:3, :3], R, atol=1e-12)
    # the trace <= 0 branches (rotations by ~pi about each axis)
    for ax in range(3):
        R = -np.eye(3)
        R[ax, ax] = 1.0
        assert np.allclose(matrix_from_quat(quat_from_matrix(R), np.zeros(3))[:3, :3], R, atol=1e-12)


def test_constant_velocity_propagation_is_exact_for_a_constant_twist():
    rng = np.random.default_rng(1)
    D = np.eye(4)
    D[:3, :3] = _rand_rot(rng)
    D[:3, 3] = rng.normal(0, 0.1, 3)
    c0 = np.eye(4)
    c0[:3, :3] = _rand_rot(rng)
    c1, c2 = D @ c0, D @ D @ c0                       # c2w[t] = D c2w[t-1]
    pred = propagate_pose(np.linalg.inv(c1), np.linalg.inv(c0))
    assert np.allclose(np.linalg.inv(pred), c2, atol=1e-10)


def test_trajectory_starts_at_identity_with_the_requested_step():
    T = synthetic.trajectory(40, step_m=0.01, step_deg=0.3)
    assert np.allclose(T[0], np.eye(4))
    step = np.linalg.norm(np.diff(T[:, :3, 3], axis=0), axis=1)
    assert 0.002 < step.mean() < 0.03
    for M in T:
        assert np.allclose(M[:3, :3] @ M[:3, :3].T, np.eye(3), atol=1e-12)
    assert ate_rmse(T, T) == 0.0


def test_posed_frame_hit_points_reproject_to_their_pixels():
    W, H, K = synthetic.intrinsics("tum_fr1", 80, 60)
    c2w = synthetic.trajectory(30, 0.02, 1.0)[17]
    depth, pts = synthetic._room_depth(W, H, K, c2w=c2w)
    w2c = np.linalg.inv(c2w)
    cam = pts.reshape(-1, 3) @ w2c[:3, :3].T + w2c[:3, 3]
    assert np.allclose(cam[:, 2], depth.reshape(-1), atol=1e-9)
    u = cam[:, 0] / cam[:, 2] * K[0, 0] + K[0, 2] - 0.5
    v = cam[:, 1] / cam[:, 2] * K[1, 1] + K[1, 2] - 0.5
    xg, yg = np.meshgrid(np.arange(W), np.arange(H))
    assert np.allclose(u, xg.reshape(-1), atol=1e-6) and np.allclose(v, yg.reshape(-1), atol=1e-6)
    # every hit lies on the room's surfaces or on a slab
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]
    on_wall = (np.isclose(x, synthetic.ROOM["x"][0]) | np.isclose(x, synthetic.ROOM["x"][1]) | np.isclose(y, synthetic.ROOM["y"][0]) |
               np.isclose(y, synthetic.ROOM["y"][1]) | np.isclose(z, synthetic.ROOM["z"][0]) | np.isclose(z, synthetic.ROOM["z"][1]))
    on_slab = np.zeros_like(on_wall)
    for s in synthetic.SLABS:
        on_slab |= np.isclose(z, s[4])
    assert bool((on_wall | on_slab).all())


def test_identity_pose_frame_equals_the_unposed_frame():
    a = synthetic.make_frame("tum_fr1", 64, 48, seed=5)
    b = synthetic.make_frame("tum_fr1", 64, 48, seed=5, c2w=np.eye(4))
    assert np.array_equal(a["depth"], b["depth"]) and np.array_equal(a["im"], b["im"])


def test_params_ls_round_trip_in_the_reference_layout(tmp_path):
    import torch
    from vtgaussian_slam_b200.slam_loop import SectionStore, export_params_ls, import_params_ls
    rng = np.random.default_rng(0)
    st = SectionStore(8, "cpu")
    secs = []
    for n in (5, 9, 3):
        p = {k: torch.tensor(rng.normal(size=(n, c)).astype(np.float32)) for k, c in SectionStore.KEYS.items()}
        secs.append(p)
        st.append(p)
    traj = [np.linalg.inv(M) for M in synthetic.trajectory(7, 0.05, 2.0)]
    path = export_params_ls(str(tmp_path / "params_ls.npy"), st, traj)
    raw = np.load(path, allow_pickle=True)                         # what the reference's eval code does
    assert raw.shape == (3,) and set(raw[0]) == set(SectionStore.KEYS) | {"cam_unnorm_rots", "cam_trans"}
    assert tuple(raw[1]["cam_unnorm_rots"].shape) == (1, 4, 7) and tuple(raw[1]["cam_trans"].shape) == (1, 3, 7)
    assert isinstance(raw[2]["means3D"], torch.Tensor) and raw[2]["means3D"].shape == (3, 3)
    st2, traj2 = import_params_ls(path)
    assert len(st2) == 3 and st2.num_gaussians == 17
    for k in range(3):
        for name in SectionStore.KEYS:
            assert torch.equal(st2.rows(k)[name], secs[k][name])
    assert len(traj2) == 7 and all(np.allclose(a, b, atol=1e-6) for a, b in zip(traj, traj2))


def test_section_builders_match_the_reference_point_clouds():
    """section_from_frame / geometric_edge_mask / densified_section (run on the CPU device here: plain torch ops, the same
    on CUDA) against get_pointcloud + geometric_edge_mask of the reference, called as initialize_params_base_timestep does."""
    import os
    import torch
    from vtgaussian_slam_b200.slam_loop import densified_section, geometric_edge_mask, section_from_frame
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "section_golden.npz"))
    im, depth, im2, depth2 = (torch.tensor(G[k]) for k in ("im", "depth", "im2", "depth2"))
    edge = geometric_edge_mask((im.permute(1, 2, 0) * 255).numpy(), dilate=True, rgb=True)
    assert np.array_equal(edge, G["edge_mask"]) and 0.01 < (edge > 0).mean() < 0.9
    n0 = int(G["n_base"])
    base = section_from_frame(im, depth, G["K"], G["pose"], "cpu")
    assert base["means3D"].shape[0] == n0 == int((G["depth"] > 0).sum())
    sec = densified_section(im, depth, G["K"], im2, depth2, G["K2"], G["pose"], edge, "cpu")
    assert sec["means3D"].shape == G["means3D"].shape and sec["means3D"].shape[0] > n0
    assert np.allclose(sec["means3D"].numpy(), G["means3D"], atol=3e-6)
    assert np.array_equal(sec["rgb_colors"].numpy(), G["rgb"])
    assert np.allclose(sec["log_scales"][:, 0].numpy(), G["log_scales"], atol=2e-6)
    # the dense Gaussians are half as wide as their neighbours of the base grid
    assert float(sec["log_scales"][n0:].mean()) < float(sec["log_scales"][:n0].mean()) - 0.4      # (~log 2, edges sit at other depths)
    assert float(sec["logit_opacities"].abs().max()) == 0.0 and float((sec["unnorm_rotations"][:, 0] - 1).abs().max()) == 0.0
