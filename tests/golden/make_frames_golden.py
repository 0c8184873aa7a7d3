"""Golden vectors for vtgaussian_slam_b200.frames from the REFERENCE's own dataset classes.

Run in the build container (needs /root/reference; not needed on the GPU box):
    python tests/golden/make_frames_golden.py

1. writes a tiny Replica-layout and a tiny TUM-layout sequence under tests/golden/frames_fixture/ (a few 64x48
   frames of the synthetic room: JPEG + 16-bit PNG depth + trajectory files) -- committed, < 100 KB;
2. imports the reference's ReplicaDataset / TUMDataset (datasets/gradslam_datasets/{replica,tum,basedataset}.py) and
   stores what they return for every frame, after the two lines the reference's main loop applies
   (src/vtgaussian_slam.py:198-202: color.permute(2,0,1) / 255, depth.permute(2,0,1)), in frames_golden.npz.

Three third-party imports of those files are absent from this image and are shimmed with equivalents:
imageio.v2.imread (Pillow decode -- imageio's own default backend), natsort.natsorted (natural sort),
kornia.geometry.linalg.compose_transformations / inverse_transformation (rigid composition formulas);
`np.unicode_` (removed in NumPy 2) is aliased to `np.str_`.
"""
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
FIX = os.path.join(HERE, "frames_fixture")

REPLICA_CAM = dict(image_height=48, image_width=64, fx=32.0, fy=32.0, cx=31.5, cy=23.5, png_depth_scale=6553.5)
TUM_CAM = dict(image_height=48, image_width=64, fx=51.73, fy=51.65, cx=31.86, cy=25.53, png_depth_scale=5000.0)


def write_fixture():
    from PIL import Image
    from vtgaussian_slam_b200 import synthetic
    n = 6
    poses = synthetic.trajectory(n, step_m=0.03, step_deg=1.0, seed=5)
    world = np.eye(4)
    world[:3, 3] = [0.4, -0.2, 0.1]                       # a non-identity first pose: relative_pose must remove it
    world[:3, :3] = synthetic.trajectory(9, 0.01, 3.0, seed=9)[7][:3, :3]
    # ---- Replica layout
    rdir = os.path.join(FIX, "replica", "room0", "results")
    os.makedirs(rdir, exist_ok=True)
    rows = []
    for i in range(n):
        fr = synthetic.make_frame("replica", 64, 48, seed=i, c2w=poses[i])
        Image.fromarray((fr["im"].transpose(1, 2, 0) * 255 + 0.5).astype(np.uint8)).save(os.path.join(rdir, f"frame{i:06d}.jpg"), quality=95)
        Image.fromarray(np.clip(fr["depth"][0] * 6553.5 + 0.5, 0, 65535).astype(np.uint16)).save(os.path.join(rdir, f"depth{i:06d}.png"))
        rows.append(" ".join(repr(float(v)) for v in (world @ poses[i]).reshape(-1)))
    with open(os.path.join(FIX, "replica", "room0", "traj.txt"), "w") as f:
        f.write("\n".join(rows) + "\n")
    # ---- TUM layout: 8 colour frames (two of them closer than 1/32 s to their predecessor), depth stamps offset by
    # a few ms, one colour frame without a pose within 0.08 s
    tdir = os.path.join(FIX, "tum", "fr1")
    os.makedirs(os.path.join(tdir, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(tdir, "depth"), exist_ok=True)
    stamps = [100.000, 100.020, 100.070, 100.140, 100.150, 100.230, 100.600, 100.680]
    src = [0, 0, 1, 2, 2, 3, 4, 5]
    rgb_l, dep_l, gt_l = ["# color images", "# file", "# timestamp filename"], ["# depth maps", "# file", "# timestamp filename"], ["# timestamp tx ty tz qx qy qz qw"]
    for t, k in zip(stamps, src):
        fr = synthetic.make_frame("tum_fr1", 64, 48, seed=10 + k, c2w=poses[k])
        Image.fromarray((fr["im"].transpose(1, 2, 0) * 255 + 0.5).astype(np.uint8)).save(os.path.join(tdir, "rgb", f"{t:.6f}.png"))
        td = t + 0.004
        Image.fromarray(np.clip(fr["depth"][0] * 5000.0 + 0.5, 0, 65535).astype(np.uint16)).save(os.path.join(tdir, "depth", f"{td:.6f}.png"))
        rgb_l.append(f"{t:.6f} rgb/{t:.6f}.png")
        dep_l.append(f"{td:.6f} depth/{td:.6f}.png")
        if abs(t - 100.600) > 1e-9:                       # no pose near 100.600
            M = world @ poses[k]
            from vtgaussian_slam_b200.slam_loop import quat_from_matrix
            w, x, y, z = quat_from_matrix(M[:3, :3])
            s = 1.7 if k == 2 else 1.0                    # an un-normalised quaternion row, as real files may hold
            gt_l.append(f"{t + 0.001:.6f} {M[0, 3]:.9f} {M[1, 3]:.9f} {M[2, 3]:.9f} {x * s:.9f} {y * s:.9f} {z * s:.9f} {w * s:.9f}")
    for name, lines in (("rgb.txt", rgb_l), ("depth.txt", dep_l), ("groundtruth.txt", gt_l)):
        with open(os.path.join(tdir, name), "w") as f:
            f.write("\n".join(lines) + "\n")
    # ---- ScanNet++ DSLR layout: 5 training + 2 test images, OpenGL-convention c2w in a nerfstudio transforms file
    import json
    sdir = os.path.join(FIX, "scannetpp", "scene0", "dslr")
    for d in ("undistorted_images", "undistorted_depths", "nerfstudio"):
        os.makedirs(os.path.join(sdir, d), exist_ok=True)
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    names = [f"DSC{100 + i:05d}.JPG" for i in range(7)]
    metas = []
    sp = synthetic.trajectory(7, step_m=0.05, step_deg=2.0, seed=8)
    for i, nme in enumerate(names):
        fr = synthetic.make_frame("scannetpp", 72, 48, seed=20 + i, c2w=sp[i])
        Image.fromarray((fr["im"].transpose(1, 2, 0) * 255 + 0.5).astype(np.uint8)).save(os.path.join(sdir, "undistorted_images", nme), quality=95)
        Image.fromarray(np.clip(fr["depth"][0] * 1000.0 + 0.5, 0, 65535).astype(np.uint16)).save(
            os.path.join(sdir, "undistorted_depths", nme.replace(".JPG", ".png")))
        gl = flip @ (world @ sp[i]) @ flip                     # OpenCV c2w -> OpenGL c2w (P is its own inverse)
        metas.append(dict(file_path=nme, transform_matrix=gl.tolist(), is_bad=bool(i == 3)))
    Ws, Hs, Ks = synthetic.intrinsics("scannetpp", 72, 48)
    with open(os.path.join(sdir, "nerfstudio", "transforms_undistorted.json"), "w") as f:
        json.dump(dict(w=72, h=48, fl_x=float(Ks[0, 0]), fl_y=float(Ks[1, 1]), cx=float(Ks[0, 2]), cy=float(Ks[1, 2]),
                       frames=[metas[i] for i in (4, 0, 2, 1, 3)], test_frames=[metas[6], metas[5]]), f)
    with open(os.path.join(sdir, "train_test_lists.json"), "w") as f:
        json.dump(dict(train=[names[i] for i in range(5)], test=[names[5], names[6]]), f)


SCANNET_CAM = dict(image_height=48, image_width=64, fx=57.7, fy=57.9, cx=31.9, cy=23.8, png_depth_scale=1000.0)


def write_scannet_fixture():
    """ScanNet (v2) export layout: color/<i>.jpg, depth/<i>.png (mm), pose/<i>.txt (4x4 c2w); 11 frames so that natural
    order (0, 1, 2, ..., 10) differs from lexicographic order (0, 1, 10, 2, ...)."""
    from PIL import Image
    from vtgaussian_slam_b200 import synthetic
    sdir = os.path.join(FIX, "scannet", "scene0000_00")
    for d in ("color", "depth", "pose"):
        os.makedirs(os.path.join(sdir, d), exist_ok=True)
    poses = synthetic.trajectory(11, step_m=0.04, step_deg=1.5, seed=15)
    world = np.eye(4)
    world[:3, 3] = [-0.3, 0.5, 0.2]
    for i in range(11):
        fr = synthetic.make_frame("tum_fr1", 64, 48, seed=40 + i, c2w=poses[i])
        Image.fromarray((fr["im"].transpose(1, 2, 0) * 255 + 0.5).astype(np.uint8)).save(os.path.join(sdir, "color", f"{i}.jpg"), quality=92)
        Image.fromarray(np.clip(fr["depth"][0] * 1000.0 + 0.5, 0, 65535).astype(np.uint16)).save(os.path.join(sdir, "depth", f"{i}.png"))
        np.savetxt(os.path.join(sdir, "pose", f"{i}.txt"), world @ poses[i])


def shim_missing_modules():
    from PIL import Image
    iio = types.ModuleType("imageio")
    v2 = types.ModuleType("imageio.v2")
    v2.imread = lambda p: np.asarray(Image.open(p))
    iio.v2 = v2
    iio.imread = v2.imread
    sys.modules.setdefault("imageio", iio)
    sys.modules.setdefault("imageio.v2", v2)
    ns = types.ModuleType("natsort")
    ns.natsorted = lambda xs: sorted(xs, key=lambda s: [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)])
    sys.modules.setdefault("natsort", ns)

    def compose(a, b):                                    # kornia.geometry.linalg.compose_transformations
        out = torch.zeros_like(a)
        out[..., :3, :3] = a[..., :3, :3] @ b[..., :3, :3]
        out[..., :3, 3:] = a[..., :3, :3] @ b[..., :3, 3:] + a[..., :3, 3:]
        out[..., 3, 3] = 1.0
        return out

    def inverse(t):
        out = torch.zeros_like(t)
        Rt = t[..., :3, :3].transpose(-1, -2)
        out[..., :3, :3] = Rt
        out[..., :3, 3:] = -Rt @ t[..., :3, 3:]
        out[..., 3, 3] = 1.0
        return out
    k = types.ModuleType("kornia")
    kg = types.ModuleType("kornia.geometry")
    kl = types.ModuleType("kornia.geometry.linalg")
    kl.compose_transformations, kl.inverse_transformation = compose, inverse
    k.geometry, kg.linalg = kg, kl
    for name, mod in (("kornia", k), ("kornia.geometry", kg), ("kornia.geometry.linalg", kl)):
        sys.modules.setdefault(name, mod)


def reference_outputs():
    shim_missing_modules()
    if not hasattr(np, "unicode_"):
        np.unicode_ = np.str_             # tum.py:47 uses the NumPy < 2 alias
    # import the three files without running the package __init__ (it pulls in every other dataset's dependencies)
    import importlib.util
    pkg = types.ModuleType("refds")
    pkg.__path__ = ["/root/reference/datasets/gradslam_datasets"]
    sys.modules["refds"] = pkg

    def load(name):
        spec = importlib.util.spec_from_file_location(f"refds.{name}", f"/root/reference/datasets/gradslam_datasets/{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refds.{name}"] = m
        spec.loader.exec_module(m)
        return m
    for n in ("geometryutils", "datautils", "basedataset"):
        load(n)
    replica, tum, spp, scannet = load("replica"), load("tum"), load("scannetpp"), load("scannet")
    out = {}
    cases = [
        ("replica_native", replica.ReplicaDataset, dict(dataset_name="replica", camera_params=REPLICA_CAM), os.path.join(FIX, "replica"), "room0", dict(desired_height=48, desired_width=64)),
        ("replica_resized", replica.ReplicaDataset, dict(dataset_name="replica", camera_params=REPLICA_CAM), os.path.join(FIX, "replica"), "room0", dict(desired_height=30, desired_width=44, start=1, end=6, stride=2)),
        ("tum_native", tum.TUMDataset, dict(dataset_name="tum", camera_params=TUM_CAM), os.path.join(FIX, "tum"), "fr1", dict(desired_height=48, desired_width=64)),
        ("tum_resized", tum.TUMDataset, dict(dataset_name="tum", camera_params=TUM_CAM), os.path.join(FIX, "tum"), "fr1", dict(desired_height=24, desired_width=32, start=1)),
    ]
    cases += [
        ("scannet_native", scannet.ScannetDataset, dict(dataset_name="scannet", camera_params=SCANNET_CAM), os.path.join(FIX, "scannet"), "scene0000_00", dict(desired_height=48, desired_width=64)),
        ("scannet_resized", scannet.ScannetDataset, dict(dataset_name="scannet", camera_params=SCANNET_CAM), os.path.join(FIX, "scannet"), "scene0000_00", dict(desired_height=33, desired_width=50, start=2, stride=3)),
    ]
    spp_cases = [
        ("scannetpp_train", dict(desired_height=48, desired_width=72)),
        ("scannetpp_train_nobad", dict(desired_height=24, desired_width=36, ignore_bad=True, start=1)),
        ("scannetpp_test", dict(desired_height=48, desired_width=72, use_train_split=False)),
    ]
    datasets = [(tag, cls(cfg, basedir, seq, device="cpu", **{"stride": 1, **kw})) for tag, cls, cfg, basedir, seq, kw in cases]
    datasets += [(tag, spp.ScannetPPDataset(os.path.join(FIX, "scannetpp"), "scene0", device="cpu", **{"stride": 1, **kw})) for tag, kw in spp_cases]
    for tag, ds in datasets:
        ims, deps, Ks, Ps = [], [], [], []
        for i in range(len(ds)):
            color, depth, intr, pose = ds[i]
            ims.append((color.permute(2, 0, 1) / 255).numpy())          # src/vtgaussian_slam.py:201
            deps.append(depth.permute(2, 0, 1).numpy())                  # :202
            Ks.append(intr[:3, :3].numpy())
            Ps.append(pose.numpy())
        out[f"{tag}.im"], out[f"{tag}.depth"] = np.stack(ims), np.stack(deps)
        out[f"{tag}.K"], out[f"{tag}.c2w"] = np.stack(Ks), np.stack(Ps)
        out[f"{tag}.files"] = np.array([os.path.relpath(p, FIX) for p in ds.color_paths])
    return out


if __name__ == "__main__":
    write_fixture()
    write_scannet_fixture()
    g = reference_outputs()
    np.savez_compressed(os.path.join(HERE, "frames_golden.npz"), **g)
    for k, v in g.items():
        print(k, v.shape, v.dtype)
