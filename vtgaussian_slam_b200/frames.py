"""Frame sources: what feeds the hot path one RGB-D frame at a time.

On-disk sequences in the two layouts the reference's headline configs use, plus the synthetic generator, behind one
interface; every source returns frames already in the form the tracking / mapping loop consumes:

    im[3,H,W] float32 in [0,1], depth[1,H,W] float32 metres, K[3,3] float32 (scaled to H x W),
    c2w[4,4] float32 relative to the sequence's first retained frame.

Reference behaviour mirrored (datasets/gradslam_datasets/):
  * basedataset.py:106-202   start / end / stride slicing, `relative_pose=True` -> inv(T_0) T_i (:274-292)
  * basedataset.py:215-272   colour resized with cv2 INTER_LINEAR as float64, depth with INTER_NEAREST then divided
                             by `png_depth_scale`
  * basedataset.py:311-365   __getitem__: colour kept 0..255 HWC there; the main loop then does
                             `color.permute(2,0,1) / 255`, `depth.permute(2,0,1)` (src/vtgaussian_slam.py:198-202,
                             :291-295) -- that last step is folded in here
  * datautils.py:73-117      scale_intrinsics (fx, cx by the width ratio; fy, cy by the height ratio)
  * replica.py:44-66         results/frame*.jpg, results/depth*.png in natural order, traj.txt = one row-major c2w per line
  * tum.py:44-160            rgb.txt / depth.txt / groundtruth.txt, nearest-timestamp association within 0.08 s, frames
                             at least 1/32 s apart, pose rows (tx ty tz qx qy qz qw)
  * scannet.py:14-61         color/*.jpg, depth/*.png, pose/*.txt (4x4 c2w per frame) in natural order
  * scannetpp.py:18-135      dslr/train_test_lists.json + dslr/nerfstudio/transforms_undistorted.json (intrinsics and
                             per-image OpenGL c2w, converted with P c2w P^T, P = diag(1,-1,-1,1)), undistorted_images /
                             undistorted_depths (mm); the test split is prefixed with the first training frame
`FrameSource.prefetch` decodes ahead on a worker thread and (on CUDA) stages through pinned memory on a copy stream,
which is what removes the reference's per-iteration synchronous image decode from the mapping loop (:2583).
"""
from __future__ import annotations

import glob
import os
import queue
import re
import threading

import numpy as np
import torch

from . import synthetic

_TUM_MAX_DT = 0.08          # tum.py:50
_TUM_FRAME_RATE = 32        # tum.py:80


def _natural_key(path):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", os.path.basename(path))]


def _cv2():
    try:
        import cv2
    except ImportError as e:                     # pragma: no cover
        raise ImportError("on-disk frame sources need OpenCV (cv2.resize, as the reference uses)") from e
    return cv2


def _read_image(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im)


def quat_xyzw_to_matrix(q):
    """Rotation matrix of a (qx, qy, qz, qw) quaternion, normalised first (scipy Rotation.from_quat semantics,
    which tum.py:69-76 relies on)."""
    x, y, z, w = np.asarray(q, np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class FrameSource:
    """Base: subclasses provide `_paths()` -> (colour paths, depth paths) and `_poses()` -> list of 4x4 c2w,
    both for the WHOLE sequence; slicing, pose normalisation, resizing and scaling live here."""

    def __init__(self, camera_params, desired_height=None, desired_width=None, start=0, end=-1, stride=1, relative_pose=True):
        cp = camera_params
        self.orig_h, self.orig_w = int(cp["image_height"]), int(cp["image_width"])
        self.fx, self.fy, self.cx, self.cy = float(cp["fx"]), float(cp["fy"]), float(cp["cx"]), float(cp["cy"])
        self.png_depth_scale = float(cp["png_depth_scale"])
        self.H = int(desired_height or self.orig_h)
        self.W = int(desired_width or self.orig_w)
        if start < 0:
            raise ValueError(f"start must be non-negative, got {start}")
        if not (end == -1 or end > start):
            raise ValueError(f"end ({end}) must be -1 (all frames) or greater than start ({start})")
        colour, depth = self._paths()
        if len(colour) != len(depth):
            raise ValueError("Number of color and depth images must be the same.")
        poses = self._poses(len(colour))
        stop = len(colour) if end == -1 else end
        sl = slice(start, stop, stride or 1)
        self.colour_paths, self.depth_paths = colour[sl], depth[sl]
        self.retained = list(range(len(colour)))[sl]
        P = torch.stack([torch.as_tensor(np.asarray(p), dtype=torch.float32) for p in poses[sl]]) if len(poses[sl]) else torch.zeros((0, 4, 4))
        if relative_pose and P.shape[0]:
            P = torch.inverse(P[0]).unsqueeze(0).expand_as(P) @ P          # basedataset.py:274-292
        self.c2w = P
        # scale_intrinsics in float32, as the reference does
        K = np.eye(3, dtype=np.float32)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2] = self.fx, self.fy, self.cx, self.cy
        K[0, 0] *= np.float32(self.W / self.orig_w)
        K[0, 2] *= np.float32(self.W / self.orig_w)
        K[1, 1] *= np.float32(self.H / self.orig_h)
        K[1, 2] *= np.float32(self.H / self.orig_h)
        self.K = K

    # -- to be provided ------------------------------------------------------------------------------------
    def _paths(self):
        raise NotImplementedError

    def _poses(self, n):
        raise NotImplementedError

    # -- access --------------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.colour_paths)

    def _decode(self, i):
        cv2 = _cv2()
        colour = np.asarray(_read_image(self.colour_paths[i]), dtype=float)
        colour = cv2.resize(colour, (self.W, self.H), interpolation=cv2.INTER_LINEAR)
        depth = np.asarray(_read_image(self.depth_paths[i]), dtype=np.int64).astype(float)
        depth = cv2.resize(depth, (self.W, self.H), interpolation=cv2.INTER_NEAREST) / self.png_depth_scale
        im = torch.from_numpy(colour).to(torch.float32).permute(2, 0, 1) / 255
        return im.contiguous(), torch.from_numpy(depth).to(torch.float32)[None].contiguous()

    def decode_raw(self, i):
        """The frame as decoded from disk, before any arithmetic: (colour uint8 [h,w,3], depth uint16 [h,w]) or None when
        the files hold other types (then only the CPU path applies)."""
        colour, depth = _read_image(self.colour_paths[i]), _read_image(self.depth_paths[i])
        if colour.dtype != np.uint8 or colour.ndim != 3 or colour.shape[2] != 3 or depth.dtype != np.uint16 or depth.ndim != 2:
            return None
        return np.array(colour), np.array(depth)          # writable copies (Pillow hands out read-only views)

    def convert_on_device(self, colour_u8, depth_u16, device, stream=None):
        """Raw frame (host arrays or device tensors) -> (im[3,H,W], depth[1,H,W]) float32 on `device` in one kernel of
        libvtgs_cuda.so (vtgs_frame_convert): resize + scaling follow the CPU loader (`_decode`), but the frame crosses
        PCIe as 5 bytes per source pixel instead of 16 per target pixel."""
        import ctypes as C
        from . import _lib
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("convert_on_device needs a CUDA device (the CPU path is __getitem__)")
        up = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(a).pin_memory()).to(device, non_blocking=True)
        c, d = up(colour_u8), up(depth_u16 if isinstance(depth_u16, torch.Tensor) else depth_u16.view(np.int16))
        sh, sw = int(c.shape[0]), int(c.shape[1])
        im = torch.empty((3, self.H, self.W), dtype=torch.float32, device=device)
        depth = torch.empty((1, self.H, self.W), dtype=torch.float32, device=device)
        st = stream if stream is not None else torch.cuda.current_stream(device)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().vtgs_frame_convert(sw, sh, self.W, self.H, C.c_void_p(c.data_ptr()), C.c_void_p(d.data_ptr()),
                                                     float(self.png_depth_scale), C.c_void_p(im.data_ptr()),
                                                     C.c_void_p(depth.data_ptr()), C.c_void_p(st.cuda_stream)))
        for t in (c, d):
            t.record_stream(st)
        return im, depth

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        im, depth = self._decode(i)
        return dict(im=im, depth=depth, K=torch.from_numpy(self.K.copy()), c2w=self.c2w[i].clone(), index=self.retained[i],
                    W=self.W, H=self.H)

    def prefetch(self, device="cpu", ahead=2, indices=None, device_convert=False):
        """Iterate frames with decoding `ahead` frames in advance on a worker thread; on a CUDA device the planes
        arrive there through pinned staging buffers on a copy stream (the consumer's stream waits on the copy).
        device_convert: upload the raw decoded bytes and resize / scale on the GPU (convert_on_device) instead of
        converting on the CPU; frames whose files are not uint8 colour + uint16 depth take the CPU path."""
        device = torch.device(device)
        order = list(range(len(self))) if indices is None else list(indices)
        q = queue.Queue(maxsize=max(1, ahead))
        stop = threading.Event()
        cuda = device.type == "cuda"
        stream = torch.cuda.Stream(device) if cuda else None

        def work():
            try:
                for i in order:
                    if stop.is_set():
                        return
                    raw = self.decode_raw(i) if (cuda and device_convert and hasattr(self, "png_depth_scale")) else None
                    if raw is not None:
                        with torch.cuda.stream(stream):
                            im, dep = self.convert_on_device(raw[0], raw[1], device, stream)
                            ev = torch.cuda.Event()
                            ev.record(stream)
                        q.put(dict(im=im, depth=dep, K=torch.from_numpy(self.K.copy()), c2w=self.c2w[i].clone(), index=self.retained[i],
                                   W=self.W, H=self.H, _ready=ev))
                        continue
                    fr = self[i]
                    if cuda:
                        with torch.cuda.stream(stream):
                            for k in ("im", "depth"):
                                fr[k] = fr[k].pin_memory().to(device, non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(stream)
                        fr["_ready"] = ev
                    q.put(fr)
                q.put(None)
            except BaseException as e:          # surface decode errors in the consumer
                q.put(e)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        try:
            while True:
                fr = q.get()
                if fr is None:
                    return
                if isinstance(fr, BaseException):
                    raise fr
                ev = fr.pop("_ready", None)
                if ev is not None:
                    cur = torch.cuda.current_stream(device)
                    cur.wait_event(ev)
                    for k in ("im", "depth"):           # allocated on the copy stream, consumed on this one
                        fr[k].record_stream(cur)
                yield fr
        finally:
            stop.set()
            while th.is_alive():                # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    th.join(timeout=0.05)


class ReplicaSource(FrameSource):
    """<basedir>/<sequence>/results/{frame*.jpg, depth*.png}, <basedir>/<sequence>/traj.txt (replica.py)."""

    def __init__(self, camera_params, basedir, sequence, **kw):
        self.folder = os.path.join(basedir, sequence)
        super().__init__(camera_params, **kw)

    def _paths(self):
        colour = sorted(glob.glob(os.path.join(self.folder, "results", "frame*.jpg")), key=_natural_key)
        depth = sorted(glob.glob(os.path.join(self.folder, "results", "depth*.png")), key=_natural_key)
        return colour, depth

    def _poses(self, n):
        with open(os.path.join(self.folder, "traj.txt")) as f:
            lines = f.readlines()
        return [np.array(lines[i].split(), dtype=np.float64).reshape(4, 4) for i in range(n)]


class TumSource(FrameSource):
    """TUM RGB-D layout: rgb.txt, depth.txt, groundtruth.txt (or pose.txt) with timestamped entries (tum.py)."""

    def __init__(self, camera_params, basedir, sequence, **kw):
        self.folder = os.path.join(basedir, sequence)
        self._assoc = None
        super().__init__(camera_params, **kw)

    @staticmethod
    def _table(path, skiprows=0):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)          # "input line contained no data" for comment lines
            return np.loadtxt(path, delimiter=" ", dtype=np.str_, skiprows=skiprows, ndmin=2)

    def _associate(self):
        if self._assoc is not None:
            return self._assoc
        pose_file = next(p for p in (os.path.join(self.folder, n) for n in ("groundtruth.txt", "pose.txt")) if os.path.isfile(p))
        rgb, dep, pose = self._table(os.path.join(self.folder, "rgb.txt")), self._table(os.path.join(self.folder, "depth.txt")), self._table(pose_file, 1)
        t_rgb, t_dep, t_pose = (a[:, 0].astype(np.float64) for a in (rgb, dep, pose))
        pairs = []
        for i, t in enumerate(t_rgb):                   # nearest depth and pose stamps, both within max_dt
            j, k = int(np.argmin(np.abs(t_dep - t))), int(np.argmin(np.abs(t_pose - t)))
            if abs(t_dep[j] - t) < _TUM_MAX_DT and abs(t_pose[k] - t) < _TUM_MAX_DT:
                pairs.append((i, j, k))
        keep = [0] if pairs else []
        for n in range(1, len(pairs)):                  # drop frames closer than 1 / frame_rate to the last kept one
            if t_rgb[pairs[n][0]] - t_rgb[pairs[keep[-1]][0]] > 1.0 / _TUM_FRAME_RATE:
                keep.append(n)
        self._assoc = ([pairs[n] for n in keep], rgb, dep, pose[:, 1:].astype(np.float64))
        return self._assoc

    def _paths(self):
        pairs, rgb, dep, _ = self._associate()
        return ([os.path.join(self.folder, rgb[i, 1]) for i, _, _ in pairs],
                [os.path.join(self.folder, dep[j, 1]) for _, j, _ in pairs])

    def _poses(self, n):
        pairs, _, _, vec = self._associate()
        out = []
        for _, _, k in pairs:
            M = np.eye(4)
            M[:3, :3] = quat_xyzw_to_matrix(vec[k, 3:7])
            M[:3, 3] = vec[k, :3]
            out.append(M)
        return out


class ScannetSource(FrameSource):
    """ScanNet (v2) export layout (scannet.py): <basedir>/<sequence>/color/*.jpg, depth/*.png, pose/*.txt (one 4x4 c2w
    per frame), all in natural order."""

    def __init__(self, camera_params, basedir, sequence, desired_height=968, desired_width=1296, **kw):
        self.folder = os.path.join(basedir, sequence)
        super().__init__(camera_params, desired_height=desired_height, desired_width=desired_width, **kw)

    def _paths(self):
        return (sorted(glob.glob(os.path.join(self.folder, "color", "*.jpg")), key=_natural_key),
                sorted(glob.glob(os.path.join(self.folder, "depth", "*.png")), key=_natural_key))

    def _poses(self, n):
        return [np.loadtxt(p) for p in sorted(glob.glob(os.path.join(self.folder, "pose", "*.txt")), key=_natural_key)]


class ScannetPPSource(FrameSource):
    """ScanNet++ DSLR layout (scannetpp.py): camera parameters come from the sequence's own metadata."""

    def __init__(self, basedir, sequence, ignore_bad=False, use_train_split=True, desired_height=1168, desired_width=1752, **kw):
        import json
        self.folder = os.path.join(basedir, sequence)
        with open(os.path.join(self.folder, "dslr", "train_test_lists.json")) as f:
            split = json.load(f)
        with open(os.path.join(self.folder, "dslr", "nerfstudio", "transforms_undistorted.json")) as f:
            meta = json.load(f)
        self._train_names = split["train"]
        self._names = split["train"] if use_train_split else split["test"]
        self._frames = meta["frames"] if use_train_split else meta["test_frames"]
        self._train_frames = meta["frames"]
        self._ignore_bad, self._use_train = ignore_bad, use_train_split
        cam = dict(png_depth_scale=1000.0, image_height=meta["h"], image_width=meta["w"], fx=meta["fl_x"], fy=meta["fl_y"],
                   cx=meta["cx"], cy=meta["cy"])
        self._listing = None
        super().__init__(cam, desired_height=desired_height, desired_width=desired_width, **kw)

    def _list(self):
        if self._listing is not None:
            return self._listing
        base = os.path.join(self.folder, "dslr")
        flip = np.diag([1.0, -1.0, -1.0, 1.0]).astype(np.float32)
        by_name = {fr["file_path"]: fr for fr in self._frames}
        entries = []
        if not self._use_train:                        # evaluation split: anchored on the first training frame
            first = self._train_names[0]
            entries.append((first, {fr["file_path"]: fr for fr in self._train_frames}[first]))
        for name in self._names:
            fr = by_name[name]
            if self._ignore_bad and fr["is_bad"]:
                continue
            entries.append((name, fr))
        colour = [os.path.join(base, "undistorted_images", n) for n, _ in entries]
        depth = [os.path.join(base, "undistorted_depths", n.replace(".JPG", ".png")) for n, _ in entries]
        poses = [flip @ np.asarray(fr["transform_matrix"], np.float32) @ flip.T for _, fr in entries]
        self._listing = (colour, depth, poses)
        return self._listing

    def _paths(self):
        return self._list()[0], self._list()[1]

    def _poses(self, n):
        return self._list()[2]


class SyntheticSource(FrameSource):
    """The analytic room of `synthetic` along `synthetic.trajectory`: no files, same interface."""

    def __init__(self, shape="tum_fr1", num_frames=60, width=None, height=None, step_m=0.01, step_deg=0.3, seed=0,
                 start=0, end=-1, stride=1):
        self.shape, self.seed = shape, seed
        self.W, self.H, K = synthetic.intrinsics(shape, width, height)
        self.K = K.astype(np.float32)
        poses = synthetic.trajectory(num_frames, step_m, step_deg, seed=seed + 3)
        stop = num_frames if end == -1 else end
        self.retained = list(range(num_frames))[slice(start, stop, stride or 1)]
        P = torch.as_tensor(poses[self.retained], dtype=torch.float32)
        self._gen_poses = poses
        self.c2w = torch.inverse(P[0]).unsqueeze(0).expand_as(P) @ P if len(self.retained) else P
        self.colour_paths = self.depth_paths = self.retained
        self._size = (width, height)

    def _decode(self, i):
        j = self.retained[i]
        # rendered from the pose RELATIVE to the first retained frame, like a dataset whose poses were normalised
        rel = np.linalg.inv(self._gen_poses[self.retained[0]]) @ self._gen_poses[j]
        fr = synthetic.make_frame(self.shape, self._size[0], self._size[1], seed=self.seed + j, c2w=rel)
        return torch.from_numpy(fr["im"]), torch.from_numpy(fr["depth"])
