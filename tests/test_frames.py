"""CPU: vtgaussian_slam_b200.frames against golden vectors produced by the REFERENCE's own dataset classes
(tests/golden/make_frames_golden.py: ReplicaDataset / TUMDataset of datasets/gradslam_datasets on the committed
fixture sequences, followed by the two lines of the reference's main loop that bring a frame into render layout)."""
import os

import numpy as np
import pytest
import torch

from vtgaussian_slam_b200 import frames

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "frames_fixture")
G = np.load(os.path.join(HERE, "golden", "frames_golden.npz"))

REPLICA_CAM = dict(image_height=48, image_width=64, fx=32.0, fy=32.0, cx=31.5, cy=23.5, png_depth_scale=6553.5)
TUM_CAM = dict(image_height=48, image_width=64, fx=51.73, fy=51.65, cx=31.86, cy=25.53, png_depth_scale=5000.0)

SCANNET_CAM = dict(image_height=48, image_width=64, fx=57.7, fy=57.9, cx=31.9, cy=23.8, png_depth_scale=1000.0)

CASES = {
    "scannet_native": lambda: frames.ScannetSource(SCANNET_CAM, os.path.join(FIX, "scannet"), "scene0000_00", desired_height=48, desired_width=64),
    "scannet_resized": lambda: frames.ScannetSource(SCANNET_CAM, os.path.join(FIX, "scannet"), "scene0000_00", desired_height=33, desired_width=50,
                                                    start=2, stride=3),
    "replica_native": lambda: frames.ReplicaSource(REPLICA_CAM, os.path.join(FIX, "replica"), "room0", desired_height=48, desired_width=64),
    "replica_resized": lambda: frames.ReplicaSource(REPLICA_CAM, os.path.join(FIX, "replica"), "room0", desired_height=30, desired_width=44,
                                                    start=1, end=6, stride=2),
    "tum_native": lambda: frames.TumSource(TUM_CAM, os.path.join(FIX, "tum"), "fr1", desired_height=48, desired_width=64),
    "tum_resized": lambda: frames.TumSource(TUM_CAM, os.path.join(FIX, "tum"), "fr1", desired_height=24, desired_width=32, start=1),
    "scannetpp_train": lambda: frames.ScannetPPSource(os.path.join(FIX, "scannetpp"), "scene0", desired_height=48, desired_width=72),
    "scannetpp_train_nobad": lambda: frames.ScannetPPSource(os.path.join(FIX, "scannetpp"), "scene0", desired_height=24, desired_width=36,
                                                            ignore_bad=True, start=1),
    "scannetpp_test": lambda: frames.ScannetPPSource(os.path.join(FIX, "scannetpp"), "scene0", desired_height=48, desired_width=72,
                                                     use_train_split=False),
}


@pytest.mark.parametrize("tag", list(CASES))
def test_sources_reproduce_the_reference_datasets(tag):
    src = CASES[tag]()
    assert len(src) == G[f"{tag}.im"].shape[0]
    assert [os.path.relpath(p, FIX) for p in src.colour_paths] == list(G[f"{tag}.files"])        # same frames, same order
    for i in range(len(src)):
        fr = src[i]
        assert fr["im"].dtype == torch.float32 and fr["im"].shape == G[f"{tag}.im"][i].shape
        # same decoder, same cv2 resize on float64, same scaling: bit-identical planes
        assert np.array_equal(fr["im"].numpy(), G[f"{tag}.im"][i])
        assert np.array_equal(fr["depth"].numpy(), G[f"{tag}.depth"][i])
        assert np.array_equal(fr["K"].numpy(), G[f"{tag}.K"][i])
        assert np.allclose(fr["c2w"].numpy(), G[f"{tag}.c2w"][i], atol=2e-6)
    assert np.allclose(src.c2w[0].numpy(), np.eye(4), atol=2e-6)                                 # poses relative to the first retained frame


def test_tum_association_drops_unposed_and_too_dense_frames():
    src = CASES["tum_native"]()
    names = [os.path.basename(p) for p in src.colour_paths]
    assert "100.600000.png" not in names             # no pose within 0.08 s
    assert "100.020000.png" not in names and "100.150000.png" not in names        # < 1/32 s after the previous kept frame
    assert names[0] == "100.000000.png" and len(names) == 5


def test_argument_checks_match_the_reference():
    with pytest.raises(ValueError):
        frames.ReplicaSource(REPLICA_CAM, os.path.join(FIX, "replica"), "room0", start=-1)
    with pytest.raises(ValueError):
        frames.ReplicaSource(REPLICA_CAM, os.path.join(FIX, "replica"), "room0", start=3, end=2)
    src = CASES["replica_native"]()
    with pytest.raises(IndexError):
        src[len(src)]
    assert src[-1]["index"] == len(src) - 1


def test_prefetch_yields_every_frame_in_order_and_surfaces_errors():
    src = CASES["replica_native"]()
    got = [fr["index"] for fr in src.prefetch("cpu", ahead=2)]
    assert got == list(range(len(src)))
    sub = [fr["index"] for fr in src.prefetch("cpu", ahead=1, indices=[4, 1, 3])]
    assert sub == [4, 1, 3]
    it = src.prefetch("cpu", ahead=1)
    next(it)
    it.close()                                       # early exit must not hang on the worker thread
    src.colour_paths[2] = os.path.join(FIX, "missing.jpg")
    with pytest.raises(FileNotFoundError):
        list(src.prefetch("cpu", ahead=2))


def test_synthetic_source_has_the_same_interface():
    src = frames.SyntheticSource("tum_fr1", num_frames=5, width=64, height=48, start=1)
    assert len(src) == 4
    fr = src[0]
    assert fr["im"].shape == (3, 48, 64) and fr["depth"].shape == (1, 48, 64) and fr["K"].shape == (3, 3)
    assert np.allclose(fr["c2w"].numpy(), np.eye(4), atol=1e-6)
    assert [f["index"] for f in src.prefetch("cpu")] == [1, 2, 3, 4]
    assert float(fr["depth"].min()) > 0.2 and 0.0 <= float(fr["im"].min()) and float(fr["im"].max()) <= 1.0


def test_quaternion_rows_are_normalised_like_scipy():
    q = np.array([0.1, -0.3, 0.2, 0.9]) * 1.7
    R = frames.quat_xyzw_to_matrix(q)
    from scipy.spatial.transform import Rotation
    assert np.allclose(R, Rotation.from_quat(q).as_matrix(), atol=1e-12)


def test_scannet_frames_come_in_natural_order():
    src = CASES["scannet_native"]()
    assert [os.path.basename(p) for p in src.colour_paths] == [f"{i}.jpg" for i in range(11)]      # not 0, 1, 10, 2, ...
    raw = src.decode_raw(3)
    assert raw[0].dtype == np.uint8 and raw[0].shape == (48, 64, 3) and raw[1].dtype == np.uint16 and raw[1].shape == (48, 64)
    with pytest.raises(RuntimeError):
        src.convert_on_device(raw[0], raw[1], "cpu")          # the device path has no CPU fallback (the CPU path is src[i])
