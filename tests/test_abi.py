"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/vtgs.h declares, struct layouts agree, argument validation works, and the
Python surface mirrors the reference's (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

from vtgaussian_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "vtgs.h")).read()
    declared = set(re.findall(r"VTGS_API\s+[\w\s\*]+?\b(vtgs_\w+)\s*\(", hdr))
    assert len(declared) >= 15
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name)
    assert L.vtgs_abi_version() == _lib.ABI_VERSION == 4
    assert b"sm_100a" in L.vtgs_build_info()


def test_struct_layouts_match_the_header(L):
    assert C.sizeof(_lib.VtgsCamera) == 4 * (2 + 2 + 16 + 16 + 3 + 1 + 1 + 2)
    assert C.sizeof(_lib.VtgsCounters) == 4 * (4 + 9 + 3 + 4 + 2 + 10)
    assert C.sizeof(_lib.VtgsBuffers) == 8 * 19            # 18 pointers / sizes + flags, reserved
    assert C.sizeof(_lib.VtgsParams) == 8 * 5 + 8 + 8
    assert C.sizeof(_lib.VtgsPose) == 16 + 16
    assert C.sizeof(_lib.VtgsLossConfig) == 32 + 8 + 8 + 8
    assert C.sizeof(_lib.VtgsParamGrads) == 88


def test_workspace_query_and_validation(L):
    sz = _lib.VtgsWorkspaceSizes()
    assert L.vtgs_workspace_query(1200, 680, 1000000, 4000000, C.byref(sz)) == 0
    assert (sz.tiles_x, sz.tiles_y) == (75, 43)
    assert sz.geom_bytes == 64 * 1000000 and sz.pair_keys_bytes == 8 * 4000000
    assert sz.final_T_bytes == 1200 * 680 * 4
    assert L.vtgs_workspace_query(0, 680, 1, 1, C.byref(sz)) == -1
    assert b"bad dimensions" in L.vtgs_last_error()
    cam = _lib.VtgsCamera()
    assert L.vtgs_mark_visible(C.byref(cam), 0, None, None, None) == -1     # zero-sized image
    assert L.vtgs_pose_scratch_floats(1000) >= 12 * 4


def test_python_surface_mirrors_the_reference():
    import diff_gaussian_rasterization as dgr
    fields = dgr.GaussianRasterizationSettings._fields
    assert fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                      "projmatrix", "sh_degree", "campos", "prefiltered")       # utils/recon_helpers.py:14-26
    s = dgr.GaussianRasterizationSettings(4, 4, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4)[None], torch.eye(4)[None], 0,
                                          torch.zeros(3), False)
    r = dgr.GaussianRasterizer(raster_settings=s)
    z3, z4, z1 = torch.zeros(2, 3), torch.zeros(2, 4), torch.zeros(2, 1)
    with pytest.raises(Exception, match="excatly one"):
        r(means3D=z3, means2D=z3, opacities=z1, scales=z3, rotations=z4)
    with pytest.raises(NotImplementedError):
        r(means3D=z3, means2D=z3, opacities=z1, shs=torch.zeros(2, 1, 3), scales=z3, rotations=z4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r(means3D=z3, means2D=z3, opacities=z1, colors_precomp=z3, scales=z3, rotations=z4)


def test_product_never_imports_the_oracle():
    for pkg in ("vtgaussian_slam_b200", "diff_gaussian_rasterization"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                    assert "libvtgs_oracle" not in src, f
