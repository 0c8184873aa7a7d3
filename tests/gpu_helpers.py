"""Helpers for the -m gpu parity tests: run the CUDA path through the C ABI and pull every
stage output back for comparison with the oracle."""
import ctypes as C

import numpy as np
import torch

from vtgaussian_slam_b200 import _lib, rasterizer


def settings_from(s, device):
    return rasterizer.GaussianRasterizationSettings(
        image_height=s["image_height"], image_width=s["image_width"], tanfovx=s["tanfovx"], tanfovy=s["tanfovy"],
        bg=torch.tensor(s["bg"], device=device), scale_modifier=s["scale_modifier"],
        viewmatrix=torch.tensor(s["viewmatrix"], device=device), projmatrix=torch.tensor(s["projmatrix"], device=device),
        sh_degree=0, campos=torch.tensor(s["campos"], device=device), prefiltered=False)


def cuda_forward_all(s, sc, device="cuda:0", tile_rows=(0, 0)):
    """-> dict of numpy stage outputs of the CUDA path (same keys as oracle.Oracle.forward)."""
    dev = torch.device(device)
    cam = rasterizer.camera_struct(settings_from(s, dev), tile_rows=tile_rows)
    t = {k: torch.tensor(v, device=dev) for k, v in sc.items()}
    color, depth, radii, ws = rasterizer.rasterize_forward(cam, t["means3D"], t["scales"], t["rotations"], t["opacities"], t["colors"])
    N = t["means3D"].shape[0]
    R = ws.num_rendered
    keys = torch.zeros(max(R, 1), dtype=torch.int64, device=dev)
    m2d = torch.zeros((max(N, 1), 2), device=dev); dep = torch.zeros(max(N, 1), device=dev); co = torch.zeros((max(N, 1), 4), device=dev)
    b = ws.struct()
    L = _lib.lib()
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(L.vtgs_export_sorted_keys(C.byref(cam), N, C.byref(b), C.c_void_p(keys.data_ptr()), R, st))
    _lib.check(L.vtgs_export_geometry(N, C.byref(b), C.c_void_p(m2d.data_ptr()), C.c_void_p(dep.data_ptr()), C.c_void_p(co.data_ptr()), st))
    torch.cuda.synchronize()
    out = dict(R=R, color=color.cpu().numpy(), depth=depth[0].cpu().numpy(), radii=radii.cpu().numpy(),
               tiles_touched=ws.tiles_touched[:N].cpu().numpy().astype(np.uint32),
               ranges=ws.tile_ranges.cpu().numpy().astype(np.uint32),
               point_list=ws.point_list[:R].cpu().numpy().astype(np.uint32),
               keys=keys[:R].cpu().numpy().astype(np.uint64),
               final_T=ws.final_T.cpu().numpy(), n_contrib=ws.n_contrib.cpu().numpy().astype(np.uint32),
               means2D=m2d[:N].cpu().numpy(), depths=dep[:N].cpu().numpy(), conic_opacity=co[:N].cpu().numpy())
    return out, cam, ws
