"""A CPU `Renderer` with the rasteriser's autograd contract whose arithmetic is the oracle (TEST INFRASTRUCTURE).

Used by tests/golden/make_get_loss_golden.py to execute the reference's own `get_loss` without its (absent)
rasteriser, and by tests/test_get_loss_golden.py to run `slam_ops.get_loss(backend="dropin")` on the CPU with the
same render, so that the comparison isolates the host logic around the two render calls."""
import numpy as np
import torch

import oracle


class _OracleRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, opacities, scales, rotations, cam):
        o = oracle.Oracle()
        out = o.forward(cam, means3D.detach().numpy(), scales.detach().numpy(), rotations.detach().numpy(),
                        opacities.detach().numpy().reshape(-1), colors_precomp.detach().numpy())
        ctx.o = o
        ctx.op_shape = opacities.shape
        color = torch.from_numpy(out["color"].copy())
        radii = torch.from_numpy(out["radii"].copy())
        depth = torch.from_numpy(out["depth"].copy())[None]
        ctx.mark_non_differentiable(radii, depth)
        return color, radii, depth

    @staticmethod
    def backward(ctx, g_color, _g_radii, _g_depth):
        g = ctx.o.backward(g_color.contiguous().numpy())
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        return (t(g["means3D"]), t(g["means2D"]), t(g["colors"]), t(g["opacities"]).reshape(ctx.op_shape),
                t(g["scales"]), t(g["rotations"]), None)


class OracleRenderer:
    """Renderer(raster_settings=cam)(**rendervar) -> (color[3,H,W], radii[N], depth[1,H,W]); `cam` is an oracle.Camera."""

    def __init__(self, raster_settings):
        self.cam = raster_settings

    def __call__(self, means3D, means2D, opacities, colors_precomp, scales, rotations):
        return _OracleRasterize.apply(means3D, means2D, colors_precomp, opacities, scales, rotations, self.cam)
