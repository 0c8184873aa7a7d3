import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from vtgaussian_slam_b200 import slam_ops
from vtgaussian_slam_b200.rasterizer import GaussianRasterizationSettings
wl = bench.build_workload("c2")
dev = torch.device("cuda:0")
fr, s = wl["frame"], wl["settings"]
settings = GaussianRasterizationSettings(image_height=fr["H"], image_width=fr["W"], tanfovx=s["tanfovx"], tanfovy=s["tanfovy"], bg=torch.tensor(s["bg"], device=dev),
    scale_modifier=1.0, viewmatrix=torch.tensor(s["viewmatrix"], device=dev), projmatrix=torch.tensor(s["projmatrix"], device=dev), sh_degree=0, campos=torch.tensor(s["campos"], device=dev), prefiltered=False)
params = {k: torch.tensor(v, device=dev) for k, v in wl["params"].items()}
N = params["means3D"].shape[0]
P_ = {k: torch.nn.Parameter(v.clone()) for k, v in params.items()}
P_["cam_unnorm_rots"] = torch.nn.Parameter(torch.tensor(wl["q"], device=dev).reshape(1, 4, 1).contiguous())
P_["cam_trans"] = torch.nn.Parameter(torch.tensor(wl["t"], device=dev).reshape(1, 3, 1).contiguous())
lrs = dict(means3D=0.0, rgb_colors=0.0, unnorm_rotations=0.0, logit_opacities=0.0, log_scales=0.0, cam_unnorm_rots=0.0004, cam_trans=0.002)
opt = slam_ops.initialize_optimizer(P_, lrs, tracking=True)
variables = dict(max_2D_radius=torch.zeros(N, device=dev))
data = dict(cam=settings, im=torch.tensor(fr["im"], device=dev), depth=torch.tensor(fr["depth"], device=dev), w2c=torch.eye(4, device=dev))
def step():
    loss, _, _ = slam_ops.get_loss(P_, data, variables, 0, bench.LOSS_W, True, bench.SIL_THRES, True, False, tracking=True, dataset_name="tum", backend="fused")
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
for _ in range(5): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(100): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host ms/step (enqueue only):", (t1 - t0) * 10, "total:", (time.perf_counter() - t0) * 10)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(22); print(st.getvalue()[:5000])
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(30); print(st.getvalue()[:7000])
# the same loop with the device kept idle-free: how long does the DEVICE need per step on this path?
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time as _t
e0.record()
for _ in range(100): step()
e1.record(); torch.cuda.synchronize()
print("device-timed ms/step:", e0.elapsed_time(e1) / 100)
