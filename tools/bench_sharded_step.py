"""torchrun --nproc-per-node N tools/bench_sharded_step.py : the keyframe-sharded mapping step alone (barrier +
vtgs_sharded_adam + barrier on the C2-sized flat vectors) for a few grid sizes, peer loads / stores against the NVLS multicast
path, next to the NCCL all-reduce of the same message + the replicated Adam it replaces.  Max over ranks, microseconds."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vtgaussian_slam_b200.fused import MappingSolver, adam_step  # noqa: E402

rank, world, local = bench.dist_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
pg = dist.group.WORLD
wl = bench.build_workload(os.environ.get("WL", "c2"))
params = {k: torch.tensor(v, device=dev) for k, v in wl["params"].items()}
ms = MappingSolver(bench.make_settings(wl, dev), params, device=dev, process_group=pg, sharded_step=True)
mc_ptr = int(getattr(ms.sharded["hdl"], "multicast_ptr", 0) or 0)
ms.step_dev.fill_(1)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


rows = []
for mc in ([0, mc_ptr] if mc_ptr else [0]):
    ms.sharded["mc"] = mc
    for blocks in (74, 148, 296, 592, 1184):
        os.environ["VTGS_SHARD_BLOCKS"] = str(blocks)
        rows.append((("nvls" if mc else "p2p"), blocks, timed(ms._sharded_step)))
n = ms.sharded["n"]
flat = torch.zeros(n + 1, device=dev)
p, m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
nccl = timed(lambda: dist.all_reduce(flat, group=pg))
adam = timed(lambda: adam_step(p, flat[:n], m, v, 1e-3, step_dev=ms.step_dev, eps=1e-15))
bar = timed(lambda: ms.sharded["hdl"].barrier(channel=0))
if rank == 0:
    print(f"world {world}, flat vector {4 * n / 1e6:.1f} MB, multicast {'available' if mc_ptr else 'unavailable'}")
    for r in rows:
        print(f"  fused step {r[0]:5s} blocks {r[1]:5d}: {r[2]:7.1f} us")
    print(f"  NCCL all-reduce {nccl:7.1f} us + replicated Adam {adam:6.1f} us = {nccl + adam:7.1f} us;  one barrier {bar:5.1f} us")
dist.barrier()
os._exit(0)
