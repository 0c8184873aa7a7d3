"""Golden vectors for vtgaussian_slam_b200.metrics from the reference's own functions.

    python tests/golden/make_metrics_golden.py          # build container only (needs /root/reference)

utils/eval_helpers.py cannot be imported here (matplotlib, lpips, open3d, kornia ... are absent), so the two pure
numpy functions `align` and `evaluate_ate` are compiled from their own source text (ast) and executed unchanged;
`calc_psnr` is imported from utils/slam_external.py.  `np.linalg.linalg` (removed in NumPy 2) is aliased to np.linalg.
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_functions():
    src = open("/root/reference/utils/eval_helpers.py").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("align", "evaluate_ate")]
    if not hasattr(np.linalg, "linalg"):
        np.linalg.linalg = np.linalg
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "eval_helpers.py", "exec"), ns)
    sys.path.insert(0, "/root/reference")
    from utils.slam_external import calc_psnr
    return ns["align"], ns["evaluate_ate"], calc_psnr


if __name__ == "__main__":
    align, evaluate_ate, calc_psnr = reference_functions()
    rng = np.random.default_rng(0)
    out = {}
    for k, n in enumerate((5, 40, 300)):
        gt = [torch.eye(4) for _ in range(n)]
        est = [torch.eye(4) for _ in range(n)]
        path = np.cumsum(rng.normal(0, 0.05, (n, 3)), 0)
        # the estimate: the same path in another rigid frame (one case with a reflection-prone, nearly planar path) + noise
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        if np.linalg.det(A) < 0:
            A[:, 0] *= -1
        if k == 0:
            path[:, 2] *= 1e-3
        noisy = path @ A.T + rng.normal(0, 0.004, (n, 3)) + np.array([0.3, -1.0, 2.0])
        for i in range(n):
            gt[i][:3, 3] = torch.tensor(path[i], dtype=torch.float32)
            est[i][:3, 3] = torch.tensor(noisy[i], dtype=torch.float32)
        out[f"ate{k}.gt"] = torch.stack(gt).numpy()
        out[f"ate{k}.est"] = torch.stack(est).numpy()
        out[f"ate{k}.value"] = np.float64(evaluate_ate(gt, est))
        R, t, err = align(out[f"ate{k}.gt"][:, :3, 3].T.astype(np.float64), out[f"ate{k}.est"][:, :3, 3].T.astype(np.float64))
        out[f"ate{k}.R"], out[f"ate{k}.t"], out[f"ate{k}.err"] = np.asarray(R), np.asarray(t), np.asarray(err)
    a, b = torch.rand(3, 24, 32, generator=torch.Generator().manual_seed(1)), torch.rand(3, 24, 32, generator=torch.Generator().manual_seed(2))
    out["psnr.a"], out["psnr.b"], out["psnr.value"] = a.numpy(), b.numpy(), calc_psnr(a, b).numpy()
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("value")})
