"""Read the JSON lines of tools/run_variants.sh from stdin -> one row per library: iterations/s and the blend kernels' microseconds."""
import json
import sys

for line in sys.stdin:
    try:
        d = json.loads(line)
    except Exception:
        continue
    k = d["per_kernel_us"]
    print(f'{d["lib"].split("/")[-1]:34s} {d["value"]:8.1f} it/s  K5 {k["blend_forward_kernel"]:7.2f}  K6 {k["blend_backward_kernel"]:7.2f}  '
          f'K7 {k["fused_preprocess_backward_kernel"]:6.2f}  sort {k["tile_sort_kernel"]:6.2f}')
