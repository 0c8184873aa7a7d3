#!/bin/bash
# bench every built variant (and the default library) with --kernels-only; one JSON line each
cd "$(dirname "$0")/.."
STEPS=${STEPS:-60}
python bench.py --steps $STEPS --warmup 5 --kernels-only 2>/dev/null | tail -1
for so in vtgaussian_slam_b200/lib/variants/*.so; do
  VTGS_LIB_PATH=$PWD/$so python bench.py --steps $STEPS --warmup 5 --kernels-only 2>/dev/null | tail -1
done
