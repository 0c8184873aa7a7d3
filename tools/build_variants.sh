#!/bin/bash
# Kernel experiments: build named variants of libvtgs_cuda.so with different compile-time constants.
#   tools/build_variants.sh name1 "-DVTGS_FWD_GC=4" name2 "-DVTGS_BWD_GC=3" ...
# Run one with:  VTGS_LIB_PATH=vtgaussian_slam_b200/lib/variants/libvtgs_<name>.so python bench.py --kernels-only
set -e
cd "$(dirname "$0")/../vtgaussian_slam_b200/csrc"
while [ $# -ge 2 ]; do
  make -s OUT=../lib/variants/libvtgs_$1.so BUILD=build_$1 EXTRA="$2" > /dev/null
  echo "built variant $1 ($2)"
  shift 2
done
