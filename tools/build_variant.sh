#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG=VALUE ...]   -> h264-h265-to-jpeg_b200/lib/variants/<name>.so (A/B measurements:
# H2J_B200_LIB=<that file> python bench.py ...).  lib/ is git-ignored and travels to the GPU box.
# H2J_VARIANT_SRC=<dir> builds another source tree's csrc (e.g. a git worktree of an older commit).
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
SRC=${H2J_VARIANT_SRC:-$HERE/h264-h265-to-jpeg_b200/csrc}
INC=${H2J_VARIANT_INC:-$HERE/include}
mkdir -p $HERE/h264-h265-to-jpeg_b200/lib/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I$INC "$@" -shared \
  -o $HERE/h264-h265-to-jpeg_b200/lib/variants/$NAME.so $SRC/h2j_api.cu $SRC/h2j_host_copy.cpp
