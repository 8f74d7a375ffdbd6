#!/bin/bash
# Development A/B builds: scripts/build_variant.sh <name> "<extra nvcc flags>" -> build/variants/libyuki_<name>.so
# (select at run time with YUKI_GPU_LIB=build/variants/libyuki_<name>.so; build/ is git-ignored but travels with gpurun)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
SRC=yuki_b200/csrc
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  --fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-pthread \
  -Iinclude -I$SRC -shared -o build/variants/libyuki_$name.so \
  $SRC/render.cu $SRC/host_scene.cpp $SRC/host_bvh.cpp $SRC/host_ply.cpp $SRC/host_exr.cpp $SRC/host_pbrt.cpp $SRC/host_mitsuba.cpp $SRC/post.cu -lz "$@" 2>&1 | grep -v "warning\|queue_push\|\^\|^$" || true
echo "built build/variants/libyuki_$name.so"
