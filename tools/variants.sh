#!/bin/bash
# build variants of libsvo_b200.so with extra nvcc flags:  tools/variants.sh tag "flags" [tag "flags" ...]
# -> build/variants/libsvo_b200_<tag>.so ; run with SVO_B200_LIB=<path>
set -e
cd "$(dirname "$0")/.."
PKG=sparse-voxel-octree-raycasting_b200
mkdir -p build/variants
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -fopenmp -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
      -Xcompiler -fPIC,-O2 -Xptxas -v $flags -shared -o build/variants/libsvo_b200_$tag.so $PKG/csrc/svo_abi.cu $PKG/csrc/svo_builder.cu $PKG/host/octree_builder.cpp $PKG/host/raycast_host.cpp $PKG/host/scene_io.cpp -Iinclude 2> build/variants/$tag.log
  grep -A1 "k_raycast_fine_2ILi11" build/variants/$tag.log | grep -o "Used [0-9]* registers" | head -1 | sed "s/^/$tag: /"
done
