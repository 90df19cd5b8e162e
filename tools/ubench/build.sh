#!/bin/bash
# builds build/ubench_lat and build/raylat[_<tag>] (extra nvcc flags select loop variants of ray.cuh):  tools/ubench/build.sh [tag "flags"]...
set -e
cd "$(dirname "$0")/../.."
PKG=sparse-voxel-octree-raycasting_b200
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Iinclude"
mkdir -p build
$NV -o build/ubench_lat tools/ubench/lat.cu
one() {
  tag=$1; flags=$2
  $NV $flags -DCOUNT_BUILD -c -o build/raylat_count$tag.o tools/ubench/raylat.cu
  $NV $flags -o build/raylat$tag tools/ubench/raylat.cu build/raylat_count$tag.o -L$PKG -lsvo_b200 -Xlinker -rpath,'$ORIGIN/../'$PKG
}
one "" ""
while [ $# -ge 2 ]; do one "_$1" "$2"; shift 2; done
