#!/bin/bash
# builds openvdb_b200/variants/libvdbrt_<name>.so with extra nvcc flags (A/B experiments; the other objects come from the regular build)
# usage: tools/build_variant.sh <name> [-DFLAG=..]...
set -e
cd "$(dirname "$0")/../openvdb_b200/csrc"
name=$1; shift
mkdir -p ../variants/$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo $ARCH --fmad=false -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v "$@" -c vdbrt.cu -o ../variants/$name/vdbrt.o 2> ../variants/$name/ptxas.log
/usr/local/cuda/bin/nvcc $ARCH -shared -o ../variants/libvdbrt_$name.so ../variants/$name/vdbrt.o ../build/vdbrt_build.o ../build/vdbrt_quant.o ../build/vdbrt_camera.o ../build/vdbrt_io.o -cudart static -lz
grep -A2 "k_render_levelsetILb0ELb0ELb0ELb0ELb0ELi0" ../variants/$name/ptxas.log | grep -E "spill|Used"
