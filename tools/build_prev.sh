#!/bin/bash
# builds openvdb_b200/variants/libvdbrt_prev.so from the kernels of a commit (default HEAD); the other objects come from the regular build
set -e
cd "$(dirname "$0")/.."
rev=${1:-HEAD}
rm -rf openvdb_b200/variants/prev_src; mkdir -p openvdb_b200/variants/prev_src openvdb_b200/variants/prev
git archive $rev openvdb_b200/csrc include | tar -x -C openvdb_b200/variants/prev_src
(cd openvdb_b200/variants/prev_src/openvdb_b200/csrc && /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --fmad=false -Xcompiler -fPIC,-ffp-contract=off -c vdbrt.cu -o ../../../prev/vdbrt.o)
cd openvdb_b200
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libvdbrt_prev.so variants/prev/vdbrt.o build/vdbrt_build.o build/vdbrt_quant.o build/vdbrt_camera.o build/vdbrt_io.o -cudart static -lz
rm -rf variants/prev_src
