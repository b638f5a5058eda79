#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG=..]...   -> tools/_variants/NAME.so (experiment builds of libvdbrt.so; not shipped)
set -e
cd "$(dirname "$0")/../openvdb_b200/csrc"
name=$1; shift
out=../../tools/_variants
mkdir -p $out/obj_$name
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --fmad=false -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v "$@" -c vdbrt.cu -o $out/obj_$name/vdbrt.o 2> $out/obj_$name/ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $out/obj_$name/vdbrt.o ../build/vdbrt_build.o ../build/vdbrt_quant.o ../build/vdbrt_camera.o ../build/vdbrt_io.o -cudart static -lz
grep -A2 "k_render_levelsetILb0ELb0" $out/obj_$name/ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $name"
