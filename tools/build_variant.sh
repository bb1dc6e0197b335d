#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFK_... flags]   -> build/variants/libirt_<name>.so (fk.cu rebuilt with the flags)
set -e
name=$1; shift
C=interactive-rate-tendons_b200/csrc
mkdir -p build/variants build/obj_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin g++"
$NV "$@" -c -o build/obj_$name/fk.o $C/fk.cu -Xptxas -v 2> build/obj_$name/fk.ptxas.log
objs=""
for f in ctx fk_api selfcol voxel_check voxel_raster env_prep jacobian rmp_io; do objs="$objs $C/$f.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o build/variants/libirt_$name.so build/obj_$name/fk.o $objs -ccbin g++
grep -A2 "fk_rk4_fp64_kernelILi6ELb1" build/obj_$name/fk.ptxas.log | grep -E "spill|registers" | head -4
