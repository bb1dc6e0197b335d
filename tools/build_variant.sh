#!/bin/bash
# usage: [SRC=fk] tools/build_variant.sh <name> [-D... flags]   -> build/variants/libirt_<name>.so ($SRC.cu rebuilt with the flags)
set -e
name=$1; shift
C=interactive-rate-tendons_b200/csrc
mkdir -p build/variants build/obj_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin g++"
SRC=${SRC:-fk}
$NV "$@" -c -o build/obj_$name/$SRC.o $C/$SRC.cu -Xptxas -v 2> build/obj_$name/$SRC.ptxas.log
objs=""
for f in ctx fk fk_api selfcol voxel_check voxel_raster env_prep jacobian rmp_io; do [ $f = $SRC ] || objs="$objs $C/$f.o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o build/variants/libirt_$name.so build/obj_$name/$SRC.o $objs -ccbin g++
[ $SRC != fk ] || grep -A2 "fk_rk4_fp64_kernelILi6ELb1" build/obj_$name/fk.ptxas.log | grep -E "spill|registers" | head -4
