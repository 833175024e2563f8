#!/bin/bash
# build an experimental variant of libfruitnerf_b200.so into tools/bin/ (same ABI; select it with FNR_LIB=...):
#   tools/r2/build_variant.sh NAME file.cu "-DFLAG ..."   -- recompiles only file.cu with the extra flags, links with the in-tree objects
set -e
NAME=$1; SRC=$2; FLAGS=$3
cd "$(dirname "$0")/../.."
mkdir -p tools/bin
C=fruitnerf_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -O2 $FLAGS -c $C/$SRC -o tools/bin/${NAME}_${SRC%.cu}.o
OBJS=$(ls $C/*.o | grep -v "/${SRC%.cu}.o")
nvcc -shared -o tools/bin/libfnr_${NAME}.so $OBJS tools/bin/${NAME}_${SRC%.cu}.o -lcudart
echo tools/bin/libfnr_${NAME}.so
