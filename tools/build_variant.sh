#!/bin/bash
# A/B builds of the library: tools/build_variant.sh NAME "-DSWB_PASS_GROUP=8u -DSWB_BLOCK_CHUNKS=8u" -> PKG/lib_NAME/libswb.so
# (run with SWB_LIB=.../lib_NAME/libswb.so; see tools/sweep.py). Only the CUDA objects depend on the macros.
set -e
PKG="$(cd "$(dirname "$0")/.." && pwd)/ece1782-smith-waterman-cuda_b200"
NAME=$1; FLAGS=$2
mkdir -p $PKG/lib_$NAME
make -C $PKG -s lib/libswb.so
ARCH="-gencode arch=compute_100a,code=sm_100a"
NV="$ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $FLAGS"
nvcc $NV -c $PKG/csrc/swb_kernels.cu -o $PKG/lib_$NAME/swb_kernels.o &
nvcc $NV -c $PKG/csrc/swb_engine.cu -o $PKG/lib_$NAME/swb_engine.o &
wait
nvcc $ARCH -shared -o $PKG/lib_$NAME/libswb.so $PKG/lib_$NAME/swb_kernels.o $PKG/lib_$NAME/swb_engine.o $PKG/lib/swb_group.o $PKG/lib/swb_plan.o $PKG/lib/swb_scoring.o $PKG/lib/swb_dbfile.o $PKG/lib/swb_microbench.o
rm -f $PKG/lib_$NAME/*.o
ls -la $PKG/lib_$NAME/libswb.so
