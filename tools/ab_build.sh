#!/bin/bash
# usage: tools/ab_build.sh TU TAG "-DFLAG=1 ..."   -> ab/libnqcb200_TAG.so with translation unit TU rebuilt with the extra flags
# (A/B variants of one kernel family for a single gpurun call; swap them in with `cp ab/libnqcb200_TAG.so nqcdynamics.jl_b200/csrc/libnqcb200.so`)
set -e
TU=$1; TAG=$2; FLAGS=$3
cd "$(dirname "$0")/../nqcdynamics.jl_b200/csrc"
mkdir -p ../../ab
OBJS=""
for t in engine tu_density_1d tu_density_spinboson tu_ring tu_classical_nrpmd tu_iesh; do
  if [ "$t" = "$TU" ]; then OBJS="$OBJS ../../ab/${t}_$TAG.o"; else OBJS="$OBJS $t.o"; fi
done
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -Xptxas -v $FLAGS -c -o ../../ab/${TU}_$TAG.o $TU.cu 2> ../../ab/${TU}_$TAG.ptxas
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../ab/libnqcb200_$TAG.so $OBJS -lcudart
