#!/bin/bash
# Tuning builds of the library with different block-shape constants.
# usage: scripts/build_variants.sh name:CHUNK:RMAX:NSTAGE:MINB:XMAX ...
set -e
cd "$(dirname "$0")/.."
out=ip-nonlinear-solver_b200/ipsolver_b200/variants
mkdir -p $out
for spec in "$@"; do
  IFS=: read name chunk rmax nstage minb xmax <<< "$spec"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
     -DIPB_CHUNK=$chunk -DIPB_RMAX=$rmax -DIPB_NSTAGE=$nstage -DIPB_MIN_BLOCKS=$minb -DIPB_XMAX=${xmax:-1152} $EXTRA \
     -o $out/lib_$name.so ip-nonlinear-solver_b200/csrc/api.cu &
done
wait
ls -la $out
