#!/bin/bash
# Builds libgtb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libgtb200.so
FLAGS="${GTB_EXTRA_FLAGS} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr"
SRCS="api.cu prep.cu search_simt.cu refine.cu sparse.cu"
for f in $(ls *.cu); do case " $SRCS " in *" $f "*) ;; *) SRCS="$SRCS $f";; esac; done
mkdir -p ../../build
pids=""
for f in $SRCS; do
  ( nvcc $FLAGS ${GTB_PTXAS_V:+-Xptxas -v} -c $f -o ../../build/${f%.cu}.o ) &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS ../../build/${f%.cu}.o"; done
nvcc -shared -o $OUT $OBJS -lcudart -lcuda
echo "built $(realpath $OUT)"
