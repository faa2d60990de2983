#!/bin/bash
# Kernel experiments: builds a variant of libgtb200.so with extra nvcc flags for search_tc.cu only, next to the in-tree
# library (graphtools_b200/variants/<name>.so; git-ignored, shipped by gpurun).  Use with GTB_LIB=<path>.
#   scripts/build_variant.sh hybrid1 -DGTB_TC_HYBRID=1
set -e
cd "$(dirname "$0")/../graphtools_b200/csrc"
name=$1; shift
mkdir -p ../variants ../../build/var_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr"
nvcc $FLAGS "$@" -c search_tc.cu -o ../../build/var_$name/search_tc.o
OBJS=""
for f in *.cu; do
  if [ "$f" != "search_tc.cu" ]; then OBJS="$OBJS ../../build/${f%.cu}.o"; fi
done
nvcc -shared -o ../variants/$name.so ../../build/var_$name/search_tc.o $OBJS -lcudart -lcuda
echo "built graphtools_b200/variants/$name.so"
