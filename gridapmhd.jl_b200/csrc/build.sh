#!/bin/bash
# Builds libmhdb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function ${V7_NT:+-DMHD_V7_NT=$V7_NT}"
OBJS=""
PIDS=""
for f in abi symbolic assembly krylov solver comm postprocess h1h1 patch hdiv_v7; do
  stale=0
  [ -f $f.o ] || stale=1
  for dep in $f.cu common.h h1h1_cell.h patch_cell.h hdiv7_cell.h hdiv7_tables.h ../../include/mhdb200.h; do
    [ $dep -nt $f.o ] && stale=1
  done
  if [ $stale = 1 ]; then
    rm -f $f.o  # a failed compile must not leave a stale object behind for the link
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o $f.o &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $f.o"
done
for p in $PIDS; do wait $p || { echo "build.sh: a translation unit failed to compile" >&2; exit 1; }; done
$NVCC -shared -o libmhdb200.so $OBJS -lcudart -ldl
echo "built $(pwd)/libmhdb200.so"
