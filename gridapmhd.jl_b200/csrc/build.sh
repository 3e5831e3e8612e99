#!/bin/bash
# Builds libmhdb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function"
OBJS=""
for f in abi symbolic assembly krylov solver comm postprocess h1h1 patch hdiv_v6; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ common.h -nt $f.o ] || [ h1h1_cell.h -nt $f.o ] || [ patch_cell.h -nt $f.o ] || [ hdiv_cell.h -nt $f.o ] || [ sumfac_uu.h -nt $f.o ] || [ ../../include/mhdb200.h -nt $f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o $f.o &
  fi
  OBJS="$OBJS $f.o"
done
wait
$NVCC -shared -o libmhdb200.so $OBJS -lcudart -ldl
echo "built $(pwd)/libmhdb200.so"
