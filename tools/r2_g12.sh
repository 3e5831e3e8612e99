#!/bin/bash
# A/B: 256 vs 384 threads per CTA in the v7 kernel (rebuilt on the box)
mkdir -p gpurun_out
TAG=r2_g12_nt256 bash tools/r2_quick.sh "tests/test_hdiv_v7_gpu.py"
cd gridapmhd.jl_b200/csrc && touch hdiv_v7.cu && V7_NT=384 bash build.sh 2>&1 | grep -i "error\|built"; cd ../..
TAG=r2_g12_nt384 bash tools/r2_quick.sh "tests/test_hdiv_v7_gpu.py"
ncu --set full --clock-control none --import-source on -k regex:hdiv_v7_jacobian -s 1 -c 1 -f -o gpurun_out/r2_g12_nt384_prof \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/r2_g12_ncu.log 2>&1
