#!/bin/bash
mkdir -p gpurun_out
TAG=r2_g8 bash tools/r2_quick_ncu.sh "tests/test_hdiv_v7_gpu.py"
timeout 600 python tools/solve_cfg2.py 64 500 10:30:1:1.0:30:2 100:30:1:1.0:30:2 10:10:3:0.5:30:2 > gpurun_out/r2_g8_solve.log 2>&1
cat gpurun_out/r2_g8_solve.log | cut -c1-700
