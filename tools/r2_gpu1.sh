#!/bin/bash
# Round 2, first GPU call: full GPU suite incl. the v6 tests, v5 vs v6 bench, sanitizer on the H1-H1 / patch kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu1_smi.txt
MHD_RUN_V6_TESTS=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_gpu1_tests.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_gpu1_bench_v5.json 2> gpurun_out/r2_gpu1_bench_v5.err
MHD_JAC_V6=1 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_gpu1_bench_v6.json 2> gpurun_out/r2_gpu1_bench_v6.err
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_case.py h1h1 patch > gpurun_out/r2_gpu1_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_case.py h1h1 patch > gpurun_out/r2_gpu1_racecheck.log 2>&1
tail -3 gpurun_out/r2_gpu1_tests.log; tail -2 gpurun_out/r2_gpu1_memcheck.log; tail -2 gpurun_out/r2_gpu1_racecheck.log
