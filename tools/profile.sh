#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list + one full ncu capture of the dominant kernels. Run under gpurun.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobian_kernel -s 1 -c 1 -f -o gpurun_out/prof_jac \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_jac.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_warp_row -s 2 -c 1 -f -o gpurun_out/prof_spmv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:residual_kernel -s 1 -c 1 -f -o gpurun_out/prof_res \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_res.log 2>&1
ls -la gpurun_out
