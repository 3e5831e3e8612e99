#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "nc64 or layout or spmv_dot or fgmres or hunt_solve or driver_end" > gpurun_out/r2_g9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_g9_tests.log; tail -5 gpurun_out/r2_g9_tests.log
timeout 600 python tools/solve_cfg2.py 64 1000 1 10:30:1:1.0:30:3 100:30:1:1.0:30:3 > gpurun_out/r2_g9_solve.log 2>&1
cat gpurun_out/r2_g9_solve.log | cut -c1-900
timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2_g9_bench.json 2> gpurun_out/r2_g9_bench.err
python -c "
import json
d = json.load(open('gpurun_out/r2_g9_bench.json')); print('value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'krylov', d['krylov'])
"
MHD_KRYLOV_GRAPH=0 timeout 400 python bench.py --no-cpu-baseline --steps 5 --warmup 3 --no-parity > gpurun_out/r2_g9_bench_nograph.json 2> gpurun_out/r2_g9_bench_nograph.err
python -c "
import json
d = json.load(open('gpurun_out/r2_g9_bench_nograph.json')); print('nograph krylov', d['krylov'])
"
