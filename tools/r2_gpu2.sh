#!/bin/bash
# Round 2, second GPU call: v7 kernel -- new tests first, then the whole suite, bench A/B (v7 vs v5), phase clocks, ncu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hdiv_v7_gpu.py -x -q > gpurun_out/r2_gpu2_v7tests.log 2>&1; echo "v7 tests rc=$?" >> gpurun_out/r2_gpu2_v7tests.log
tail -5 gpurun_out/r2_gpu2_v7tests.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_hdiv_v7_gpu.py > gpurun_out/r2_gpu2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_gpu2_tests.log
tail -3 gpurun_out/r2_gpu2_tests.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_gpu2_bench_v7.json 2> gpurun_out/r2_gpu2_bench_v7.err
MHD_JAC_DEBUG=16 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 1 --no-parity > /dev/null 2> gpurun_out/r2_gpu2_clocks_v7.err
grep "phase clocks" gpurun_out/r2_gpu2_clocks_v7.err | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_v7.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hdiv_v7_jacobian -s 1 -c 1 -f -o gpurun_out/r2_prof_jac_v7 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_jac.log 2>&1
python -c "
import json
for f in ('gpurun_out/r2_gpu2_bench_v7.json',):
    try:
        d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['residual'], d['parity'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
"
