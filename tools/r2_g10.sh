#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_g10_tests.log; tail -5 gpurun_out/r2_g10_tests.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2_g10_bench.json 2> gpurun_out/r2_g10_bench.err
python -c "
import json
d = json.load(open('gpurun_out/r2_g10_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'res', d['residual'], 'parity', d['parity'] and d['parity']['ok'], 'e2e', d['e2e']['value'])
print('krylov', d['krylov']); print('blas1', d['blas1']); print('solve', d['solve']); print('exp6k', d['expansion6k'])
"
MHD_JAC_DEBUG=16 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 1 --no-parity --no-extra > /dev/null 2> gpurun_out/r2_g10_clocks.err
grep "phase clocks" gpurun_out/r2_g10_clocks.err | tail -1
ncu --set full --clock-control none --import-source on -k regex:hdiv_v7_jacobian -s 1 -c 1 -f -o gpurun_out/r2_g10_prof_jac \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/r2_g10_ncu.log 2>&1
