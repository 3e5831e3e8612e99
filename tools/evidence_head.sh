#!/bin/bash
# refresh of the HEAD evidence on one GPU: suite, bench line, ncu summary of the assembly kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/head_gpu_suite.log 2>&1; tail -2 gpurun_out/head_gpu_suite.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/head_bench.json 2> gpurun_out/head_bench.err
python -c "
import json
d=json.load(open('gpurun_out/head_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['parity']['ok'], d['parity']['jac_rel'], d['krylov']['ms_per_iteration'], d['solve']['iterations'], d['solve']['solve_ms'])"
ncu --set full --clock-control none --import-source on -k regex:hdiv_v7_jacobian -s 1 -c 1 -f -o gpurun_out/head_prof_jac \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/head_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
