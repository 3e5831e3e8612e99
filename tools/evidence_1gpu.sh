#!/bin/bash
# round-2 evidence run on ONE GPU: GPU suite, bench line, reference arm, ncu launch list, full ncu capture of the Jacobian kernel, sanitizers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_final_gpu_suite.log 2>&1; tail -3 gpurun_out/r2_final_gpu_suite.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
python -c "
import json
d = json.load(open('gpurun_out/r2_final_bench.json')); r = json.load(open('gpurun_out/r2_final_reference.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'parity', d['parity']['ok'], d['parity']['jac_rel'], 'e2e', d['e2e']['value'], 'krylov', d['krylov']['ms_per_iteration'], 'cpu', d['cpu_baseline'], 'ref arm', r['value'], r['cpu_baseline'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/r2_final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hdiv_v7_jacobian -s 1 -c 1 -f -o gpurun_out/r2_final_prof_jac \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/r2_final_ncu.log 2>&1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_case.py hdiv > gpurun_out/r2_final_sanitizer_${tool}_hdiv_v7.log 2>&1
  tail -2 gpurun_out/r2_final_sanitizer_${tool}_hdiv_v7.log
done
