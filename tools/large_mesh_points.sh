#!/bin/bash
# N=1: parity at the large single-GPU points (--nc-global): sampled rows vs the oracle through a device gather + whole-matrix SpMV check
mkdir -p gpurun_out
export MHD_BENCH_BIG_NNZ=500000000
for nc in "128 128" "256 256"; do
  t=${nc// /x}
  timeout 1200 python bench.py --no-cpu-baseline --no-extra --steps 5 --warmup 3 --nc-global $nc > gpurun_out/r2_g16_nc$t.json 2> gpurun_out/r2_g16_nc$t.err
  python -c "
import json
d = json.load(open('gpurun_out/r2_g16_nc$t.json'))
print('nc $nc: value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'nnz', d['config']['nnz_local'], 'parity', d['parity'], 'spmv', d['spmv']['ms'], d['spmv']['roofline']['frac'], 'e2e', d['e2e']['value'])
" || tail -15 gpurun_out/r2_g16_nc$t.err
done
