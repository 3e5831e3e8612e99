#!/bin/bash
# N=1: e2e after overlapping the clearing with the copy of x; large single-GPU points (--nc-global)
mkdir -p gpurun_out
TAG=r2_g15 bash tools/r2_quick.sh
for nc in "128 128" "256 256"; do
  t=${nc// /x}
  timeout 900 python bench.py --no-cpu-baseline --no-extra --steps 5 --warmup 3 --nc-global $nc > gpurun_out/r2_g15_nc$t.json 2> gpurun_out/r2_g15_nc$t.err
  python -c "
import json
d = json.load(open('gpurun_out/r2_g15_nc$t.json'))
print('nc $nc: value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'nnz', d['config']['nnz_local'], 'parity', d['parity'], 'spmv', d['spmv']['ms'], d['spmv']['roofline']['frac'], 'e2e', d['e2e']['value'])
" || tail -5 gpurun_out/r2_g15_nc$t.err
done
