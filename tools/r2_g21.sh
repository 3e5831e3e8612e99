#!/bin/bash
# A/B: persisting L2 window over the Krylov basis
mkdir -p gpurun_out
for l2 in 1 0; do
  MHD_KRYLOV_L2=$l2 python bench.py --no-cpu-baseline --no-parity --steps 3 --warmup 3 > gpurun_out/r2_g21_l2$l2.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2_g21_l2$l2.json')); print('L2 window $l2: krylov', d['krylov']['ms_per_iteration'], d['krylov']['residual_reduction'], 'solve', d['solve']['iterations'], d['solve']['solve_ms'], d['solve']['true_relative_residual'], 'spmv', d['spmv']['ms'])"
done
