#!/bin/bash
# SpMV tuning sweep on cfg2 (run under gpurun): unroll x CTAs/SM
for u in 1 2; do for c in 16 32 64 200; do
  echo -n "unroll=$u ctas=$c : "
  MHD_SPMV_UNROLL=$u MHD_SPMV_CTAS_PER_SM=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['spmv']['kernel_ms'], d['spmv']['roofline']['frac'])"
done; done
