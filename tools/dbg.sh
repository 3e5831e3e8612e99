#!/bin/bash
# Timing experiments on the Jacobian kernel (MHD_JAC_DEBUG bit mask: 1 no stores, 2 no main phase, 4 preparation only once, 8 no MMA)
for d in 0 1 2 4 5 6; do
  MHD_JAC_DEBUG=$d python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith(chr(123))][-1]); print('dbg=$d', d['ms_per_step'], d['roofline']['kernel_ms'], d['residual']['kernel_ms'])"
done
