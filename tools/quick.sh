#!/bin/bash
# quick A/B of the Jacobian kernel: bench line + phase clocks (+ optional pytest selection in $1)
mkdir -p gpurun_out
TAG=${TAG:-quick}
if [ -n "$1" ]; then timeout 900 python -m pytest $1 -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_tests.log; fi
timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
MHD_JAC_DEBUG=16 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 1 --no-parity --no-extra > /dev/null 2> gpurun_out/${TAG}_clocks.err
grep "phase clocks" gpurun_out/${TAG}_clocks.err | tail -1
python -c "
import json
d = json.load(open('gpurun_out/${TAG}_bench.json')); print('value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'res', d['residual'], 'parity', d['parity'] and d['parity']['ok'], d['parity'] and d['parity']['jac_rel'], 'e2e', d['e2e']['value'])
"
