#!/bin/bash
# N GPUs (default 8): parity + stress of the fused halo SpMV on the Hunt and Expansion-6k partitions, then the bench line
N=${N:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export N
MHD_CHECK_STRESS=${STRESS:-10000} timeout 400 bash -c "$(declare -f run); run 29511 tests/multigpu_check.py" > gpurun_out/r2_mg${N}_hunt.log 2>&1; grep -h "MULTIGPU\|stress" gpurun_out/r2_mg${N}_hunt.log | tail -3
MHD_CHECK_CASE=expansion6k MHD_CHECK_STRESS=${STRESS:-10000} timeout 400 bash -c "$(declare -f run); run 29512 tests/multigpu_check.py" > gpurun_out/r2_mg${N}_exp6k.log 2>&1; grep -h "MULTIGPU\|stress" gpurun_out/r2_mg${N}_exp6k.log | tail -3
timeout 900 bash -c "$(declare -f run); run 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline" > gpurun_out/r2_mg${N}_bench.json 2> gpurun_out/r2_mg${N}_bench.err
python -c "
import json
d = json.load(open('gpurun_out/r2_mg${N}_bench.json'))
print('N=$N value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'parity', d['parity'] and d['parity']['ok'], 'spmv ms', d['spmv']['ms'], 'krylov', d['krylov']['ms_per_iteration'], 'e2e', d['e2e']['value'])
" || tail -5 gpurun_out/r2_mg${N}_bench.err
