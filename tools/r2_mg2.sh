#!/bin/bash
# 2-GPU checks: parity (Hunt, Expansion 6k with an RCB partition, H1-H1), stress of the fused halo SpMV, bench at N=2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); run 29511 tests/multigpu_check.py" > gpurun_out/r2_mg2_hunt.log 2>&1; grep -h "MULTIGPU\|rank" gpurun_out/r2_mg2_hunt.log | tail -4
MHD_CHECK_CASE=expansion6k MHD_CHECK_STRESS=2000 timeout 400 bash -c "$(declare -f run); run 29512 tests/multigpu_check.py" > gpurun_out/r2_mg2_exp6k.log 2>&1; grep -h "MULTIGPU\|rank" gpurun_out/r2_mg2_exp6k.log | tail -6
MHD_CHECK_FORMULATION=h1h1 timeout 300 bash -c "$(declare -f run); run 29513 tests/multigpu_check.py" > gpurun_out/r2_mg2_h1h1.log 2>&1; grep -h "MULTIGPU\|rank" gpurun_out/r2_mg2_h1h1.log | tail -4
timeout 600 bash -c "$(declare -f run); run 29514 bench.py --gpus 2 --steps 10 --warmup 3" > gpurun_out/r2_mg2_bench.json 2> gpurun_out/r2_mg2_bench.err
python -c "
import json
d = json.load(open('gpurun_out/r2_mg2_bench.json'))
print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'parity', d['parity'], 'spmv', d['spmv']['ms'], 'krylov', d['krylov'])
"
