#!/bin/bash
# 2 GPUs: where does the fused halo SpMV lose its time?  default / debug statistics / no-wait (timing only, results invalid)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
show() { python -c "
import json
d = json.load(open('gpurun_out/$1.json'))
print('$1: spmv ms', d['spmv']['ms'], 'kernel', d['spmv']['kernel_ms'], 'krylov', d['krylov']['ms_per_iteration'], 'parity', d['parity'] and d['parity']['ok'])
"; }
B="bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
timeout 600 bash -c "$(declare -f run); run 29514 $B" > gpurun_out/r2_mg2c_default.json 2> gpurun_out/r2_mg2c_default.err; show r2_mg2c_default
MHD_HALO_DEBUG=1 timeout 600 bash -c "$(declare -f run); run 29515 $B --no-parity" > gpurun_out/r2_mg2c_debug.json 2> gpurun_out/r2_mg2c_debug.err; show r2_mg2c_debug
grep "mhd halo" gpurun_out/r2_mg2c_debug.err
MHD_HALO_NCCL=1 timeout 600 bash -c "$(declare -f run); run 29516 $B --no-parity" > gpurun_out/r2_mg2c_nccl.json 2> gpurun_out/r2_mg2c_nccl.err; show r2_mg2c_nccl
MHD_CHECK_CASE=expansion6k MHD_CHECK_STRESS=2000 timeout 400 bash -c "$(declare -f run); run 29512 tests/multigpu_check.py" > gpurun_out/r2_mg2c_exp6k.log 2>&1; grep -h "MULTIGPU\|rank" gpurun_out/r2_mg2c_exp6k.log | tail -4
timeout 300 bash -c "$(declare -f run); run 29511 tests/multigpu_check.py" > gpurun_out/r2_mg2c_hunt.log 2>&1; grep -h "MULTIGPU\|rank" gpurun_out/r2_mg2c_hunt.log | tail -3
