#!/bin/bash
# 2 GPUs: fused halo SpMV after the per-CTA acquire (bench N=2) + the pytest 2-GPU test
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); run 29514 bench.py --gpus 2 --steps 10 --warmup 3" > gpurun_out/r2_mg2b_bench.json 2> gpurun_out/r2_mg2b_bench.err
python -c "
import json
d = json.load(open('gpurun_out/r2_mg2b_bench.json'))
print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'parity', d['parity']['ok'], 'spmv ms', d['spmv']['ms'], d['spmv']['roofline']['frac'], 'krylov', d['krylov']['ms_per_iteration'])
"
MHD_HALO_NCCL=1 timeout 600 bash -c "$(declare -f run); run 29515 bench.py --gpus 2 --steps 5 --warmup 3 --no-parity" > gpurun_out/r2_mg2b_bench_nccl.json 2> gpurun_out/r2_mg2b_bench_nccl.err
python -c "
import json
d = json.load(open('gpurun_out/r2_mg2b_bench_nccl.json'))
print('N=2 NCCL halo: spmv ms', d['spmv']['ms'], 'krylov', d['krylov']['ms_per_iteration'])
"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k multigpu 2>&1 | tail -2
