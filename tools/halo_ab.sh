#!/bin/bash
# A/B of the SpMV ghost exchange on N GPUs (run under gpurun --gpus N): fused peer-memory kernel vs NCCL send/recv
N=${1:-2}
for mode in fused nowait nccl; do
  unset MHD_HALO_NCCL MHD_FUSED_NOWAIT
  if [ $mode = nccl ]; then export MHD_HALO_NCCL=1; fi
  if [ $mode = nowait ]; then export MHD_FUSED_NOWAIT=1; fi   # diagnostic only: skips the arrival wait (wrong results)
  port=$((29600 + RANDOM % 300))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/halo_ab_$mode.err | \
      python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$mode', 'N=$N', 'value', round(d['value'],3), 'spmv ms', round(d['spmv']['ms'],4), 'kernel_ms', round(d['spmv']['kernel_ms'],4), 'GB/s', round(d['spmv']['value'],1))" || tail -5 gpurun_out/halo_ab_$mode.err
done
