#!/bin/bash
# N GPUs (default 8), final state: multigpu check (short stress), weak-scaling bench line, strong-scaling point (--nc-global 128 128)
N=${N:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export N
MHD_CHECK_STRESS=2000 timeout 400 bash -c "$(declare -f run); run 29511 tests/multigpu_check.py" > gpurun_out/r2_mg${N}b_hunt.log 2>&1; grep -h "MULTIGPU\|stress" gpurun_out/r2_mg${N}b_hunt.log | tail -2
timeout 900 bash -c "$(declare -f run); run 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline" > gpurun_out/r2_mg${N}b_bench.json 2> gpurun_out/r2_mg${N}b_bench.err
timeout 900 bash -c "$(declare -f run); run 29515 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra --nc-global 128 128" > gpurun_out/r2_mg${N}b_strong.json 2> gpurun_out/r2_mg${N}b_strong.err
for f in bench strong; do python -c "
import json
d = json.load(open('gpurun_out/r2_mg${N}b_$f.json'))
print('N=$N $f value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'parity', d['parity'] and d['parity']['ok'], d['parity'] and d['parity']['jac_rel'], 'spmv ms', d['spmv']['ms'], 'krylov', d['krylov']['ms_per_iteration'], 'e2e', d['e2e']['value'], d['scaling'], d['config'].get('ncells'))
" || tail -5 gpurun_out/r2_mg${N}b_$f.err; done
