#!/bin/bash
# quick A/B + one full ncu capture of the v7 Jacobian kernel
TAG=${TAG:-quick}
bash tools/quick.sh "$1"
ncu --set full --clock-control none --import-source on -k regex:hdiv_v7_jacobian -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_jac \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/${TAG}_ncu.log 2>&1
