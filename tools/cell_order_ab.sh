#!/bin/bash
# A/B of the breadth-first cell traversal: bench line + DRAM bytes of the Jacobian kernel, with and without (MHD_V7_ORDER=0)
mkdir -p gpurun_out
for ord in 1 0; do
  export MHD_V7_ORDER=$ord
  TAG=r2_g17_ord$ord bash tools/quick.sh
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:hdiv_v7_jacobian -s 1 -c 1 --csv \
      --log-file gpurun_out/r2_g17_ord${ord}_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/r2_g17_ord${ord}_dram.csv | awk -F, '{print $(NF-2), $(NF-1), $NF}'
done
