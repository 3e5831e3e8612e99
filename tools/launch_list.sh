#!/bin/bash
# launch list (per-kernel durations) of one bench run incl. the Krylov cycle
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_g19_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > gpurun_out/r2_g19_b.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_g19_launches.csv')) if len(r) > 5]
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split('(')[0][:60]; v = float(r[vi].replace(',', '')); v = v / 1000.0 if r[ui] in ('ns', 'nsecond') else v
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f'{n:60s} n={c:4d} total={t:10.1f} us avg={t/c:8.1f} us')
P
