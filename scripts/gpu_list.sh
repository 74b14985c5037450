#!/bin/bash
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
last=rows[-9:]
tot=0
for r in last:
    print(f"{r[ki][:50]:50s} {float(r[vi])/1000:8.2f} us"); tot+=float(r[vi])/1000
print('sum',round(tot,1))
PY
