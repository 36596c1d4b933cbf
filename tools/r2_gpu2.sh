#!/bin/bash
mkdir -p gpurun_out
python bench.py --batch 32 --steps 6 --warmup 3 --no-configs --e2e-steps 2 --depth 1 > gpurun_out/plain_b32.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_b32_d1.csv python bench.py --batch 32 --steps 6 --warmup 3 --no-configs --e2e-steps 2 --depth 1 > gpurun_out/ncu_b32.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_b32_d1.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:70]:
    print(r[4][:60], r[7], r[8], r[-1])
PY
