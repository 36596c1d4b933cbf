timeout 600 python -m pytest tests/test_roi_strip_gpu.py -x -q 2>&1 | tail -4
python tools/roi_strip_probe.py
python tools/roi_strip_min.py > gpurun_out/rw_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rs_|roi_align|roi_level" -c 40 --csv --log-file gpurun_out/r2_roi_launches.csv python tools/roi_strip_min.py > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/r2_roi_launches.csv')):
    if len(r)>10 and r[0].isdigit(): print(r[4][:50], r[7], r[8], r[-1])
PY
