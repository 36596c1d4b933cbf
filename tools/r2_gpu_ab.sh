#!/bin/bash
# headline check after a decode-kernel change: tests, bench (no configs) twice, cfg4 pipelined
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_yolo_gpu.py tests/test_fullsize_gpu.py tests/test_golden_gpu.py tests/test_properties_gpu.py -x -q 2>&1 | tail -3
for round in 1 2; do
  python bench.py --no-configs --e2e-steps 2 > gpurun_out/ab_new_$round.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_new_*.json')):
    j = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.0f' % j['value'], 'ms %.4f' % j['ms_per_step'], 'serial %.4f' % j['step_ms_serial'])
PY
timeout 100 python tools/cfg4_pipe.py 1 3
