#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nms_gpu.py tests/test_yolo_gpu.py tests/test_fullsize_gpu.py tests/test_peer_gpu.py -x -q 2>&1 | tail -3
timeout 100 python tools/cfg1_probe.py 2>&1 | tail -9
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value %.0f' % j['value'], 'ms/step %.4f' % j['ms_per_step'], 'serial %.4f' % j['step_ms_serial'], 'frac %.3f' % j['roofline']['frac'], 'e2e %.0f' % j['e2e']['value'])
for k, v in j.get('configs', {}).items():
    if isinstance(v, dict):
        print(k, 'ms %.4f' % v['ms'], 'frac %.3f' % v['roofline']['frac'], ('pipelined %.4f ms frac %.3f' % (v['pipelined']['ms'], v['pipelined']['roofline']['frac'])) if 'pipelined' in v else '', v.get('eager_view_after_view_ms', ''), v.get('graph_equals_eager', ''))
PY
