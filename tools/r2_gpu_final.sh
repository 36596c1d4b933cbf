#!/bin/bash
# end-of-round validation on one B200: full GPU test suite, smoke, the bench line at the driver's flags (with configs) and the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_n1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value %.0f' % j['value'], 'ms/step %.4f' % j['ms_per_step'], 'serial %.4f' % j['step_ms_serial'], 'frac %.3f' % j['roofline']['frac'],
      'launches/step', j['gpu_launches'] / j['steps'], 'e2e %.0f' % j['e2e']['value'], 'clocks', j.get('clocks'))
for k, v in j.get('configs', {}).items():
    if not isinstance(v, dict):
        continue
    print(k, 'ms %.4f' % v['ms'], 'frac %.3f' % v['roofline']['frac'], ('pipelined %.4f ms frac %.3f' % (v['pipelined']['ms'], v['pipelined']['roofline']['frac'])) if 'pipelined' in v else '')
r = json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('unit'), 'steps', r.get('steps'))
PY
