#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_n1.err
for d in 2 3 4; do
  python bench.py --batch 32 --steps 800 --warmup 20 --no-configs --depth $d --e2e-steps 2 > gpurun_out/bench_b32_d$d.json 2>> gpurun_out/bench_b32.err
  python bench.py --batch 256 --steps 100 --warmup 10 --no-configs --depth $d --e2e-steps 2 > gpurun_out/bench_b256_d$d.json 2>> gpurun_out/bench_b256.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_b*_d*.json'))+['gpurun_out/bench_n1.json']:
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f'%j['value'], 'ms/step %.4f'%j['ms_per_step'], 'serial %.4f'%j['step_ms_serial'], 'frac %.3f'%j['roofline']['frac'], 'launches/step', j['gpu_launches']/j['steps'], 'e2e %.0f'%j['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
