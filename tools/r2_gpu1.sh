#!/bin/bash
# round-2 GPU call 1: tests, bench (headline + emulated 32-image shard), stage timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_n1.err
for d in 1 2 3; do
  python bench.py --batch 32 --steps 600 --warmup 20 --no-configs --depth $d --e2e-steps 2 > gpurun_out/bench_b32_d$d.json 2>> gpurun_out/bench_b32.err
  python bench.py --batch 256 --steps 100 --warmup 10 --no-configs --depth $d --e2e-steps 2 > gpurun_out/bench_b256_d$d.json 2>> gpurun_out/bench_b256.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_b*_d*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f'%j['value'], 'ms/step %.4f'%j['ms_per_step'], 'serial %.4f'%j['step_ms_serial'], 'frac %.3f'%j['roofline']['frac'], 'launches', j['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
python tools/perf_probe.py cfg1 cfg2 > gpurun_out/probe_12.txt 2>&1; tail -30 gpurun_out/probe_12.txt
