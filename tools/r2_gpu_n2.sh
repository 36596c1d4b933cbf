#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_peer_gpu.py tests/test_fake_ops.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 400 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
tail -c 1200 gpurun_out/bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --no-configs > gpurun_out/bench_n1_same_box.json 2>> gpurun_out/bench_n1_same_box.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'n', j['n_gpus'], 'value %.0f'%j['value'], 'ms/step %.4f'%j['ms_per_step'], 'frac %.3f'%j['roofline']['frac'], 'parity', j.get('gather_parity'), 'e2e %.0f'%j['e2e']['value'], 'zc %.0f full %.0f'%(j['e2e_zero_copy']['value'], j['e2e_full_copy']['value']), 'numa', j.get('numa_node'))
    except Exception as e: print(f, 'ERR', e)
PY
