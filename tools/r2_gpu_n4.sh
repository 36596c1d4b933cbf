#!/bin/bash
# driver-style launch at N GPUs (default flags), then N=1 on the same box for the ratio
mkdir -p gpurun_out
N=${1:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err; echo "bench N=$N exit $?"
tail -c 600 gpurun_out/bench_n${N}_final.err
timeout 300 python bench.py --gpus 1 --no-configs > gpurun_out/bench_n1_same_box_final.json 2> /dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n*_final.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'n', j['n_gpus'], 'value %.0f'%j['value'], 'ms/step %.4f'%j['ms_per_step'], 'scaling', j['scaling'], 'frac %.3f'%j['roofline']['frac'], 'parity', j.get('gather_parity'), 'e2e %.0f'%j['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
