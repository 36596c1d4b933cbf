#!/bin/bash
# the driver's own command line (--steps 20 --warmup 5) at N GPUs, then default flags
mkdir -p gpurun_out
N=${1:-2}
for a in "--steps 20 --warmup 5" "--steps 200 --warmup 20"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $a > gpurun_out/drv_n$N.json 2> gpurun_out/drv_n$N.err; echo "bench N=$N [$a] exit $?"
  python - <<PY
import json
j = json.loads(open('gpurun_out/drv_n$N.json').read().strip().splitlines()[-1])
print('n', j['n_gpus'], 'steps', j['steps'], 'value %.0f' % j['value'], 'ms/step %.4f' % j['ms_per_step'], 'parity', j.get('gather_parity'), 'e2e %.0f' % j['e2e']['value'], '|', j['config']['launch'][-60:])
PY
done
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n 1 steps %d value %.0f ms %.4f' % (j['steps'], j['value'], j['ms_per_step']))"
