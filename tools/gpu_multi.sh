#!/bin/bash
# multi-GPU checks: 2-GPU bit-identity test, then the bench at N GPUs (both arms), driver-style launch
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py -m gpu -q -k two_gpus 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/multi_n${N}.json 2> gpurun_out/multi_n${N}.err
tail -c 600 gpurun_out/multi_n${N}.err | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/multi_n${N}_ref.json 2>> gpurun_out/multi_n${N}.err
python - <<PY
import json
for f in ('gpurun_out/multi_n${N}.json', 'gpurun_out/multi_n${N}_ref.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4e' % d['value'], 'e2e %.4e' % d['e2e']['value'], 'n_gpus', d['n_gpus'], 'ms/step %.1f' % d['ms_per_step'])
    except Exception as ex:
        print(f, 'failed', ex)
PY
