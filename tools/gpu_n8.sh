#!/bin/bash
# 8-GPU box: 2-GPU bit-identity test, then the column-sharded bench at N = 8 (both arms, driver-style launch)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_configs.py -m gpu -q -k two_gpus 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
$TR --nproc-per-node 8 --master-port 29542 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/scale_n8_ref.json 2>> gpurun_out/scale_n8.err
grep -c "NCCL INFO" gpurun_out/scale_n8.err
python - <<PY
import json
for f in ('scale_n8', 'scale_n8_ref'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'value %.4e' % d['value'], 'n_gpus', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'] if d.get('e2e') else '')
    except Exception as ex:
        print(f, 'failed', ex)
        print(open('gpurun_out/scale_n8.err').read()[-1500:])
PY
