#!/bin/bash
# 8-GPU box: the column-sharded bench at N = 8 (both arms, driver-style launch) and the wavelength-sharded stress
# column (BASELINE config 5 experiment) at 8, 4, 2 and 1 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
$TR --nproc-per-node 8 --master-port 29542 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/scale_n8_ref.json 2>> gpurun_out/scale_n8.err
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29550 + n)) bench.py --workload lambda_shard --gpus $n > gpurun_out/lambda_shard_n$n.json 2> gpurun_out/lambda_shard_n$n.err
done
python bench.py --workload lambda_shard --gpus 1 > gpurun_out/lambda_shard_n1.json 2> gpurun_out/lambda_shard_n1.err
grep -c "NCCL INFO" gpurun_out/scale_n8.err
python - <<PY
import json
for f in ('scale_n8', 'scale_n8_ref', 'lambda_shard_n8', 'lambda_shard_n4', 'lambda_shard_n2', 'lambda_shard_n1'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        extra = d.get('lambda_shard') or {}
        print(f, 'value %.4e' % d['value'], 'n_gpus', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'],
              'e2e %.4e' % d['e2e']['value'] if d.get('e2e') else '',
              {k: round(extra[k], 4) for k in ('ms_formal_solution', 'ms_gamma_allreduce', 'ms_stat_equil', 'allreduce_share_of_iteration') if k in extra})
    except Exception as ex:
        print(f, 'failed', ex)
        import os
        p = 'gpurun_out/%s.err' % f
        if os.path.isfile(p):
            print(open(p).read()[-1200:])
PY
