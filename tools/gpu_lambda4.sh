#!/bin/bash
# wavelength-sharded stress column (BASELINE config 5 experiment) at 4, 2 and 1 GPUs with the round's final kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 4 2; do
  $TR --nproc-per-node $n --master-port $((29550 + n)) bench.py --workload lambda_shard --gpus $n > gpurun_out/lambda_shard_n$n.json 2> gpurun_out/lambda_shard_n$n.err
done
python bench.py --workload lambda_shard --gpus 1 > gpurun_out/lambda_shard_n1.json 2> gpurun_out/lambda_shard_n1.err
python - <<PY
import json
for f in ('lambda_shard_n4', 'lambda_shard_n2', 'lambda_shard_n1'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        extra = d.get('lambda_shard') or {}
        print(f, 'value %.4e' % d['value'], 'n_gpus', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'],
              {k: round(extra[k], 4) for k in ('ms_formal_solution', 'ms_gamma_allreduce', 'ms_gamma_exchange_itself', 'ms_stat_equil') if k in extra})
    except Exception as ex:
        print(f, 'failed', ex)
        print(open('gpurun_out/%s.err' % f).read()[-800:])
PY
