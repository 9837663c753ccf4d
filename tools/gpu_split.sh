#!/bin/bash
# experiment: the resident solve as K column ranges on K streams (finish kernels of one range under the formal solution
# of another)
mkdir -p gpurun_out
for k in 2 3 4; do
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --split $k > gpurun_out/split_$k.json 2> gpurun_out/split_$k.err
  python -c "
import json
d = json.loads(open('gpurun_out/split_$k.json').read().strip().splitlines()[-1])
print('SPLIT $k', d['value'], d.get('split_streams_experiment'))
" || tail -5 gpurun_out/split_$k.err
done
