#!/bin/bash
# resident bench in both arithmetic modes
mkdir -p gpurun_out
for mode in exact contracted; do
  MALI_ARITH=$mode python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/arith_${mode}.json 2> gpurun_out/arith_${mode}.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/arith_${mode}.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('ARITH %-10s value %.4e  ms/step %.2f  fs_ms %.3f  fp64 frac %.4f' % ('${mode}', d['value'], d['ms_per_step'], r['mean_launch_ms'], r['fp64']['frac']))
except Exception as ex:
    print('ARITH ${mode} failed', ex)
    print(open('gpurun_out/arith_${mode}.err').read()[-1500:])
PY
done
