#!/bin/bash
# benches every lightspinner_b200/_lib/libmali_b200_<tag>.so variant given on the command line (resident solve only)
mkdir -p gpurun_out
for tag in "$@"; do
  lib=libmali_b200_${tag}.so
  [ "$tag" = base ] && lib=libmali_b200.so
  MALI_LIB_NAME=$lib python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_formal or c1_free or determin" 2>&1 | tail -1
  MALI_LIB_NAME=$lib python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/var_${tag}.json 2> gpurun_out/var_${tag}.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/var_${tag}.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('VARIANT %-12s value %.4e  ms/step %.2f  fs_ms %.3f  fp64 frac %.4f' % ('${tag}', d['value'], d['ms_per_step'], r['mean_launch_ms'], r['fp64']['frac']))
except Exception as ex:
    print('VARIANT ${tag} failed', ex)
    print(open('gpurun_out/var_${tag}.err').read()[-1500:])
PY
done
