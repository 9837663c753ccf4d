#!/bin/bash
# quick GPU check during kernel work: the whole -m gpu suite (stop at the first failure) + the resident bench
mkdir -p gpurun_out
tag=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('value %.4e  e2e %.4e  ms/step %.2f  fs_ms %.3f  hbm frac %.4f  fp64 frac %.4f  share %.3f' % (
        d['value'], d['e2e']['value'], d['ms_per_step'], r['mean_launch_ms'], r['frac'], r['fp64']['frac'], r['share_of_step']))
except Exception as ex:
    print('bench failed', ex)
    print(open('gpurun_out/${tag}_bench.err').read()[-2000:])
PY
