#!/bin/bash
# round 2, second session: GPU suite + bench with the continuum-group kernels, then the e2e chunking sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('value %.4e  e2e %.4e (%.1f ms, upload only %.1f)  ms/step %.2f  fs_ms %.3f  hbm frac %.4f  fp64 frac %.4f  share %.3f exact %.4e clocks %s' % (
        d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('upload_only_ms_per_step', -1), d['ms_per_step'], r['mean_launch_ms'], r['frac'], r['fp64']['frac'], r['share_of_step'], d['exact_arith']['value'], d['clocks']))
except Exception as ex:
    print('bench failed', ex)
    print(open('gpurun_out/r2b_bench.err').read()[-2000:])
PY
for cfg in "512:128:3" "256:64:3" "256:0:3" "512:128:6" "384:128:3"; do
  IFS=: read chunk first es <<< "$cfg"
  python bench.py --steps 3 --warmup 2 --no-cpu --chunk $chunk --first-chunk $first --e2e-steps $es 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CHUNK $cfg value %.4e e2e %.4e (%.1f ms/step, upload only %.1f) host_phi %.4e' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('upload_only_ms_per_step', -1), d['e2e_host_phi']['value']))"
done
python tools/gpu_thermo_time.py 1024 > gpurun_out/r2b_thermo.json 2> gpurun_out/r2b_thermo.err; cat gpurun_out/r2b_thermo.json; tail -2 gpurun_out/r2b_thermo.err
