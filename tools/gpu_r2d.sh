#!/bin/bash
# round 2, third session: suite + resident bench of a kernel change, then the e2e upload-schedule sweep (--sched)
mkdir -p gpurun_out
tag=${1:-v19}
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
show() {
python - "$1" "$2" <<PY
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d['roofline']
    e = d.get('e2e') or {}
    print('%s value %.4e  e2e %.4e (%.1f ms, upload only %.1f)  ms/step %.2f  fs_ms %.3f  fp64 frac %.4f  exact %.4e clocks %s' % (
        sys.argv[2], d['value'], e.get('value', 0), e.get('ms_per_step', 0), e.get('upload_only_ms_per_step', -1), d['ms_per_step'],
        r['mean_launch_ms'], r['fp64']['frac'], d['exact_arith']['value'], d['clocks']))
except Exception as ex:
    print(sys.argv[2], 'failed', ex)
PY
}
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench_quick.err
show gpurun_out/${tag}_bench_quick.json base
for sc in "64,128,256,576" "96,160,256,512" "128,256,640" "128,384,512" "32,64,128,256,544"; do
  python bench.py --steps 3 --warmup 3 --no-cpu --sched $sc --e2e-steps 3 > gpurun_out/${tag}_sched.json 2>/dev/null
  show gpurun_out/${tag}_sched.json "SCHED $sc"
done
