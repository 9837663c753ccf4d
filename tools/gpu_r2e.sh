#!/bin/bash
# round 2, third session: suite + bench of an upload-path change, and the launch list of the e2e path
mkdir -p gpurun_out
tag=${1:-v22}
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_quick.json 2> gpurun_out/${tag}_bench_quick.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_bench_quick.json').read().strip().splitlines()[-1])
    r = d['roofline']; e = d['e2e']
    print('value %.4e  e2e %.4e (%.1f ms, upload only %.1f)  host-phi e2e %.4e  ms/step %.2f  fs_ms %.3f clocks %s' % (
        d['value'], e['value'], e['ms_per_step'], e.get('upload_only_ms_per_step', -1), d['e2e_host_phi']['value'], d['ms_per_step'], r['mean_launch_ms'], d['clocks']))
except Exception as ex:
    print('bench failed', ex); print(open('gpurun_out/${tag}_bench_quick.err').read()[-1500:])
PY
CMDE="python bench.py --ncol 512 --iters 1 --steps 1 --warmup 1 --no-cpu --e2e-steps 1"
$CMDE > gpurun_out/${tag}_e2e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_e2e_launches.csv $CMDE > gpurun_out/${tag}_e2e_ncu.log 2>&1
python - <<PY
import csv, collections
tot = collections.Counter(); cnt = collections.Counter()
for r in csv.DictReader(l for l in open('gpurun_out/${tag}_e2e_launches.csv') if l.startswith('"')):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    tot[r['Kernel Name'][:60]] += v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(u, 1.0); cnt[r['Kernel Name'][:60]] += 1
for k, v in tot.most_common(8): print('E2E-LAUNCHES %-60s n=%4d total %.3f ms' % (k, cnt[k], v))
PY
