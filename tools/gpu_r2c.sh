#!/bin/bash
# round 2, third session: the whole GPU evidence set of one kernel version in one call -- suite, both bench arms, the
# ncu launch list with DRAM bytes (roofline.traffic), and one --set full capture of the formal-solution kernels
mkdir -p gpurun_out
tag=${1:-v18}
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
tail -c 600 gpurun_out/${tag}_bench_reference.json
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('value %.4e  e2e %.4e (%.1f ms, upload only %.1f)  ms/step %.2f  fs_ms %.3f  hbm frac %.4f  fp64 frac %.4f  share %.3f exact %.4e clocks %s' % (
        d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('upload_only_ms_per_step', -1), d['ms_per_step'], r['mean_launch_ms'], r['frac'], r['fp64']['frac'], r['share_of_step'], d['exact_arith']['value'], d['clocks']))
except Exception as ex:
    print('bench failed', ex)
    print(open('gpurun_out/${tag}_bench.err').read()[-2000:])
PY
CMD="python bench.py --ncol 1024 --iters 2 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/${tag}_traffic_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches_traffic.csv $CMD > gpurun_out/${tag}_traffic_ncu.log 2>&1
python tools/traffic_from_launches.py gpurun_out/${tag}_launches_traffic.csv gpurun_out/${tag}_traffic.json | tail -3
bash tools/gpu_ncu_fs.sh ${tag} | tail -3
# launch list of the e2e path (upload kernels: re-layout, line profiles) on a 512-column chunk
CMDE="python bench.py --ncol 512 --iters 1 --steps 1 --warmup 1 --no-cpu --e2e-steps 1"
$CMDE > gpurun_out/${tag}_e2e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_e2e_launches.csv $CMDE > gpurun_out/${tag}_e2e_ncu.log 2>&1
python - <<PY
import csv, collections
tot = collections.Counter(); cnt = collections.Counter()
for r in csv.DictReader(l for l in open('gpurun_out/${tag}_e2e_launches.csv') if l.startswith('"')):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    tot[r['Kernel Name'][:60]] += v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(u, 1.0); cnt[r['Kernel Name'][:60]] += 1
for k, v in tot.most_common(14): print('E2E-LAUNCHES %-60s n=%4d total %.3f ms' % (k, cnt[k], v))
PY
python tools/gpu_latency.py > gpurun_out/${tag}_latency.json 2> gpurun_out/${tag}_latency.err; tail -c 800 gpurun_out/${tag}_latency.json
