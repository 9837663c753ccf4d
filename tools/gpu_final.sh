#!/bin/bash
# final lines of the round: both bench arms with their defaults (what the driver runs), smoke
mkdir -p gpurun_out
tag=${1:-v23}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1])
ro = d['roofline']
print('value %.4e e2e %.4e (%.1f ms) ms/step %.2f fs_ms %.3f frac %.4f fp64 %.4f cpu %.4e ref-arm %.4e ratio e2e %.1f value %.1f clocks %s' % (
    d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], ro['mean_launch_ms'], ro['frac'], ro['fp64']['frac'],
    d['cpu_baseline']['value'], r['value'], d['e2e']['value'] / r['value'], d['value'] / r['value'], d['clocks']))
for k in ('single_column', 'to_convergence', 'response_function', 'response_function_from_thermodynamic_state', 'config4_from_thermodynamic_state'):
    v = d.get(k) or {}
    print(k, {q: v[q] for q in v if q in ('seconds_to_converge', 'setup_seconds', 'solve_seconds', 'device_loop', 'matches_reference', 'iterations_max', 'iterations_min')})
PY
