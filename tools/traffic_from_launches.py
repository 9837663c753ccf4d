#!/usr/bin/env python
"""DRAM traffic of the formal-solution kernels from an ncu launch list

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file launches.csv python bench.py --ncol 1024 --iters 2 --steps 1 --warmup 1 --no-e2e --no-cpu
    python tools/traffic_from_launches.py launches.csv [out.json]

Sums, over the three fs_gamma_kernel_m launches of the LAST formal solution in the list, DRAM bytes read + written
(per launch of the formal-solution stage, as bench.py's roofline.traffic wants it) and their ncu durations."""
import csv
import json
import sys


def main():
    rows = []
    with open(sys.argv[1], newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    order = []
    for r in rd:
        name = r['Kernel Name']
        key = (r['ID'], name)
        if key not in per:
            per[key] = {}
            order.append(key)
        val = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1.0)
        per[key][r['Metric Name']] = val * scale
    # bench.py also runs one solve in the exact arithmetic mode: keep the default (contracted, "(bool)1") launches only
    fs = [k for k in order if 'fs_gamma_kernel_m' in k[1] and ', 0>' not in k[1] and '(bool)0' not in k[1]]
    last = {}
    for k in fs:           # the last launch of each register class
        cls = k[1].split('fs_gamma_kernel_m')[1].split('(')[0]
        last[cls] = k
    out = {'source': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; '
                     'bench.py --ncol 1024 --iters 2 --steps 1 --warmup 1 --no-e2e --no-cpu; sum over the fs_gamma_kernel_m '
                     'launches of the last formal solution',
           'ncol': 1024, 'fixture': 'c2_falc_cah', 'per_kernel': {}}
    rdb = wrb = t = 0.0
    for cls, k in sorted(last.items()):
        m = per[k]
        out['per_kernel'][k[1]] = m
        rdb += m.get('dram__bytes_read.sum', 0.0)
        wrb += m.get('dram__bytes_write.sum', 0.0)
        t += m.get('gpu__time_duration.sum', 0.0)
    out.update(dram_bytes_read=rdb, dram_bytes_write=wrb, traffic_bytes_per_launch=rdb + wrb, ncu_time_ms=t)
    others = {}
    for k in order:        # time of every other kernel kind (last launch), for the share-of-step comparison
        if 'fs_gamma_kernel_m' not in k[1]:
            others[k[1].split('(')[0]] = per[k].get('gpu__time_duration.sum', 0.0)
    out['other_kernels_ms_last_launch'] = others
    print('read %.2f GB  write %.2f GB  total %.2f GB  fs time %.3f ms' % (rdb / 1e9, wrb / 1e9, (rdb + wrb) / 1e9, t))
    if len(sys.argv) > 2:
        json.dump(out, open(sys.argv[2], 'w'), indent=1)


if __name__ == '__main__':
    main()
