#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key metrics per kernel + opcode histogram.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_pipe_lsu.sum',
        # the unit that bounds the formal-solution kernels: the L1 data pipe (one 128-byte wavefront per cycle and SM;
        # LDS / STS / SHFL and global accesses all pass through it)
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.sum']
for vals in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keys:
            print('%-70s %-12s %s' % (h, units[i], vals[i]))
    st = [(float(vals[i]), h) for i, h in enumerate(hdr)
          if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and vals[i] not in ('', 'n/a')]
    print('stalls per issue: ' + ', '.join('%s %.2f' % (h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v)
                                           for v, h in sorted(st, reverse=True)[:7]))
    print('---')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        secs.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
seen = set()
for sec in secs:
    if sec['name'] in seen:
        continue
    seen.add(sec['name'])
    h = sec['rows'][0]
    data = [r for r in sec['rows'][1:] if len(r) == len(h)]
    iA, iI, iS = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    tot = sum(int(r[iI]) for r in data)
    tots = max(1, sum(int(r[iS]) for r in data))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iA])
        o = m.group(2).split('.')[0] if m else '?'
        op[o] += int(r[iI])
        ops[o] += int(r[iS])
    print(sec['name'], 'instr', tot, 'SASS lines', len(data))
    print('  ' + '  '.join('%s %.1f%%/%.1f%%' % (o, 100 * c / tot, 100 * ops[o] / tots) for o, c in op.most_common(24)))
