#!/usr/bin/env python
"""Opcode histogram of the innermost loops of a kernel, from `cuobjdump -sass` output.

    cuobjdump -sass lightspinner_b200/_lib/libmali_b200.so > all.sass
    python tools/sass_loops.py all.sass fs_gamma_kernel_mILi2E [min_len]

A loop = a backward branch (BRA to a lower address); nested outer loops are skipped when an inner one exists.
Used to see where the issue slots of the formal-solution kernel's depth loop go (fp64 vs integer vs LDS/SHFL)."""
import collections
import re
import sys


def main():
    path, fn = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    ins = []
    on = False
    pat = re.compile(r'^\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);')
    for line in open(path):
        if 'Function :' in line:
            on = fn in line
            continue
        if not on:
            continue
        m = pat.match(line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print('%s: %d instructions' % (fn, len(ins)))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r'\bBRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)', t)
        if m is None:
            m = re.search(r'BRA.*?(0x[0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                loops.append((addr_idx[tgt], i))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    inner = [l for l in inner if l[1] - l[0] + 1 >= min_len]
    for lo, hi in inner:
        ops = collections.Counter()
        for a, t in ins[lo:hi + 1]:
            t = re.sub(r'^@!?U?P\d+\s+', '', t)
            op = t.split()[0]
            ops[op.split('.')[0]] += 1
        n = hi - lo + 1
        f64 = sum(v for k, v in ops.items() if k in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX'))
        print('loop 0x%x..0x%x: %d instr, fp64 %d (%.0f%%)' % (ins[lo][0], ins[hi][0], n, f64, 100.0 * f64 / n))
        print('   ' + ' '.join('%s:%d' % kv for kv in ops.most_common()))


if __name__ == '__main__':
    main()
