#!/usr/bin/env python
"""Per-instruction view of one kernel of an .ncu-rep (source page): execution-count classes and the hot loop."""
import collections
import csv
import subprocess
import sys


def main(rep, kernel, show=False):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix['Instructions Executed'] and r[ix['Instructions Executed']].isdigit()]
    data = data[:len(data) // 2] if len(data) > 1 and data[0][ix['Source']] == data[len(data) // 2][ix['Source']] else data
    E = lambda r: int(r[ix['Instructions Executed']])
    S = lambda r: int(r[ix['# Samples']])
    tot = sum(E(r) for r in data)
    ts = sum(S(r) for r in data)
    mx = max(E(r) for r in data)
    print('instructions', tot, 'samples', ts, 'max exec', mx)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in data:
        e = E(r)
        c = 'hot (>0.5 max)' if e > 0.5 * mx else 'warm (>0.05 max)' if e > 0.05 * mx else 'cold'
        agg[c][0] += e
        agg[c][1] += S(r)
        agg[c][2] += 1
    for k, v in agg.items():
        print('  {:18s} inst {:.3f} samples {:.3f} n {}'.format(k, v[0] / tot, v[1] / ts, v[2]))
    ops = collections.Counter()
    for r in data:
        if E(r) > 0.5 * mx:
            s = r[ix['Source']].split()
            op = s[1] if s[0].startswith('@') else s[0]
            ops[op.split('.')[0]] += 1
    print('  hot mix:', dict(ops.most_common()))
    if show:
        for i, r in enumerate(data):
            if E(r) > 0.05 * mx:
                print(i, r[ix['Source']][:72].ljust(72), E(r), S(r))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], len(sys.argv) > 3)
