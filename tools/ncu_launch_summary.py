#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name."""
import collections
import csv
import sys


def main(path, top=40):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg, tot = collections.OrderedDict(), 0.0
    for row in rows:
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(row['Metric Unit'], 1e-6)
        a = agg.setdefault(row['Kernel Name'], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f'# {path}: {len(rows)} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)')
    print(f'# {"ms":>10} {"share":>6} {"count":>6}  kernel')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{t:12.3f} {100 * t / tot:5.1f}% {c:6d}  {k[:140]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
