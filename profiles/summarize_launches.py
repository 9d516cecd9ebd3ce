"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[ki].split('(')[0]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: per-kernel device time ({rows[1][ui]}), cold-cache + serialised under ncu: compare SHARES')
    print(f'{"kernel":60s} {"n":>5s} {"total":>12s} {"mean":>10s} {"share":>6s}')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{n:60s} {c:5d} {t:12.1f} {t / c:10.1f} {100 * t / tot:5.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
