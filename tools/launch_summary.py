"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
        a = agg.setdefault(r[ki][:64], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:64s} n={n:5d} total={t:10.1f} us avg={t / n:9.2f} us share={100 * t / tot:5.1f}%")
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")


if __name__ == '__main__':
    main(sys.argv[1])
