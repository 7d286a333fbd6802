"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel/grid."""
import collections
import csv
import sys


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def summarise(rows, out=sys.stdout):
    agg = collections.OrderedDict()
    for r in rows:
        k = r['Kernel Name'].split('(')[0] + ' grid=' + r['Grid Size']
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += float(r['Metric Value']) / 1e3
    tot = sum(v[1] for v in agg.values())
    for k, v in agg.items():
        out.write('%-64s n=%4d total %10.1f us avg %8.1f us %5.1f%%\n' % (k[:64], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    out.write('total %.1f us over %d launches\n' % (tot, sum(v[0] for v in agg.values())))


if __name__ == '__main__':
    rows = load(sys.argv[1])
    marker = sys.argv[2] if len(sys.argv) > 2 else 'gray_kernel'
    idx = [i for i, r in enumerate(rows) if r['Kernel Name'].startswith(marker)]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else -2      # which occurrence of the marker starts the batch
    if len(idx) >= 2:
        a = idx[which]
        b = idx[which + 1] if which + 1 != 0 and which + 1 < len(idx) else (idx[-1] if which < 0 else len(rows))
        print('# one batch: launches %d..%d' % (a, b))
        summarise(rows[a:b])
    else:
        summarise(rows)
