"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share.  usage: python tools/summarize_launches.py profiles/<file>.csv"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        tot[k] = tot.get(k, 0) + v
        cnt[k] += 1
    T = sum(tot.values())
    print('| kernel | launches | total ms | share |')
    print('|---|---:|---:|---:|')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v / T < 0.0005:
            continue
        print('| `%s` | %d | %.3f | %.1f %% |' % (k[:80], cnt[k], v, 100 * v / T))
    print('| **all** | %d | %.3f | 100 %% |' % (sum(cnt.values()), T))


if __name__ == '__main__':
    main(sys.argv[1])
