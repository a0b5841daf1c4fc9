"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:64]
    agg[name][0] += v
    agg[name][1] += 1
    tot += v
print('total %.1f us over %d launches' % (tot, sum(v[1] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-66s %10.1f us %5.1f%% %5d  avg %8.1f' % (k, v[0], 100 * v[0] / tot, v[1], v[0] / v[1]))
