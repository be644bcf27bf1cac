"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel n / mean / min / max / share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
d, grid = collections.defaultdict(list), {}
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki])
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    d[name].append(v / 1000)
    grid[name] = r[gi]
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-45s n=%3d mean=%9.1f us  min=%9.1f max=%9.1f share=%5.1f%% grid=%s" % (k[:45], len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot, grid[k]))
