#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'void |hb::<unnamed>::|hb::\(anonymous namespace\)::|unnamed>::', '', name)
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row['Metric Unit'], 1)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {T/1e6:.3f} ms of kernel time (cold-cache, serialised under ncu)")
print("share_pct,total_ms,launches,avg_us,kernel")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v/T*100:.2f},{v/1e6:.3f},{cnt[k]},{v/cnt[k]/1e3:.2f},{k}")
