"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid) count,
mean duration and share of one step.  usage: summarize_launches.py launches.csv steps_in_capture"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    key = (r['Kernel Name'].split('(')[0][-58:], r['Grid Size'], r['Block Size'])
    agg.setdefault(key, []).append(float(r['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':58s} {'grid':>16s} {'block':>12s} {'n':>4s} {'mean us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[0]:58s} {k[1]:>16s} {k[2]:>12s} {len(v):4d} {sum(v)/len(v)/1e3:9.1f} {100*sum(v)/tot:5.1f}%")
