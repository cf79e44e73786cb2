"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r'<.*', '', row['Kernel Name']).replace('void ', '')
    name = re.sub(r'\(.*', '', name)
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':40s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:40s} {v[0]:8d} {v[1]:12.1f} {v[1]/v[0]:10.2f} {100*v[1]/tot:6.1f}%")
print(f"{'total':40s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}")
