"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
shapes = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(row['Metric Unit'], 1)
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    agg[name][0] += 1
    agg[name][1] += v
    if 'gemm_tc' in name:
        shapes[(name[-28:], round(v / 1e3, -1))] += 1
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{v[1]/1e6:10.2f} ms {100*v[1]/tot:5.1f}%  n={v[0]:5d}  avg={v[1]/v[0]/1e3:9.1f} us  {k[:90]}")
print("total ms", tot / 1e6)
for k, c in sorted(shapes.items(), key=lambda kv: -kv[0][1])[:16]:
    print(k, c)
