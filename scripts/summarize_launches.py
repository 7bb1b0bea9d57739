"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections, csv, re, sys
def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0]); per = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum': continue
        name = re.sub(r'^void ', '', re.sub(r'\(.*', '', row['Kernel Name']))
        name = name.replace('sdb::', '')
        v = float(row['Metric Value'].replace(',', ''))
        if row['Metric Unit'] == 'ns': v /= 1000.0
        agg[name][0] += 1; agg[name][1] += v; per.append((name, row['Grid Size'], v))
    return agg, per
agg, per = load(sys.argv[1])
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.0f} us over {sum(v[0] for v in agg.values())} launches (ncu per-launch times: cold-cache, serialised)")
fam = collections.defaultdict(float)
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.0f} us {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:7.1f} us  {k}")
    fam[re.sub(r'<.*', '', k)] += t
print("--- by family")
for k, t in sorted(fam.items(), key=lambda kv: -kv[1]):
    print(f"{t:9.0f} us {100*t/tot:5.1f}%  {k}")
if len(sys.argv) > 2:
    g = collections.defaultdict(lambda: [0, 0.0])
    for n_, gr, v in per:
        if sys.argv[2] in n_: g[(n_, gr)][0] += 1; g[(n_, gr)][1] += v
    for k, (n, t) in sorted(g.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f"{t:8.0f} us n={n:3d} avg {t/n:7.1f} {k}")
