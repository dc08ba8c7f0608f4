"""Per-kernel totals and the launch sequence of an `ncu --metrics gpu__time_duration.sum --csv` log.
Usage: python scripts/launch_seq.py <launches.csv> [substring filter for the sequence] [first] [count]"""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e6 if unit == 'ns' else v / 1e3 if unit == 'us' else v * 1e3 if unit == 's' else v
    seq.append((name, v, row['Grid Size'], row['Block Size']))
tot = collections.defaultdict(lambda: [0, 0.0])
for n, v, _, _ in seq:
    tot[n][0] += 1
    tot[n][1] += v
print('total ms:', round(sum(v[1] for v in tot.values()), 3), 'launches:', len(seq))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f'{v[1]:9.3f} ms {v[0]:5d} x {v[1] / v[0]:8.4f}  {k[:90]}')
if len(sys.argv) > 2:
    pat = sys.argv[2].split('|')
    sel = [s for s in seq if any(p in s[0] for p in pat)]
    a = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    b = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    for n, v, g, bl in sel[a:a + b]:
        print(f'{v:8.3f} ms  {n[:60]:60s} grid {g} block {bl}')
