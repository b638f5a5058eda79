# prints the per-kernel times of one c2 frame from an ncu launch list (gpurun_out/rounds.csv)
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot = 0; agg = {}
for r in rows:
    name = r[4].split('(')[0].replace('vdbrt::', '').replace('void ', ''); v = float(r[-1].replace(',', '')) / 1000; tot += v
    agg.setdefault(name, []).append(v)
for k, v in agg.items(): print('  %-28s total %8.1f us : %s' % (k[:28], sum(v), ' '.join('%.0f' % x for x in v)))
print('  total %.1f us' % tot)
