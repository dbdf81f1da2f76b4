"""Summarise `ncu --page source --csv` output: stall mix of the kernel and its hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(r for r in rows if 'Source' in r and '# Samples' in r)
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows if len(r) == len(h) and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in data)
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = {n: sum(int(r[ix[n]] or 0) for r in data) for n in stalls}
print('samples', tot, ' stall mix:', ', '.join(f"{k[6:]} {100*v/tot:.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    st = {n: int(r[ix[n]] or 0) for n in stalls}
    print(f"{100*int(r[ix['# Samples']])/tot:5.1f}%  {max(st, key=st.get)[6:]:14s} {r[ix['Source']].strip()[:100]}")
