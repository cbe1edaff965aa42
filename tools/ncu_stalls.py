"""Aggregate the warp-stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv` (read on the CPU box).

    ncu -i gpurun_out/X.ncu-rep --page source --csv > /tmp/src.csv ; python tools/ncu_stalls.py /tmp/src.csv <kernel index> [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        secs.append(cur)
    elif r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
for i, s in enumerate(secs):
    print(i, s['name'][:90], len(s['rows']))
s = secs[int(sys.argv[2])]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
h = s['hdr']
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = {n: 0 for n in stalls}
nsamp = ninst = 0
for r in s['rows']:
    for n in stalls:
        tot[n] += int(r[ix[n]] or 0)
    nsamp += int(r[ix['# Samples']] or 0)
    ninst += int(r[ix['Instructions Executed']] or 0)
print('samples', nsamp, 'warp instructions', ninst)
for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    print(f'  {n:28s} {v:8d} {100 * v / max(nsamp, 1):5.1f} %')
top = sorted(s['rows'], key=lambda r: -int(r[ix['# Samples']] or 0))[:ntop]
for r in top:
    st = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
    print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']].strip()[:56].ljust(56), st)
