"""Aggregate an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) into segments split at
WARPSYNC / BAR.SYNC instructions: warp-instructions, shared-memory wavefronts and stall samples per segment."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(x):
    try: return float(x)
    except ValueError: return 0.0
segs, cur = [], None
def new(): return {'inst': 0, 'wf': 0, 'wfx': 0, 'samples': 0, 'n': 0, 'ops': {}, 'st': {}}
cur = new()
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in data:
    src = r[ix['Source']]
    toks = src.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    e = f(r[ix['Instructions Executed']])
    cur['inst'] += e; cur['wf'] += f(r[ix['L1 Wavefronts Shared']]); cur['wfx'] += f(r[ix['L1 Wavefronts Shared Excessive']])
    cur['samples'] += f(r[ix['# Samples']]); cur['n'] += 1
    k = op.split('.')[0]
    cur['ops'][k] = cur['ops'].get(k, 0) + e
    for h in stalls: cur['st'][h] = cur['st'].get(h, 0) + f(r[ix[h]])
    if 'WARPSYNC' in src or 'BAR.SYNC' in src:
        cur['end'] = op; segs.append(cur); cur = new()
segs.append(cur)
tot = sum(s['samples'] for s in segs)
for s in segs:
    if s['inst'] == 0: continue
    top = sorted(s['ops'].items(), key=lambda x: -x[1])[:7]
    st = sorted(s['st'].items(), key=lambda x: -x[1])[:4]
    print(f"{s['n']:5d} sass {s['inst']/units:9.1f} winst {s['wf']/units:8.1f} wf (excess {s['wfx']/units:6.1f}) samples {100*s['samples']/tot:5.1f}% end={s.get('end','')[:12]:12s} "
          f"{[(k, round(v/units, 1)) for k, v in top]} {[(k[6:], round(100*v/max(1,s['samples']))) for k, v in st]}")
print('total', sum(s['inst'] for s in segs)/units, 'winst', sum(s['wf'] for s in segs)/units, 'wf')
