"""Aggregate an `ncu --page source --csv` SASS dump: instruction mix and stall samples per opcode and per
code region (regions split at BAR.SYNC).  usage: sass_mix.py file.csv"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0; byop = collections.Counter(); samp = collections.Counter()
regions = []; cur = collections.Counter(); cur_s = 0; cur_n = 0
def num(x):
    try: return int(float(x))
    except: return 0
for r in data:
    src = r[ix['Source']]; n = num(r[ix['Instructions Executed']]); s = num(r[ix['Warp Stall Sampling (All Samples)']])
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else src[:10]
    byop[op] += n; samp[op] += s; tot += n
    cur[op] += n; cur_s += s; cur_n += n
    if op == 'BAR':
        regions.append((cur_n, cur_s, cur)); cur = collections.Counter(); cur_s = 0; cur_n = 0
regions.append((cur_n, cur_s, cur))
stot = sum(samp.values())
print('total warp instr', tot, 'samples', stot)
for op, n in byop.most_common(24):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%  samples {100*samp[op]/max(1,stot):5.1f}%")
print('--- regions between barriers: instr%, samples%, top ops')
for n, s, c in regions:
    if n == 0: continue
    print(f"{100*n/tot:5.1f}% {100*s/max(1,stot):5.1f}%  " + ' '.join(f"{k}:{100*v/n:.0f}" for k, v in c.most_common(7)))
