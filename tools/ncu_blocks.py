#!/usr/bin/env python
"""Basic-block table of one kernel from the SASS source page of an .ncu-rep holding several kernels:
python tools/ncu_blocks.py REP KERNEL_SUBSTRING [NTH] — blocks with >= 0.8 % of issue slots or stall samples, top stall reason each."""
import csv, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()))
secs, title = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': title = r[1]; continue
    if r and r[0] == 'Address': secs.append((title, r, [])); continue
    if secs and len(r) == len(secs[-1][1]): secs[-1][2].append(r)
secs = [s for s in secs if name in s[0] and s[2]]
title, hdr, data = secs[min(nth, len(secs) - 1)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_i = sum(int(r[ix['Instructions Executed']]) for r in data); tot_s = sum(int(r[ix['# Samples']]) for r in data) or 1
print(f"# {title[:90]}: {len(data)} SASS instructions, {tot_i} warp-instructions executed, {tot_s} samples")
cur, segs = None, []
for k, r in enumerate(data):
    i = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    st = {h: int(r[ix[h]] or 0) for h in stalls}
    if cur and abs(i - cur['i']) <= 0.02 * max(cur['i'], 1):
        cur['n'] += 1; cur['ti'] += i; cur['s'] += s; cur['end'] = k
        for h in stalls: cur['st'][h] += st[h]
    else:
        if cur: segs.append(cur)
        cur = dict(start=k, end=k, i=i, n=1, ti=i, s=s, st=st, first=r[ix['Source']].strip()[:44])
segs.append(cur)
for g in segs:
    if g['ti'] / tot_i > 0.008 or g['s'] / tot_s > 0.008:
        top = sorted(g['st'].items(), key=lambda kv: -kv[1])[:2]
        print(f"{g['start']:5d}-{g['end']:5d} n={g['n']:4d} exec={g['i'] / 1e3:8.1f}k issue%={100 * g['ti'] / tot_i:5.2f} samp%={100 * g['s'] / tot_s:5.2f} {top[0][0][6:]}:{top[0][1]} {top[1][0][6:]}:{top[1][1]} | {g['first']}")
