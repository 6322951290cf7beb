"""Aggregate an ncu `--page source --csv --print-source cuda,sass` export by source function.
usage: ncu -i rep --page source --csv --print-source cuda,sass > x.csv; python scripts/ncu_by_function.py x.csv src.cuh"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().split('\n')
hdr = rows[2]
ci = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
fn_at, cur = {}, '?'
for i, l in enumerate(src, 1):
    m = re.search(r'^(?:__device__|__global__|__host__)[^(]*?(\w+)\s*\(', l)
    if m: cur = m.group(1)
    fn_at[i] = cur
data = []
for r in rows[3:]:
    try: ln = int(r[0])
    except Exception: continue
    if r[2] != '-':  # sass rows have an address; the cuda rows carry the per-line sums
        continue
    data.append((ln, r[1], int(r[ci] or 0), int(r[si] or 0)))
tot = sum(d[2] for d in data) or 1; ts = sum(d[3] for d in data) or 1
print("total warp-instructions", tot, "samples", ts)
agg = {}
for ln, s, ie, sm in data:
    a = agg.setdefault(fn_at.get(ln, '?'), [0, 0]); a[0] += ie; a[1] += sm
for f, (ie, sm) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{f:28s} inst {ie/tot*100:5.1f}%  samples {sm/ts*100:5.1f}%")
print("-- top lines by samples")
for ln, s, ie, sm in sorted(data, key=lambda d: -d[3])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{ln:5d} inst {ie/tot*100:5.1f}% smp {sm/ts*100:5.1f}%  {s.strip()[:120]}")
