"""Summarise one step out of an ncu launch list (gpu__time_duration.sum csv): per-kernel totals and the ordered list."""
import csv, sys, collections
path, nsteps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 5
rows = list(csv.reader(open(path)))
hdr = None; recs = []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: recs.append((d['Kernel Name'], d['Grid Size'], d['Block Size'], float(d['Metric Value'].replace(',', '')) / 1e3))
        except ValueError: pass
n = len(recs) // nsteps
last = recs[-n:]
if '--list' in sys.argv:
    for k, g, b, t in last: print(f"{t:8.1f} {g:16s} {b:14s} {k[:110]}")
agg = collections.defaultdict(list)
for k, g, b, t in last: agg[k.split('(')[0][-60:]].append(t)
tot = sum(t for *_, t in last)
print(f"launches/step {n}  total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):3d} sum={sum(v):8.1f} avg={sum(v)/len(v):7.1f} min={min(v):6.1f} max={max(v):6.1f} share={sum(v)/tot:.3f}")
