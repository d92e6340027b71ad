"""Summarise the source page of an ncu report for one kernel id: python tools/ncu_src.py rep.ncu-rep <launch index> [top]"""
import csv, subprocess, sys, collections, io
rep, idx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ia, isamp, iex = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
stall = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
body = [r for r in rows[2:] if len(r) == len(h)]
print(rows[0][1][:100])
tot = sum(int(r[isamp]) for r in body)
agg = collections.Counter()
for r in body:
    for i in stall:
        if r[i].isdigit(): agg[h[i]] += int(r[i])
print('samples', tot, 'instr', len(body), agg.most_common(9))
for n, r in enumerate(body): r.append(n)
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top]:
    st = sorted(((int(r[i]), h[i][6:]) for i in stall if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    print(str(r[-1]).rjust(6), r[isamp].rjust(5), r[iex].rjust(8), r[ia][:70].ljust(70), st)
