"""Warp-stall samples of one kernel of an ncu report, grouped by source-line ranges of the .cu file (the warp roles of the
warp-specialised kernels): python tools/ncu_roles.py rep.ncu-rep file.cubin kernel_substr file.cu line:name [line:name ...]
Every SASS instruction is attributed to the last line of file.cu seen in program order (inlined helpers inherit it)."""
import csv, subprocess, sys, io, re, collections
rep, cubin, sub, cu = sys.argv[1:5]
marks = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[5:])
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infunc = [], None, False
for ln in txt.splitlines():
    if ln.startswith('.text.'):
        infunc = sub in ln; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if m.group(1).endswith(cu): cur = int(m.group(2))
        continue
    if infunc and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ia, isamp, iex = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
body = [r for r in rows[2:] if len(r) == len(h)]
assert len(body) == len(lines), (len(body), len(lines))
def role(l):
    name = '?'
    for a, n in marks:
        if l is not None and l >= a: name = n
    return name
tot, wait, ex = collections.Counter(), collections.Counter(), collections.Counter()
for r, l in zip(body, lines):
    k = role(l); s = int(r[isamp]); tot[k] += s; ex[k] += int(r[iex])
    if 'SYNCS.PHASECHK' in r[ia] or 'NANOSLEEP' in r[ia] or (' BRA ' in r[ia] and s > 30): wait[k] += s
allsamp = sum(tot.values())
print('role'.ljust(12), 'samples', 'share', 'in-wait', 'warp-instr')
for a, n in marks:
    print(n.ljust(12), str(tot[n]).rjust(7), f"{tot[n]/allsamp:6.1%}", f"{wait[n]/max(tot[n],1):7.1%}", str(ex[n]).rjust(10))
