"""Attribute ncu warp-stall samples to SOURCE LINES: python tools/ncu_lines.py rep.ncu-rep <launch idx> <obj.o> <kernel substr> [top]
(SASS order in the report == order in nvdisasm of the same build; lines come from -lineinfo.)"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, idx, obj, sub = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infunc = [], None, False
for ln in txt.splitlines():
    if ln.startswith('.text.'):
        infunc = sub in ln; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if infunc and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
isamp = h.index('Warp Stall Sampling (All Samples)')
stall = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
body = [r for r in rows[2:] if len(r) == len(h)]
print(rows[0][1][:90], "| sass", len(body), "vs disasm", len(lines))
agg = collections.defaultdict(lambda: collections.Counter())
for n, r in enumerate(body):
    key = lines[n] if n < len(lines) else ("?", 0)
    for i in stall:
        if r[i].isdigit() and int(r[i]):
            agg[key][h[i][6:]] += int(r[i])
tot = sum(sum(c.values()) for c in agg.values())
print("total samples", tot)
for key, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    print(f"{key[0]}:{key[1]:<5d} {sum(c.values()):6d}  {dict(c.most_common(3))}")
