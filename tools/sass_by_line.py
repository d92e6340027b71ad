"""SASS instruction count per source-line range of one kernel: python tools/sass_by_line.py file.cubin kernel_substr [bucket]"""
import re, sys, collections, subprocess
cubin, sub = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur = None; infunc = False; per = collections.Counter(); files = collections.Counter()
for ln in txt.splitlines():
    if ln.startswith('.text.'):
        infunc = sub in ln; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if infunc and cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        per[cur] += 1
print('total', sum(per.values()))
agg = collections.Counter()
for (f, l), c in per.items():
    agg[(f, l // bucket * bucket)] += c
for (f, l), c in sorted(agg.items()):
    print(f"{f}:{l:5d}  {c}")
