"""Per-kernel SASS opcode census of libcti_sm100.so: which functions carry tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM),
TMA (UTMALDG / UTMASTG / UTMAREDG / UBLKCP), mbarrier (SYNCS) and legacy HMMA instructions.

    python tools/sass_census.py [path/to/lib.so] > profiles/rNN_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "iccv19_vqa-cti_b200", "libcti_sm100.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "ELECT", "HMMA"]
per = collections.OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        per[cur]["_total"] += 1
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + "."):
                per[cur][o] += 1


def demangle(n):
    try:
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        d = n
    d = re.sub(r"cti::\(anonymous namespace\)::|cti::<unnamed>::|void ", "", d)
    depth = 0
    for i, ch in enumerate(d):          # cut the parameter list: the first "(" outside the template arguments
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            d = d[:i]
            break
    return d.replace("(bool)", "")[:64]


print("| kernel | SASS instr | " + " | ".join(ops) + " |")
print("|---|---|" + "---|" * len(ops))
tot = collections.Counter()
for f, c in sorted(per.items(), key=lambda kv: -kv[1]["_total"]):
    print(f"| `{demangle(f)}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in ops) + " |")
    tot.update(c)
print(f"| **all {len(per)} kernels** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in ops) + " |")
