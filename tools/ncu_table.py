"""Markdown table from `ncu --csv --page raw` of a report captured with SpeedOfLight / LaunchStats /
MemoryWorkloadAnalysis sections: the last `n` launches (one step), by kernel and in launch order.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv [n]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[0]
body = [r for r in rows[2:] if len(r) == len(h)]
if n:
    body = body[-n:]
col = {k: i for i, k in enumerate(h)}


def g(r, k, default=0.0):
    if k not in col or r[col[k]] in ("", "n/a"):
        return default
    try:
        return float(r[col[k]].replace(",", ""))
    except ValueError:
        return default


def short(name):
    name = name.replace("cti::<unnamed>::", "").replace("void ", "")
    return name.split("(")[0][:60]


T = "gpu__time_duration.sum"
units = rows[1][col[T]] if T in col else "ns"
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3}.get(units, 1e-3)
tens = next((k for k in h if "pipe_tensor" in k and "pct_of_peak_sustained_active" in k and "hmma" in k), None) or \
    next((k for k in h if k.startswith("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")), None)
agg = collections.OrderedDict()
for r in body:
    k = short(r[col["Kernel Name"]])
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += g(r, T) * scale
    a[2] += g(r, "dram__bytes_read.sum")
    a[3] += g(r, "dram__bytes_write.sum")
tot = sum(a[1] for a in agg.values())
ru = rows[1][col["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in col else "byte"
bscale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(ru, 1e-6)
print(f"{len(body)} launches, {tot:.1f} us of kernel time\n")
print("| kernel | launches | time us | share | DRAM read MB | DRAM write MB |")
print("|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} | {a[2] * bscale:.1f} | {a[3] * bscale:.1f} |")
print("\n| # | kernel | grid | block | time us | tensor pipe % | DRAM read MB | DRAM write MB | DRAM % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(body):
    print(f"| {i} | `{short(r[col['Kernel Name']])}` | {r[col['Grid Size']] if 'Grid Size' in col else ''} | "
          f"{r[col['Block Size']] if 'Block Size' in col else ''} | {g(r, T) * scale:.1f} | {g(r, tens) if tens else 0:.1f} | "
          f"{g(r, 'dram__bytes_read.sum') * bscale:.1f} | {g(r, 'dram__bytes_write.sum') * bscale:.1f} | "
          f"{g(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {int(g(r, 'launch__registers_per_thread'))} |")
