"""profiles/rNN_traffic.json from an `ncu --page raw --csv` export of one step: measured DRAM bytes per C-ABI call.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python tools/traffic_json.py raw.csv <launches of one step> <bench line .json> > profiles/r02_traffic.json

`bench.py` reads `per_launch_dram_bytes[op]` into `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of
the kernels one C-ABI call launches, averaged over that op's calls in one step).  The number of calls per step comes from
the bench line's `kernels` list, the bytes from the ncu capture of the same step.
"""
import collections
import csv
import json
import sys

# kernel function (substring) -> C-ABI entry point that launches it (kernels.py `_call` names)
OPS = [("gemm_bf16_kernel", "cti_gemm_bf16"), ("pool_kernel<1>", "cti_tri_pool_bwd"), ("pool_kernel<0>", "cti_tri_pool_fwd"),
       ("trilinear_fwd_tc_kernel", "cti_trilinear_logits_fwd"), ("trilinear_bwd1_tc_kernel", "cti_trilinear_logits_bwd"),
       ("trilinear_bwd2_tc_kernel", "cti_trilinear_logits_bwd"), ("dlogits_to_dlm_kernel", "cti_trilinear_logits_bwd"),
       ("cast_rows_mask_kernel", "cti_cast_rows_mask"), ("act_bwd_bias_kernel", "cti_act_bwd_bias"),
       ("wn_multi_dot_kernel", "cti_wn_grad_multi"), ("wn_multi_grad_kernel", "cti_wn_grad_multi"),
       ("wn_multi_sumsq_kernel", "cti_wn_pack_multi"), ("wn_multi_scale_kernel", "cti_wn_pack_multi"),
       ("softmax_bwd_kernel", "cti_masked_softmax_bwd"), ("softmax_fwd", "cti_masked_softmax_fwd"),
       ("glimpse_", "cti_glimpse_glue"), ("rank_proj", "cti_rank_proj_dropout")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2])
    bench = json.loads([l for l in open(sys.argv[3]) if l.startswith("{")][-1])
    h = rows[0]
    col = {k: i for i, k in enumerate(h)}
    body = [r for r in rows[2:] if len(r) == len(h)][-n:]
    unit = rows[1][col["dram__bytes_read.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    per_step = collections.defaultdict(float)
    kernels = collections.defaultdict(int)
    for r in body:
        name = r[col["Kernel Name"]]
        op = next((o for s, o in OPS if s in name), None)
        if op is None:
            continue
        f = lambda k: float(r[col[k]].replace(",", "") or 0) * scale
        per_step[op] += f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
        kernels[op] += 1
    calls = collections.defaultdict(int)
    for k in bench.get("kernels") or []:
        calls[k["kernel"]] += k["launches_per_step"]
    out = {"source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per kernel launch, one step (%d launches of this library) "
                     "of `bench.py --steps 1 --warmup 3 --no-graph --resident-only`, summed per C-ABI entry point and divided "
                     "by its calls per step" % len(body),
           "per_step_dram_bytes": dict(per_step), "kernel_launches_per_step": dict(kernels), "calls_per_step": {},
           "per_launch_dram_bytes": {}}
    for op, b in per_step.items():
        c = calls.get(op) or kernels[op]
        out["calls_per_step"][op] = c
        out["per_launch_dram_bytes"][op] = b / c
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
