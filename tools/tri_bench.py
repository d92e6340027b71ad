"""Micro-benchmark of the trilinear contraction kernels (cti_trilinear_logits_fwd / _bwd) at the bench shape, with a
parity check against an fp32 einsum on the same GPU and -- on CTI_PROF builds -- the per-role cycle accounting of the
forward kernel.

    python tools/tri_bench.py [--rows 1024] [--A 6] [--iters 20] [--prof]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cti_b200  # noqa: E402
from cti_b200 import _lib, kernels as K_  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1024)
ap.add_argument("--A", type=int, default=6)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--prof", action="store_true")
ap.add_argument("--no-bwd", action="store_true")
args = ap.parse_args()

B, K, Q, A, G, R, d = args.rows, 50, 12, args.A, 2, 32, 16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda n: (torch.relu(torch.randn(B * n, R * d, generator=g, device=dev)) * 0.5).to(torch.bfloat16)
vc, qc, ac = mk(K), mk(Q), mk(A)
teff = torch.randn(R, d, d, d, G, generator=g, device=dev)            # (r,i,j,l,g)
tpack = teff.permute(0, 3, 1, 4, 2).reshape(R, d, d * G * d).to(torch.bfloat16).contiguous()   # [r][l][(i,g,j)]
mask = (torch.rand(B * K, generator=g, device=dev) < 0.2).to(torch.uint8)


def reference():
    t = tpack.float().view(R, d, d, G, d)                             # (r,l,i,g,j)
    V, Qm, Am = vc.float().view(B, K, R, d), qc.float().view(B, Q, R, d), ac.float().view(B, A, R, d)
    out = torch.zeros(B, G, K, Q, A, device=dev)
    for lo in range(0, B, 128):
        s = slice(lo, lo + 128)
        n1 = torch.einsum("barl,rligj->barigj", Am[s], t)
        m = torch.einsum("bqrj,barigj->briqag", Qm[s], n1)
        out[s] = torch.einsum("bkri,briqag->bgkqa", V[s], m)
    return out


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


res = {"rows": B, "A": A}
out = K_.trilinear_fwd(vc, qc, ac, tpack, mask, B, K, Q, A, G, R)
ref = reference()
fin = torch.isfinite(out)
assert torch.equal(~fin, (mask.view(B, 1, K, 1, 1) != 0).expand_as(out)), "-inf set differs from the mask"
res["fwd_max_abs_err"] = (out[fin] - ref[fin]).abs().max().item()
res["fwd_scale"] = ref.abs().max().item()
flops = B * K_.trilinear_min_flops(K, Q, A, G, R)
med, best = timed(lambda: K_.trilinear_fwd(vc, qc, ac, tpack, mask, B, K, Q, A, G, R), args.iters)
res["fwd_us"] = med
med_s, _ = timed(lambda: K_.trilinear_fwd(vc, qc, ac, tpack, mask, B, K, Q, A, G, R, save_n1=True), args.iters)
res["fwd_save_n1_us"] = med_s
res["fwd_us_best"] = best
res["fwd_tflops_Tmin"] = flops / med / 1e6
if not args.no_bwd:
    dl = torch.randn(B, G, K, Q, A, generator=g, device=dev) * (mask.view(B, 1, K, 1, 1) == 0)
    _, n1 = K_.trilinear_fwd(vc, qc, ac, tpack, mask, B, K, Q, A, G, R, save_n1=True)
    med, best = timed(lambda: K_.trilinear_bwd(vc, qc, ac, tpack, dl, B, K, Q, A, G, R, n1=n1), args.iters)
    res["bwd_us"] = med
    res["bwd_us_best"] = best
    res["bwd_tflops_Tmin"] = 2 * flops / med / 1e6
if args.prof and not args.no_bwd:
    lib = _lib.load()
    lib.cti_debug_prof_read_bwd1.restype = ctypes.c_int
    lib.cti_debug_prof_read_bwd1.argtypes = [ctypes.c_void_p, ctypes.c_int]
    K_.trilinear_bwd(vc, qc, ac, tpack, dl, B, K, Q, A, G, R, n1=n1)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (148 * 128))()
    if lib.cti_debug_prof_read_bwd1(buf, 148 * 128) == 0:
        t = torch.tensor(list(buf), dtype=torch.float64).view(148, 16, 8)
        quads = (B + 147) // 148 * (R // 4)
        roles = ["tma: opempty n1empty dlempty", "b2: opfull dlfull b2empty issue", "f2: opfull n1full f2empty issue",
                 "b1: dlfull mfull b1empty issue", "b34: opfull n1full dtfull b3empty b4empty issue",
                 "c3: b2full dtempty work", "c2: f2full mempty work", "e1: opfull b1full work tmem",
                 "e4: - - - b4full e4work", "b4: opfull - dtfull - b4empty issue", "e3: opfull b3full work"]
        roles[4] = "b3: - n1full dtfull b3empty - issue"
        res["bwd1_prof_cycles_per_quad"] = {roles[i]: (t[:, i].mean(0) / quads).round().tolist() for i in range(11)}
        tb = (ctypes.c_ulonglong * (16 * 128))()
        if lib.cti_debug_prof_read_bwd1(tb, 16 * 128) == 0:
            tr = torch.tensor(list(tb), dtype=torch.float64).view(16, 128)
            names = ["tma", "b2", "c3", "c3done", "b3", "e3", "e4done", "f2", "c2done", "b1", "e1", "e1done"]
            t0 = tr[0, 16].item()
            print("# bwd1 timeline of block 0 (cycles relative to the TMA issue of quad 16)", file=sys.stderr)
            for c in range(16, 34):
                print(f"  c={c:3d} " + " ".join(f"{n}={int(tr[i, c].item() - t0):6d}" for i, n in enumerate(names)), file=sys.stderr)
if args.prof:
    lib = _lib.load()
    lib.cti_debug_prof_read.restype = ctypes.c_int
    lib.cti_debug_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    K_.trilinear_fwd(vc, qc, ac, tpack, mask, B, K, Q, A, G, R)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (148 * 64))()
    rc = lib.cti_debug_prof_read(buf, 148 * 64)
    if rc == 0:
        t = torch.tensor(list(buf), dtype=torch.float64).view(148, 8, 8)
        units = (B + 147) // 148 * R
        res["prof_cycles_per_unit_mean_over_blocks"] = (t.mean(0) / units).round().tolist()
        res["prof_units_per_block"] = units
        tb = (ctypes.c_ulonglong * (8 * 256))()
        if lib.cti_debug_prof_read(tb, 8 * 256) == 0:
            tr = torch.tensor(list(tb), dtype=torch.float64).view(8, 256)
            t0 = tr[0, 64].item()
            names = ["tma_issue", "f1_issue", "f2_issue", "iii_issue", "c1_start", "c2_start", "c1_done", "c2_done"]
            print("trace (block 0, cycles relative to the TMA issue of unit 64):")
            for u in range(64, 84):
                print(f"  u={u:3d} " + " ".join(f"{n}={int(tr[i, u].item() - t0):6d}" for i, n in enumerate(names)))
    else:
        res["prof"] = "library not built with CTI_PROF"
print(json.dumps(res))
