#!/usr/bin/env python
"""Benchmark of the CTI hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows B] [--impl b200|reference]

A step = one forward + backward pass of the CTI multiple-choice hot path over one synthetic batch
of ``rows`` module rows per GPU (TriAttention -> per glimpse TCNet.forward_with_weights + q_prj /
a_prj residuals, reference src/MC/base_model.py:143-150; K=50 regions x 2048, 12 question tokens,
6 answer tokens, 2 glimpses, rank 32, random init).  Dropout is off (eval-mode modules with
gradients taken), which is the mode the parity tolerance is defined in.

Prints ONE JSON line (rank 0).  ``value`` = rows/s with inputs resident in HBM; ``e2e`` = rows/s
through the public module API with host (pinned) inputs, H2D and D2H inside the timed region.
``--impl reference`` times the CPU oracle (a port of the reference's own op order; the reference is
pure PyTorch and is not present on the GPU box) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_REGIONS, Q_TOK, A_TOK, GLIMPSE = 50, 12, 6, 2
V_DIM, HID, H_MM, RANK = 2048, 1024, 512, 32
METRIC = "cti_mc_hot_path_fwd_bwd_rows_per_s"
UNIT = "rows/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--rows", type=int, default=1024, help="module rows per GPU per step (weak scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-rows", type=int, default=128, help="rows per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="skip the e2e / forward-only arms (for ncu runs)")
    ap.add_argument("--hot-only", action="store_true", help="skip the legs outside the hot path (trainer tail, GRUs, whole model)")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly from Python (no CUDA graph)")
    return ap.parse_args()


def workload_config(rows, n):
    return {"workload": "CTI MC hot path fwd+bwd (TriAttention + 2x TCNet.forward_with_weights + q_prj/a_prj), "
                        "dropout off", "rows_per_gpu": rows, "global_rows": rows * n, "K": K_REGIONS, "Q": Q_TOK,
            "A": A_TOK, "glimpse": GLIMPSE, "rank": RANK, "h_mm": H_MM, "v_dim": V_DIM, "num_hid": HID,
            "rows_are": "rows/4 questions x 4 answer candidates, image features cloned per candidate",
            "l2": "inputs larger than L2 (v alone is %.0f MB per step)" % (rows * K_REGIONS * V_DIM * 4 / 1e6),
            "parallelism": f"dp{n}"}


# --------------------------------------------------------------------------- #
# CPU arm: the oracle (port of the reference's op order) on the host cores
# --------------------------------------------------------------------------- #
def cpu_rows_per_s(rows, steps, warmup):
    import torch
    from oracle import cti_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = {k: t.requires_grad_(True) for k, t in O.random_cti_params(glimpse=GLIMPSE, seed=1204).items()}
    v, q, a = O.synthetic_inputs(rows, K_REGIONS, Q_TOK, A_TOK, seed=1204)
    q.requires_grad_(True)
    a.requires_grad_(True)
    cot = torch.randn(rows, HID)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for t in list(params.values()) + [q, a]:
            t.grad = None
        joint, _, _ = O.cti_hot_path(v, q, a, params, GLIMPSE)
        (joint * cot).sum().backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return rows / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    val, cores, dt = cpu_rows_per_s(args.cpu_rows, steps, warmup)
    sample = f"{steps} timed steps of {args.cpu_rows} rows after {warmup} warm-up (same model and shapes)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.cpu_rows, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, mxv = float(f[0]), float(f[1])
            except ValueError:
                continue
            mx = mxv
            if t0 <= t <= t1 + 0.1:
                sm.append(clk)
                for n, val in zip(names, f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import cti_b200
    from cti_b200 import kernels as KS
    from cti_b200.dp import GradAllReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.rows
    torch.manual_seed(1204)                      # same replica on every rank
    att = cti_b200.TriAttention(V_DIM, HID, HID, H_MM, 1, RANK, GLIMPSE, 1)
    pools = [cti_b200.TCNet(V_DIM, HID, HID, H_MM, 1, RANK, 1, k=2) for _ in range(GLIMPSE)]
    q_prj = [cti_b200.FCNet([HID, HID], '', .2) for _ in range(GLIMPSE)]
    a_prj = [cti_b200.FCNet([HID, HID], '', .2) for _ in range(GLIMPSE)]
    mods = torch.nn.ModuleList([att, *pools, *q_prj, *a_prj]).to(dev).eval()
    params = [p for p in mods.parameters()]
    reducer = GradAllReducer(params) if world > 1 else None

    # multiple-choice batch: B rows = B / 4 questions x 4 answer candidates; the loader yields ONE feature tensor per
    # question and the trainer clones it per candidate on the device (reference src/MC/train.py:69-76)
    CLONE = 4 if B % 4 == 0 else 1
    Bq = B // CLONE
    g = torch.Generator().manual_seed(1204 + rank)
    v_h = torch.relu(torch.randn(Bq, K_REGIONS, V_DIM, generator=g))
    nb = torch.randint(10, K_REGIONS + 1, (Bq,), generator=g)
    v_h = (v_h * (torch.arange(K_REGIONS)[None, :] < nb[:, None]).float()[:, :, None]).pin_memory()

    def clone_rows(v):                           # v.unsqueeze(1).expand(...).contiguous().view(...) of train.py:75-76
        return v.unsqueeze(1).expand(Bq, CLONE, K_REGIONS, V_DIM).contiguous().view(B, K_REGIONS, V_DIM)
    q_h = torch.tanh(torch.randn(B, Q_TOK, HID, generator=g)).pin_memory()
    a_h = torch.tanh(torch.randn(B, A_TOK, HID, generator=g)).pin_memory()
    cot = torch.randn(B, HID, generator=g).to(dev)
    out_h = torch.empty(B, HID).pin_memory()

    def step(v, q, a):
        cti_b200.prepack(mods)                   # every weight-norm fold of the step in two launches (a training step
        for p in params:                         # would do this right after the optimizer update)
            p.grad = None
        q.requires_grad_(True)
        a.requires_grad_(True)
        p_att, _ = att(v, q, a)
        qe, ae = q, a
        for gi in range(GLIMPSE):
            b_emb = pools[gi].forward_with_weights(v, qe, ae, p_att[:, :, :, :, gi])
            qe = q_prj[gi](b_emb.unsqueeze(1)) + qe
            ae = a_prj[gi](b_emb.unsqueeze(1)) + ae
        joint = qe.sum(1) + ae.sum(1)
        (joint * cot).sum().backward()
        if reducer is not None and not step.graphed:
            reducer.finish()
        return joint

    step.graphed = False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, all_ranks=True):
        """CUDA-event time of `steps` calls.  all_ranks: barrier on both sides and the max over ranks (the contract's
        timing); rank-0-only legs pass False -- a collective there would wait for ranks that never call it."""
        sync = barrier if all_ranks else torch.cuda.synchronize
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        timed.host_ms = (time.perf_counter() - w0) * 1e3 / steps      # host time to ISSUE a step (launch-bound check)
        sync()
        w1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1 and all_ranks:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), w0, w1

    # ---- device-resident arm -------------------------------------------------
    vq_d, q_d, a_d = v_h.to(dev), q_h.to(dev), a_h.to(dev)
    v_d = clone_rows(vq_d)                       # the cloned (B, K, 2048) tensor the reference's model receives

    def resident_step():
        step(v_d, q_d.detach(), a_d.detach())

    def shared_step():                           # extension: the un-cloned features go straight in
        step(vq_d, q_d.detach(), a_d.detach())

    for _ in range(max(args.warmup, 3)):
        resident_step()
    # launches of one step (counted on an eager step: a graph replay issues the same kernels without Python)
    KS.STATS.launches = 0
    resident_step()
    launches_per_step = KS.STATS.launches
    use_graph = not args.no_graph
    eager_ms = None
    run_resident = resident_step
    if use_graph:
        ms_eager, _, _ = timed(resident_step, max(3, args.steps // 4))
        eager_ms = {"ms_per_step": ms_eager / max(3, args.steps // 4), "host_issue_ms_per_step": timed.host_ms}
        try:
            # multi-GPU: forward + backward are replayed from the graph, the bucketed NCCL all-reduce is issued
            # eagerly right after (capturing NCCL launched from autograd hooks hung)
            if reducer is not None:
                reducer.set_hooks_enabled(False)
                step.graphed = True
            graphed = cti_b200.GraphedStep(resident_step, [mods], [v_d])
            if reducer is None:
                run_resident = graphed.replay
            else:
                def run_resident():
                    graphed.replay()
                    reducer.reduce_now()
            for _ in range(3):
                run_resident()
        except Exception as exc:
            use_graph = False
            step.graphed = False
            if reducer is not None:
                reducer.set_hooks_enabled(True)
            eager_ms["graph_capture_failed"] = repr(exc)[:200]
    sampler = ClockSampler(local) if rank == 0 else None
    ms, w0, w1 = timed(run_resident, args.steps)
    host_ms = timed.host_ms
    launches = launches_per_step * args.steps
    clocks = sampler.stop(w0, w1) if sampler else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- same step with the un-cloned image features handed in (rows of one question share the image) ----------
    shared = None
    if CLONE > 1 and not args.resident_only:
        try:
            for _ in range(3):
                shared_step()
            run_shared = shared_step
            if use_graph:
                g_sh = cti_b200.GraphedStep(shared_step, [mods], [vq_d])
                if reducer is None:
                    run_shared = g_sh.replay
                else:
                    def run_shared():
                        g_sh.replay()
                        reducer.reduce_now()
                for _ in range(3):
                    run_shared()
            ms_sh, _, _ = timed(run_shared, args.steps)
            shared = {"value": world * B * args.steps / (ms_sh / 1e3), "unit": UNIT, "ms_per_step": ms_sh / args.steps,
                      "note": "v passed once per question (B/4 samples), no x4 clone: identical outputs, image-side "
                              "GEMMs once per image; needs the clone line of src/MC/train.py:75-76 removed"}
        except Exception as exc:
            shared = {"failed": repr(exc)[:200]}

    # ---- end-to-end arm: pinned host inputs -> H2D -> modules -> D2H ----------
    def e2e_step():
        v = clone_rows(v_h.to(dev, non_blocking=True))               # H2D per question, clone on the device
        q = q_h.to(dev, non_blocking=True)
        a = a_h.to(dev, non_blocking=True)
        joint = step(v, q, a)
        out_h.copy_(joint.detach(), non_blocking=True)

    e2e = fwd = None
    if not args.resident_only:
        for _ in range(3):
            e2e_step()
        run_e2e = e2e_step
        if use_graph:
            try:
                g_e2e = cti_b200.GraphedStep(e2e_step, [mods], [])
                if reducer is None:
                    run_e2e = g_e2e.replay
                else:
                    def run_e2e():
                        g_e2e.replay()
                        reducer.reduce_now()
                for _ in range(3):
                    run_e2e()
            except Exception:
                run_e2e = e2e_step
        ms_e2e, _, _ = timed(run_e2e, args.steps)
        h2d = (v_h.numel() + q_h.numel() + a_h.numel()) * 4
        e2e_serial = {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps}
        e2e = dict(e2e_serial, h2d_bytes_per_step=h2d, d2h_bytes_per_step=out_h.numel() * 4, mode="serial")

        # Same per-step work, but the host->device copy of step i+1 runs on a copy stream while step i computes
        # (two sets of static input buffers, one captured graph per set) -- what a prefetching loader gives a user.
        try:
            copy_stream = torch.cuda.Stream()
            bufs = [(torch.empty_like(vq_d), torch.empty_like(q_d), torch.empty_like(a_d)) for _ in range(2)]
            outs = [torch.empty(B, HID).pin_memory() for _ in range(2)]
            ev_copied = [torch.cuda.Event() for _ in range(2)]
            ev_done = [torch.cuda.Event() for _ in range(2)]

            def make_compute(i, share):
                vb, qb, ab = bufs[i]

                def compute():
                    joint = step(vb if share else clone_rows(vb), qb.detach(), ab.detach())
                    outs[i].copy_(joint.detach(), non_blocking=True)
                return compute

            def build_computes(share):
                cs = [make_compute(i, share) for i in range(2)]
                if use_graph:
                    graphs = [cti_b200.GraphedStep(cs[i], [mods], [bufs[i][0]]) for i in range(2)]
                    cs = [g.replay for g in graphs]
                return cs
            computes = build_computes(False)
            state = {"i": 0, "primed": False}

            def h2d_into(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_done[i])             # the buffers' previous consumer has finished
                    for dst, src in zip(bufs[i], (v_h, q_h, a_h)):
                        dst.copy_(src, non_blocking=True)
                    ev_copied[i].record(copy_stream)

            def pipelined_step():
                i = state["i"] & 1
                if not state["primed"]:
                    h2d_into(i)
                    state["primed"] = True
                main = torch.cuda.current_stream()
                main.wait_event(ev_copied[i])
                computes[i]()
                if reducer is not None and use_graph:
                    reducer.reduce_now()
                ev_done[i].record(main)
                h2d_into(i ^ 1)                                    # prefetch the next step's inputs
                state["i"] += 1
            for _ in range(4):
                pipelined_step()
            torch.cuda.synchronize()
            state["primed"] = False                                # the timed region pays its own first copy
            ms_p, _, _ = timed(pipelined_step, args.steps)
            e2e = {"value": world * B * args.steps / (ms_p / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": ms_p / args.steps,
                   "mode": "host v per question -> H2D -> x4 clone on the device (src/MC/train.py:69-76) -> modules; "
                           "H2D of step i+1 overlapped with compute of step i (double-buffered inputs)",
                   "serial": e2e_serial}
            if CLONE > 1:
                computes[:] = build_computes(True)
                state.update(i=0, primed=False)
                for _ in range(4):
                    pipelined_step()
                torch.cuda.synchronize()
                state["primed"] = False
                ms_ps, _, _ = timed(pipelined_step, args.steps)
                e2e["shared_v"] = {"value": world * B * args.steps / (ms_ps / 1e3), "unit": UNIT,
                                   "ms_per_step": ms_ps / args.steps,
                                   "note": "same host buffers, un-cloned v handed to the modules"}
        except Exception as exc:
            e2e["pipelined_failed"] = repr(exc)[:200]

    # ---- forward-only (config[1] of BASELINE.json) -----------------------------
    def fwd_only():
        with torch.no_grad():
            p_att, _ = att(v_d, q_d, a_d)
            qe, ae = q_d, a_d
            for gi in range(GLIMPSE):
                b_emb = pools[gi].forward_with_weights(v_d, qe, ae, p_att[:, :, :, :, gi])
                qe = q_prj[gi](b_emb.unsqueeze(1)) + qe
                ae = a_prj[gi](b_emb.unsqueeze(1)) + ae
            return qe.sum(1) + ae.sum(1)

    if not args.resident_only:
        for _ in range(3):
            fwd_only()
        run_fwd = fwd_only
        if use_graph:
            try:
                run_fwd = cti_b200.GraphedStep(fwd_only, [], [v_d, q_d, a_d]).replay   # inference: weight packs stay cached
                for _ in range(3):
                    run_fwd()
            except Exception:
                run_fwd = fwd_only
        ms_f, _, _ = timed(run_fwd, args.steps)
        fwd = {"value": world * B * args.steps / (ms_f / 1e3), "unit": UNIT, "ms_per_step": ms_f / args.steps}

    # ---- trainer tail on the hot path's parameters: fused clip + Adamax vs the reference's sequence (rank 0) -------
    tail = None
    if world == 1 and not args.resident_only and not args.hot_only:
        try:
            resident_step()                                     # leaves eager gradients in p.grad
            gparams = [p for p in params if p.grad is not None]
            n_el = sum(p.numel() for p in gparams)
            fused = cti_b200.FusedClipAdamax(gparams, lr=1e-3, clip_norm=0.25)
            ms_fused, _, _ = timed(lambda: fused.step(grad_denom=float(B)), args.steps, all_ranks=False)
            ref_opt = torch.optim.Adamax(gparams, lr=1e-3)

            def ref_tail():                                     # src/MC/trainer.py:208-219 + optimizer.step()
                flat = torch.cat([p.grad.reshape(-1) for p in gparams])
                flat.div_(float(B))
                norm = flat.norm()
                flat.mul_(torch.clamp(0.25 / (norm + 1e-6), max=1.0))
                off = 0
                for p in gparams:
                    p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                    off += p.numel()
                ref_opt.step()
            for _ in range(2):
                ref_tail()
            ms_ref, _, _ = timed(ref_tail, max(3, args.steps // 4), all_ranks=False)
            peaks_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
                os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            bw = peaks_.get("hbm_gbs", 6650.0)
            gbs = 32.0 * n_el / (ms_fused / args.steps * 1e-3) / 1e9
            tail = {"params": n_el, "fused_ms": ms_fused / args.steps, "bytes_per_param": 32, "gbs": gbs,
                    "frac_of_hbm_peak": gbs / bw, "torch_sequence_ms": ms_ref / max(3, args.steps // 4),
                    "note": "rescale + global-norm clip + Adamax over the hot path's parameters; not part of `value`"}
        except Exception as exc:
            tail = {"failed": repr(exc)[:200]}

    # ---- the step before the path: the two GRUs that produce q and a (rank 0; not part of `value`) ---------------
    gru = None
    if world == 1 and not args.resident_only and not args.hot_only:
        try:
            torch.manual_seed(7)
            ours = [cti_b200.QuestionEmbedding(600, HID, 1, False, .0).to(dev) for _ in range(2)]
            refs = [torch.nn.GRU(600, HID, 1, batch_first=True).to(dev) for _ in range(2)]
            xs = [torch.randn(B, T_, 600, device=dev, requires_grad=True) for T_ in (Q_TOK, A_TOK)]
            cots = [torch.randn(B, T_, HID, device=dev) for T_ in (Q_TOK, A_TOK)]

            def run_pair(mods_, call):
                for m_ in mods_:
                    for p_ in m_.parameters():
                        p_.grad = None
                loss = 0
                for m_, x_, c_ in zip(mods_, xs, cots):
                    x_.grad = None
                    loss = loss + (call(m_, x_) * c_).sum()
                loss.backward()
            ours_step = lambda: run_pair(ours, lambda m_, x_: m_.forward_all(x_))
            ref_step = lambda: run_pair(refs, lambda m_, x_: m_(x_)[0])
            for _ in range(6):                                  # cuDNN picks its algorithm during the first calls
                ours_step()
                ref_step()
            run_ours = ours_step
            if use_graph:
                try:
                    run_ours = cti_b200.GraphedStep(ours_step, [], []).replay
                except Exception:
                    run_ours = ours_step
            ms_o, _, _ = timed(run_ours, args.steps, all_ranks=False)
            ms_r, _, _ = timed(ref_step, args.steps, all_ranks=False)
            gru = {"rows": B, "tokens": [Q_TOK, A_TOK], "in_dim": 600, "hidden": HID,
                   "ours_fwd_bwd_ms": ms_o / args.steps, "torch_cudnn_fp32_fwd_bwd_ms": ms_r / args.steps,
                   "note": "question + answer GRU, forward_all + backward; not part of `value`"}
        except Exception as exc:
            gru = {"failed": repr(exc)[:200]}

    # ---- the whole MC model (embeddings -> GRUs -> hot path -> classifier -> BCE -> clip + Adamax), rank 0 ---------
    full = None
    if world == 1 and not args.resident_only and not args.hot_only:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from mc_model import MCModel
            torch.manual_seed(1204)
            model = MCModel(ntoken=3000, v_dim=V_DIM, num_hid=HID, h_mm=H_MM, rank=RANK, gamma=GLIMPSE).to(dev).eval()
            tparams = [p for p in model.parameters() if p.requires_grad]
            opt = cti_b200.FusedClipAdamax(tparams, lr=7e-4, clip_norm=0.25)
            gq = torch.Generator().manual_seed(5)
            q_tok = torch.randint(0, 3000, (Bq, Q_TOK), generator=gq).repeat_interleave(CLONE, 0).to(dev)
            a_tok = torch.randint(0, 3001, (B, A_TOK), generator=gq).to(dev)
            labels = torch.zeros(B, 2)
            labels[torch.arange(B), torch.randint(0, 2, (B,), generator=gq)] = 1.0
            labels = labels.to(dev)

            def make_fb(vv):
                def fb():
                    cti_b200.prepack(model)      # all weight-norm folds of the model, two launches
                    for p in tparams:
                        p.grad = None
                    logits, _ = model(vv, None, q_tok, a_tok)
                    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="sum") / B
                    loss.backward()
                    return loss
                return fb
            full = {"params": sum(p.numel() for p in tparams),
                    "note": "train step of the whole MC model: forward, BCE, backward (graph replay) + fused clip/Adamax; "
                            "not part of `value`"}
            for key, vv in (("cloned_v", v_d), ("shared_v", vq_d)):
                fb = make_fb(vv)
                for _ in range(3):
                    fb()
                    opt.step(grad_denom=1.0)
                run_fb = cti_b200.GraphedStep(fb, [model], [vv]).replay if use_graph else fb

                def train_step():
                    run_fb()
                    opt.step(grad_denom=1.0)
                for _ in range(2):
                    train_step()
                ms_t, _, _ = timed(train_step, args.steps, all_ranks=False)
                full[key] = {"ms_per_step": ms_t / args.steps, "rows_per_s": B * args.steps / (ms_t / 1e3),
                             "questions_per_s": Bq * args.steps / (ms_t / 1e3)}
        except Exception as exc:
            full = {"failed": repr(exc)[:300]}

    # ---- per-kernel CUDA-event timing of the same step (rank 0) -----------------
    roofline, kernels = None, None
    if rank == 0 and not args.no_profile:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        which = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        tf_burst = peaks.get("bf16_tflops", 1590.0)
        bw_peak = peaks.get("hbm_gbs", 6650.0)
        n_prof = 3
        KS.STATS.prof = []
        torch.cuda.synchronize()
        for _ in range(n_prof):
            resident_step()
        torch.cuda.synchronize()
        rec, KS.STATS.prof = KS.STATS.prof, None
        agg = {}
        for name, tag, flops, nbytes, e0, e1 in rec:
            key = (name, tag)
            d = agg.setdefault(key, [0, 0.0, flops, nbytes])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
        total = sum(d[1] for d in agg.values())
        kernels = []
        for (name, tag), (n, t, flops, nbytes) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            avg = t / n
            ent = {"kernel": name, "tag": tag, "launches_per_step": n // n_prof, "avg_ms": avg, "share": t / total}
            if flops:
                ent["tflops"] = flops / (avg * 1e-3) / 1e12
                ent["frac_of_bf16_peak"] = ent["tflops"] / tf_peak
            if nbytes:
                ent["gbs"] = nbytes / (avg * 1e-3) / 1e9
                ent["frac_of_hbm_peak"] = ent["gbs"] / bw_peak
            kernels.append(ent)
        gem = [(k, d) for k, d in agg.items() if k[0] == "cti_gemm_bf16"]
        if gem:
            gflops = sum(d[2] * d[0] for _, d in gem)
            gms = sum(d[1] for _, d in gem)
            (kname, ktag), d = max(gem, key=lambda kv: kv[1][1])
            ach = d[2] / (d[1] / d[0] * 1e-3) / 1e12
            roofline = {"kernel": f"gemm_bf16_kernel [{ktag}]", "bound": "tensor", "achieved": ach, "peak": tf_peak,
                        "unit": "TFLOP/s", "frac": ach / tf_peak,
                        # dram__bytes_read.sum + dram__bytes_write.sum of this launch in the committed ncu --set full
                        # capture (profiles/r01_ncu_gemm_dominant.md: 214.0 + 76.9 MB); algorithmic 210 + 4 + 105 MB
                        "traffic": 290.9e6 if ktag == "fwd M=51200 N=1024 K=2048" else None,
                        "peak_source": which + ", sustained bf16 (kernel timed inside a long step); burst %.0f" % tf_burst,
                        "share_of_step": d[1] / total,
                        "all_gemm_launches": {"tflops": gflops / (gms * 1e-3) / 1e12, "share_of_step": gms / total,
                                              "frac": gflops / (gms * 1e-3) / 1e12 / tf_peak}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        val, cores, dt = cpu_rows_per_s(args.cpu_rows, 3, 1)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"3 timed fwd+bwd steps of {args.cpu_rows} rows after 1 warm-up, oracle on host cores, "
                                  f"{dt:.2f} s/step"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "host_issue_ms_per_step": host_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(B, world), launch="cuda_graph_replay" if use_graph else "eager"),
                "eager": eager_ms, "e2e": e2e, "gpu_launches": launches, "gpu_launches_per_step": launches_per_step, "clocks": clocks,
                "fwd_only": fwd, "shared_v": shared, "trainer_tail": tail, "gru": gru, "full_model": full, "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
