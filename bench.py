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
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows per step of the CPU arms (0: --impl reference steps the "
                    "same --rows as the GPU arm, the in-line cpu_baseline of the GPU arm a bounded 128-row sample)")
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
# CPU arm: the reference's own modules (baseline/_ref, unmodified) on the host cores; the oracle port when the
# reference is not installed on this box
# --------------------------------------------------------------------------- #
def _reference_hot_path_modules(device="cpu"):
    """The hot path built from the REFERENCE's classes (src.attention.TriAttention, src.tc.TCNet, src.fc.FCNet), or None."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_env
    if ref_env.import_reference() is None:
        return None
    import torch
    from src.attention import TriAttention
    from src.fc import FCNet
    from src.tc import TCNet
    torch.manual_seed(1204)
    att = TriAttention(V_DIM, HID, HID, H_MM, 1, RANK, GLIMPSE, 1)
    pools = [TCNet(V_DIM, HID, HID, H_MM, 1, RANK, 1, k=2) for _ in range(GLIMPSE)]
    q_prj = [FCNet([HID, HID], '', .2) for _ in range(GLIMPSE)]
    a_prj = [FCNet([HID, HID], '', .2) for _ in range(GLIMPSE)]
    mods = torch.nn.ModuleList([att, *pools, *q_prj, *a_prj]).to(device).eval()
    return mods, att, pools, q_prj, a_prj


def _hot_path(att, pools, q_prj, a_prj, v, q, a):
    """The hot-path slice of TanModel.forward (reference src/MC/base_model.py:143-150)."""
    p_att, _ = att(v, q, a)
    qe, ae = q, a
    for gi in range(GLIMPSE):
        b_emb = pools[gi].forward_with_weights(v, qe, ae, p_att[:, :, :, :, gi])
        qe = q_prj[gi](b_emb.unsqueeze(1)) + qe
        ae = a_prj[gi](b_emb.unsqueeze(1)) + ae
    return qe.sum(1) + ae.sum(1)


def cpu_rows_per_s(rows, steps, warmup):
    """-> (rows/s, cores, s/step, kind): fwd+bwd of the hot path on all host cores, eval mode (dropout off)."""
    import torch
    from oracle import cti_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    v, q, a = O.synthetic_inputs(rows, K_REGIONS, Q_TOK, A_TOK, seed=1204)
    q.requires_grad_(True)
    a.requires_grad_(True)
    cot = torch.randn(rows, HID)
    ref = _reference_hot_path_modules()
    if ref is not None:
        mods, att, pools, q_prj, a_prj = ref
        leaves = list(mods.parameters()) + [q, a]
        run = lambda: _hot_path(att, pools, q_prj, a_prj, v, q, a)
        kind = "reference"
    else:
        params = {k: t.requires_grad_(True) for k, t in O.random_cti_params(glimpse=GLIMPSE, seed=1204).items()}
        leaves = list(params.values()) + [q, a]
        run = lambda: O.cti_hot_path(v, q, a, params, GLIMPSE)[0]
        kind = "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for t in leaves:
            t.grad = None
        (run() * cot).sum().backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return rows / dt, cores, dt, kind


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (unmodified modules from baseline/_ref when
    present, else the oracle port), all host threads, a bounded sample of the same workload.  Under torchrun only rank 0
    works; the line describes ONE host process whatever --gpus says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    args.cpu_rows = args.cpu_rows or args.rows
    val, cores, dt, kind = cpu_rows_per_s(args.cpu_rows, steps, warmup)
    sample = (f"{steps} timed fwd+bwd steps of {args.cpu_rows} rows after {warmup} warm-up: the GPU arm's own workload "
              f"(same modules, shapes and rows per step), bounded in the number of steps")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.cpu_rows, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "one host process on rank 0; the other ranks of a torchrun launch exit without work"}
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
    CAP_MODE = "thread_local" if world > 1 else "global"      # NCCL inside the captured step (graphs.GraphedStep)
    reducer = None
    if world > 1 or os.environ.get("CTI_FORCE_REDUCER"):   # (diagnostic: the reducer's local work on one GPU)
        # the deferred weight-norm backward writes dV / dg (97 % of the gradient bytes) straight into the all-reduce slab,
        # group by group in the order backward finishes them: glimpse 1's pooling + projections, glimpse 0's, the attention.
        # Each group owns one bucket whose all-reduce starts the moment the group is done and overlaps the rest of backward.
        grad_groups = [[pools[gi], q_prj[gi], a_prj[gi]] for gi in reversed(range(GLIMPSE))] + [[att]]
        # Transport: copy-engine pushes over NVLink peer memory (dp.PeerRegion) -- overlapped NCCL took SMs from the
        # persistent kernels (3.03 ms overlapped vs 2.97 ms serial at N = 2); CTI_TRANSPORT=nccl selects NCCL.
        TRANSPORT = os.environ.get("CTI_TRANSPORT", "peer")
        try:
            reducer = GradAllReducer(params, param_groups=cti_b200.weight_norm_param_groups(mods, grad_groups),
                                     transport=TRANSPORT)
        except RuntimeError as exc:          # no peer access on this box (raised on EVERY rank, dp.PeerRegion): NCCL
            if TRANSPORT != "peer" or world == 1:
                raise
            print(f"[bench] peer-memory transport unavailable ({exc}); using NCCL", file=sys.stderr, flush=True)
            TRANSPORT = "nccl"
            reducer = GradAllReducer(params, param_groups=cti_b200.weight_norm_param_groups(mods, grad_groups),
                                     transport=TRANSPORT)
        cti_b200.bind_grad_buffers(mods, reducer, groups=None if (os.environ.get("CTI_NO_OVERLAP") or
                                                    (TRANSPORT == "nccl" and not os.environ.get("CTI_NCCL_OVERLAP")))
                                   else grad_groups)        # (overlapped NCCL measured slower than NCCL after backward)
        if os.environ.get("CTI_NO_COLL"):                  # diagnostic: everything but the transfers themselves
            reducer.set_collectives_enabled(False)

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

    TRACE = bool(os.environ.get("CTI_PEER_TRACE")) and world > 1   # debug: timeline of the overlapped transfers

    def step(v, q, a):
        if TRACE and reducer.peer is not None:
            reducer.peer.stamp("main: step start")
        cti_b200.prepack(mods)                   # every weight-norm fold of the step in two launches (a training step
        for p in params:                         # would do this right after the optimizer update)
            p.grad = None
        q.requires_grad_(True)
        a.requires_grad_(True)
        p_att, _ = att(v, q, a)
        if step.fused:                           # opt-in: the glimpse loop + token sums as one call
            joint = cti_b200.glimpse_joint(pools, q_prj, a_prj, v, q, a, p_att)
        else:                                    # the reference model's own lines (src/MC/base_model.py:145-150)
            qe, ae = q, a
            for gi in range(GLIMPSE):
                b_emb = pools[gi].forward_with_weights(v, qe, ae, p_att[:, :, :, :, gi])
                qe = q_prj[gi](b_emb.unsqueeze(1)) + qe
                ae = a_prj[gi](b_emb.unsqueeze(1)) + ae
            joint = qe.sum(1) + ae.sum(1)
        (joint * cot).sum().backward()
        if TRACE and reducer.peer is not None:
            reducer.peer.stamp("main: backward done")
        if reducer is not None and step.reduce:
            # eager: the hooks launched the bucketed all-reduces during backward; captured: the hook-free reduce is part
            # of the graph (one fused copy of the few gradients that are not written in place + one all-reduce of the slab)
            if step.graphed:
                reducer.reduce_now()
            else:
                reducer.finish()
        return joint

    step.graphed = False
    step.fused = False
    step.reduce = True                           # False: rank-0-only passes (per-kernel profile) must not issue collectives

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

    sampler = ClockSampler(local) if rank == 0 else None      # nvidia-smi needs ~1 s to start: begin before the warm-up
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
            # multi-GPU: forward, backward AND the gradient all-reduce are replayed from the graph (hook-free reduce at the
            # end of the step; NCCL is capturable with capture_error_mode="thread_local")
            if reducer is not None:
                reducer.set_hooks_enabled(False)
                step.graphed = True
            if TRACE and reducer.peer is not None:
                reducer.peer.start_trace()
                orig_start = reducer.peer.start_trace
                graphed = None

                def one_trace():                   # labels are rebuilt by every (warm-up / capture) pass of the step
                    reducer.peer.trace_labels = []
                    resident_step()
                    reducer.peer.stamp("main: step end")
                graphed = cti_b200.GraphedStep(one_trace, [mods], [v_d], capture_error_mode=CAP_MODE)
            else:
                graphed = cti_b200.GraphedStep(resident_step, [mods], [v_d], capture_error_mode=CAP_MODE)
            run_resident = graphed.replay
            for _ in range(3):
                run_resident()
        except Exception as exc:
            use_graph = False
            step.graphed = False
            if reducer is not None:
                reducer.set_hooks_enabled(True)
            eager_ms["graph_capture_failed"] = repr(exc)[:200]
    ms, w0, w1 = timed(run_resident, args.steps)
    if TRACE and reducer.peer is not None and use_graph:
        torch.cuda.synchronize()
        for lab, t in reducer.peer.read_trace():
            print(f"[trace rank {rank}] {t / 1e3:9.1f} us  {lab}", file=sys.stderr, flush=True)
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
                g_sh = cti_b200.GraphedStep(shared_step, [mods], [vq_d], capture_error_mode=CAP_MODE)
                run_shared = g_sh.replay
                for _ in range(3):
                    run_shared()
            ms_sh, _, _ = timed(run_shared, args.steps)
            shared = {"value": world * B * args.steps / (ms_sh / 1e3), "unit": UNIT, "ms_per_step": ms_sh / args.steps,
                      "note": "v passed once per question (B/4 samples), no x4 clone: identical outputs, image-side "
                              "GEMMs once per image; needs the clone line of src/MC/train.py:75-76 removed"}
        except Exception as exc:
            shared = {"failed": repr(exc)[:200]}

    # ---- same step with the caller's glue fused (cti_b200.glimpse_joint replaces src/MC/base_model.py:145-150) -------
    fused_leg = None
    if not args.resident_only:
        fused_leg = {"note": "TriAttention + cti_b200.glimpse_joint (glimpse loop, residual adds and token sums as one autograd "
                             "node: same modules, parameters and values); opt-in, needs the six glue lines of the model's "
                             "forward replaced by one call"}
        try:
            step.fused = True
            for key, fn, vv in (("cloned_v", resident_step, v_d), ("shared_v", shared_step, vq_d)):
                if key == "shared_v" and CLONE == 1:
                    continue
                for _ in range(3):
                    fn()
                KS.STATS.launches = 0
                fn()
                n_launch = KS.STATS.launches
                run_f = fn
                if use_graph:
                    run_f = cti_b200.GraphedStep(fn, [mods], [vv], capture_error_mode=CAP_MODE).replay
                    for _ in range(3):
                        run_f()
                ms_f_, _, _ = timed(run_f, args.steps)
                fused_leg[key] = {"value": world * B * args.steps / (ms_f_ / 1e3), "unit": UNIT, "ms_per_step": ms_f_ / args.steps,
                                  "gpu_launches_per_step": n_launch}
        except Exception as exc:
            fused_leg["failed"] = repr(exc)[:300]
        finally:
            step.fused = False

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
                g_e2e = cti_b200.GraphedStep(e2e_step, [mods], [], capture_error_mode=CAP_MODE)
                run_e2e = g_e2e.replay
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
        # Two host formats: "bf16" = the loader wire format of this package (cti_b200.FeatureStoreBF16 / FeatureBatch:
        # padded bf16 features + zero-row mask; half the H2D bytes, no cast and no mask pass on the device) and "fp32" =
        # the fp32 batches the reference's own loader yields (src/utils.py:127-136).
        try:
            copy_stream = torch.cuda.Stream()
            v_h16 = v_h.to(torch.bfloat16).pin_memory()
            m_h = (v_h.abs().sum(2) == 0).to(torch.uint8).pin_memory()

            q_h16, a_h16 = q_h.to(torch.bfloat16).pin_memory(), a_h.to(torch.bfloat16).pin_memory()

            def pipelined(feat, share):
                # "bf16": every input in the wire format (features + mask, question / answer token embeddings in bf16);
                # "bf16v": bf16 features, fp32 tokens; "fp32": what the reference's loader yields
                host_src = {"bf16": (v_h16, m_h, q_h16, a_h16), "bf16v": (v_h16, m_h, q_h, a_h)}.get(feat, (v_h, q_h, a_h))
                bufs = [tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_src) for _ in range(2)]
                outs = [torch.empty(B, HID).pin_memory() for _ in range(2)]
                ev_copied = [torch.cuda.Event() for _ in range(2)]
                ev_done = [torch.cuda.Event() for _ in range(2)]

                def make_compute(i):
                    if feat != "fp32":
                        vb, mb, qb, ab = bufs[i]
                    else:
                        vb, qb, ab = bufs[i]
                        mb = None

                    def compute():
                        vv = vb
                        if mb is not None:
                            vv = cti_b200.prime_features(vb, mb.view(-1))
                        if not share:
                            vv = clone_rows(vv)
                            if mb is not None:                 # the clone of src/MC/train.py:75-76 carries the mask along
                                cti_b200.prime_features(vv, mb.unsqueeze(1).expand(Bq, CLONE, K_REGIONS).reshape(-1))
                        return step(vv, qb.detach(), ab.detach()).detach()
                    return compute
                cs = [make_compute(i) for i in range(2)]
                if use_graph:
                    cs = [cti_b200.GraphedStep(cs[i], [mods], [], capture_error_mode=CAP_MODE).replay for i in range(2)]
                state = {"i": 0, "primed": False}

                def h2d_into(i):
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(ev_done[i])             # the buffers' previous consumer has finished
                        for dst, src in zip(bufs[i], host_src):
                            dst.copy_(src, non_blocking=True)
                        ev_copied[i].record(copy_stream)

                def pipelined_step():
                    i = state["i"] & 1
                    if not state["primed"]:
                        h2d_into(i)
                        state["primed"] = True
                    main = torch.cuda.current_stream()
                    main.wait_event(ev_copied[i])
                    joint = cs[i]()
                    ev_done[i].record(main)
                    h2d_into(i ^ 1)                                    # prefetch the next step's inputs ...
                    with torch.cuda.stream(copy_stream):               # ... and read this step's result back behind them
                        copy_stream.wait_event(ev_done[i])
                        outs[i].copy_(joint, non_blocking=True)
                    state["i"] += 1
                for _ in range(4):
                    pipelined_step()
                torch.cuda.synchronize()
                state["primed"] = False                                # the timed region pays its own first copy
                ms_p, _, _ = timed(pipelined_step, args.steps)
                h2d_ = sum(t.numel() * t.element_size() for t in host_src)
                return {"value": world * B * args.steps / (ms_p / 1e3), "unit": UNIT, "ms_per_step": ms_p / args.steps,
                        "h2d_bytes_per_step": h2d_, "d2h_bytes_per_step": out_h.numel() * 4}
            e2e = dict(pipelined("bf16", False),
                       mode="pinned host batch in the loader wire format (bf16 features per question + zero-row mask, bf16 "
                            "question / answer token embeddings) -> H2D -> x4 clone on the device (src/MC/train.py:69-76) -> "
                            "modules -> D2H of the joint embedding; H2D of step i+1 overlapped with compute of step i "
                            "(double-buffered inputs)",
                       serial_fp32=e2e_serial)
            e2e["fp32_tokens"] = dict(pipelined("bf16v", False), note="bf16 features + mask, fp32 q / a (cast on the device)")
            e2e["fp32_features"] = dict(pipelined("fp32", False), note="same pipeline fed the fp32 batches the reference's "
                                        "loader yields (cast + mask pass on the device)")
            if CLONE > 1:
                e2e["shared_v"] = dict(pipelined("bf16", True), note="same bf16 host buffers, un-cloned v handed to the modules")
        except Exception as exc:
            e2e["pipelined_failed"] = repr(exc)[:300]

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
                run_fwd = cti_b200.GraphedStep(fwd_only, [], [v_d, q_d, a_d], capture_error_mode=CAP_MODE).replay   # inference: weight packs stay cached
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
                    run_ours = cti_b200.GraphedStep(ours_step, [], [], capture_error_mode=CAP_MODE).replay
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
                run_fb = cti_b200.GraphedStep(fb, [model], [vv], capture_error_mode=CAP_MODE).replay if use_graph else fb

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

    # ---- legs for the other BASELINE configs and modes (rank 0 at N = 1; none is part of `value`) -------------------
    def synth(rows, seed):
        gg = torch.Generator().manual_seed(seed)
        nq = rows // CLONE
        vq = torch.relu(torch.randn(nq, K_REGIONS, V_DIM, generator=gg))
        nb_ = torch.randint(10, K_REGIONS + 1, (nq,), generator=gg)
        vq = (vq * (torch.arange(K_REGIONS)[None, :] < nb_[:, None]).float()[:, :, None]).to(dev)
        vv = vq.unsqueeze(1).expand(nq, CLONE, K_REGIONS, V_DIM).contiguous().view(rows, K_REGIONS, V_DIM)
        return (vv, torch.tanh(torch.randn(rows, Q_TOK, HID, generator=gg)).to(dev),
                torch.tanh(torch.randn(rows, A_TOK, HID, generator=gg)).to(dev), torch.randn(rows, HID, generator=gg).to(dev))

    def time_hot_path(vv, qq, aa, cc, label, fixed_dropout=False, with_reducer=False):
        def fb():
            cti_b200.prepack(mods)
            for p in params:
                p.grad = None
            joint = _hot_path(att, pools, q_prj, a_prj, vv, qq.detach().requires_grad_(True), aa.detach().requires_grad_(True))
            (joint * cc).sum().backward()
            if with_reducer and reducer is not None:              # part of the step (and of its graph), as in the main arm
                reducer.reduce_now() if use_graph else reducer.finish()
        if with_reducer and reducer is not None:
            reducer.forget_sources()
        for _ in range(3):
            fb()
        run = fb
        if use_graph:
            run = cti_b200.GraphedStep(fb, [mods], [vv], allow_fixed_dropout=fixed_dropout, capture_error_mode=CAP_MODE).replay
        for _ in range(3):
            run()
        ms_, _, _ = timed(run, args.steps, all_ranks=with_reducer)
        rows_ = vv.shape[0]
        return {"rows_per_gpu": rows_, "ms_per_step": ms_ / args.steps,
                "value": (world if with_reducer else 1) * rows_ * args.steps / (ms_ / 1e3), "unit": UNIT, "note": label}

    extra = {}
    if not args.resident_only and not args.hot_only:
        # BASELINE config 5 also asks for scaling at EQUAL GLOBAL batch: 4096 rows split over the N ranks
        try:
            g_rows = 4096
            if g_rows % (world * 4) == 0:
                vv, qq, aa, cc = synth(g_rows // world, 77 + rank)
                extra["strong_scaling_4096"] = dict(time_hot_path(
                    vv, qq, aa, cc, "global 4096 rows split evenly over the ranks, fwd+bwd + gradient all-reduce; "
                    "at N = 1 this is the top of BASELINE config 2's 64-4096 row sweep", with_reducer=world > 1),
                    global_rows=g_rows, scaling="strong")
                del vv, qq, aa, cc
        except Exception as exc:
            extra["strong_scaling_4096"] = {"failed": repr(exc)[:300]}
    if world == 1 and not args.resident_only and not args.hot_only:
        # training mode: every dropout of the reference active, the R per-rank nets with independent masks
        # (reference src/tc.py:29-31); timed by replaying one captured step (fixed masks: timing only)
        try:
            mods.train()
            extra["train_dropout"] = time_hot_path(v_d, q_d, a_d, cot, "modules in train(): input dropout of every FCNet, "
                                                   "independent masks per rank net (reference semantics)", fixed_dropout=True)
        except Exception as exc:
            extra["train_dropout"] = {"failed": repr(exc)[:300]}
        finally:
            mods.eval()

        # the reference's own modules, eager fp32, on this GPU ("the stronger comparator")
        try:
            ref = _reference_hot_path_modules(dev)
            if ref is None:
                extra["reference_gpu_fp32"] = {"unavailable": "baseline/_ref not installed"}
            else:
                r_mods, r_att, r_pools, r_qp, r_ap = ref
                r_params = list(r_mods.parameters())

                def ref_step():
                    for p in r_params:
                        p.grad = None
                    joint = _hot_path(r_att, r_pools, r_qp, r_ap, v_d, q_d.detach().requires_grad_(True),
                                      a_d.detach().requires_grad_(True))
                    (joint * cot).sum().backward()
                for _ in range(2):
                    ref_step()
                n_ref = max(3, args.steps // 4)
                ms_r, _, _ = timed(ref_step, n_ref, all_ranks=False)
                extra["reference_gpu_fp32"] = {"rows_per_gpu": B, "ms_per_step": ms_r / n_ref, "value": B * n_ref / (ms_r / 1e3),
                                               "unit": UNIT, "note": "unmodified reference modules (baseline/_ref), eager "
                                               "PyTorch fp32 on the same B200, same rows / shapes, fwd+bwd, eval mode"}
                del r_mods, r_params, ref
        except Exception as exc:
            extra["reference_gpu_fp32"] = {"failed": repr(exc)[:300]}

        # BASELINE configs 3 and 4: whole free-form models, 256 rows, Distillation_Loss(T=5, alpha=0.005), fwd+bwd
        def ffoe_leg(kind):
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ref_env
            n_ans, A_ = (3129, 0) if kind == "ban" else (1484, 3)
            rows_ = 256
            torch.manual_seed(1204)
            how = "reference builder (baseline/_ref, unmodified) on the substituted modules"
            if ref_env.import_reference() is not None:
                import src.FFOE.base_model as ff
                a_, ds_ = ref_env.fake_args_dataset(n_ans)
                cti_b200.install()
                try:
                    model = (ff.build_ban if kind == "ban" else ff.build_cti)(a_, ds_)
                finally:
                    cti_b200.uninstall()
            else:
                from mc_model import BanStudent, CTIFreeForm
                how = "mirror of the reference model (tools/mc_model.py)"
                model = (BanStudent(3000, V_DIM, HID, GLIMPSE, n_ans) if kind == "ban"
                         else CTIFreeForm(3000, V_DIM, HID, H_MM, RANK, GLIMPSE, n_ans))
            model = model.to(dev).eval()
            tp = [p for p in model.parameters() if p.requires_grad]
            gg = torch.Generator().manual_seed(9)
            vv = torch.relu(torch.randn(rows_, K_REGIONS, V_DIM, generator=gg))
            nb_ = torch.randint(10, K_REGIONS + 1, (rows_,), generator=gg)
            vv = (vv * (torch.arange(K_REGIONS)[None, :] < nb_[:, None]).float()[:, :, None]).to(dev)
            bb = torch.rand(rows_, K_REGIONS, 6, generator=gg).to(dev)
            qt = torch.randint(0, 3000, (rows_, Q_TOK), generator=gg).to(dev)
            at = torch.randint(0, 3001, (rows_, 3), generator=gg).to(dev)
            teacher = torch.randn(rows_, n_ans, generator=gg).half().to(dev)          # fp16, as the reference stores them
            target = torch.zeros(rows_, n_ans)
            target[torch.arange(rows_), torch.randint(0, n_ans, (rows_,), generator=gg)] = 1.0
            target = target.to(dev)
            crit = cti_b200.Distillation_Loss(5, 0.005)

            def fb():
                cti_b200.prepack(model)
                for p in tp:
                    p.grad = None
                out = model(vv, bb, qt, None)[0] if kind == "ban" else model(vv, qt, at)
                crit(out, teacher, target).backward()
            for _ in range(3):
                fb()
            run = cti_b200.GraphedStep(fb, [model], [vv], capture_error_mode=CAP_MODE).replay if use_graph else fb
            for _ in range(3):
                run()
            ms_, _, _ = timed(run, args.steps, all_ranks=False)
            return {"rows_per_gpu": rows_, "ms_per_step": ms_ / args.steps, "value": rows_ * args.steps / (ms_ / 1e3),
                    "unit": UNIT, "classes": n_ans, "params": sum(p.numel() for p in tp), "model": how,
                    "note": "whole model fwd + Distillation_Loss(T=5, alpha=0.005, fp16 teacher logits) + bwd, eval-mode modules"}
        for kind, key in (("ban", "ban_config3"), ("cti", "ffoe_cti_config4")):
            try:
                extra[key] = ffoe_leg(kind)
            except Exception as exc:
                extra[key] = {"failed": repr(exc)[:300]}

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
        step.reduce = False                      # this pass runs on rank 0 only
        if reducer is not None:
            reducer.set_hooks_enabled(False)
            reducer.set_collectives_enabled(False)
        for _ in range(n_prof):
            resident_step()
        torch.cuda.synchronize()
        step.reduce = True
        if reducer is not None:
            reducer.set_collectives_enabled(True)
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
                ent["frac_of_bf16_peak"] = ent["tflops"] / tf_burst
            if nbytes:
                ent["gbs"] = nbytes / (avg * 1e-3) / 1e9
                ent["frac_of_hbm_peak"] = ent["gbs"] / bw_peak
            kernels.append(ent)
        # dominant kernel = the kernel function with the largest share of the step (all its launches together);
        # per-launch CUDA-event timings inside a 3 ms step run at the burst clock -> burst peak
        by_name = {}
        for (name, tag), (n, t, flops, nbytes) in agg.items():
            d = by_name.setdefault(name, [0, 0.0, 0.0, 0.0])
            d[0] += n
            d[1] += t
            d[2] += flops * n
            d[3] += nbytes * n
        traffic_tab = {}
        tr_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tr_path):
            traffic_tab = json.load(open(tr_path)).get("per_launch_dram_bytes", {})

        def roof_entry(name):
            n, t, fl, nb = by_name[name]
            ent = {"kernel": name, "launches_per_step": n // n_prof, "avg_launch_ms": t / n, "share_of_step": t / total,
                   "traffic": traffic_tab.get(name)}
            if fl:
                ach = fl / (t * 1e-3) / 1e12
                ent.update(bound="tensor", achieved=ach, peak=tf_burst, unit="TFLOP/s", frac=ach / tf_burst)
            else:
                ach = nb / (t * 1e-3) / 1e9
                ent.update(bound="hbm", achieved=ach, peak=bw_peak, unit="GB/s", frac=ach / bw_peak)
            return ent
        dom = max(by_name, key=lambda k: by_name[k][1])
        roofline = roof_entry(dom)
        roofline["peak_source"] = which + (": burst bf16 figure (per-launch CUDA-event timings inside a ~3 ms step at the "
                                           "burst clock); sustained %.0f" % tf_peak if roofline["bound"] == "tensor" else
                                           ": copy bandwidth")
        roofline["traffic_source"] = ("profiles/r02_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                      "averaged over the kernel's launches of one step)" if traffic_tab else None)
        roofline["algorithmic"] = ("2MNK summed over the launches" if dom == "cti_gemm_bf16" else
                                   "SURVEY 8d per-row figure x rows of the launch")
        # the step as a whole and the ops the verdict tracks
        step_flops = 1547.8e6 * B * (A_TOK == 6)
        tf_step = step_flops / (ms / args.steps * 1e-3) / 1e12 if step_flops else None
        roofline["step"] = {"flops_per_row_fwd_bwd": 1547.8e6, "tflops": tf_step, "frac_of_burst": tf_step / tf_burst if tf_step else None,
                            "frac_of_sustained": tf_step / tf_peak if tf_step else None, "ms_per_step": ms / args.steps,
                            "tensor_bound_floor_ms": step_flops / (tf_burst * 1e12) * 1e3 if step_flops else None}
        if "cti_gemm_bf16" in by_name:
            roofline["gemm_all"] = roof_entry("cti_gemm_bf16")
        tri = [k for k in ("cti_trilinear_logits_fwd", "cti_trilinear_logits_bwd") if k in by_name]
        if tri:
            t_tri = sum(by_name[k][1] for k in tri)
            f_tri = sum(by_name[k][2] for k in tri)
            roofline["contraction"] = {"kernels": tri, "us_per_step": t_tri / n_prof * 1e3,
                                       "flops": "3 x B x T_min (SURVEY 8d: cheapest contraction order a -> q -> v)",
                                       "tflops": f_tri / (t_tri * 1e-3) / 1e12, "frac": f_tri / (t_tri * 1e-3) / 1e12 / tf_burst,
                                       "share_of_step": t_tri / total, "target": 0.5}
            for k in tri:
                roofline["contraction"][k] = roof_entry(k)
        longest = max(agg.items(), key=lambda kv: kv[1][1] / kv[1][0])
        (lname, ltag), (ln, lt, lfl, lnb) = longest
        roofline["longest_launch"] = {"kernel": lname, "tag": ltag, "avg_ms": lt / ln,
                                      "tflops": lfl / (lt / ln * 1e-3) / 1e12 if lfl else None,
                                      "frac": lfl / (lt / ln * 1e-3) / 1e12 / tf_burst if lfl else None}
        roofline["hbm_kernels"] = [roof_entry(k) for k in sorted(by_name, key=lambda k: -by_name[k][1])
                                   if not by_name[k][2] and by_name[k][3]][:6]

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        args.cpu_rows = args.cpu_rows or 128
        val, cores, dt, kind = cpu_rows_per_s(args.cpu_rows, 3, 1)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"3 timed fwd+bwd steps of {args.cpu_rows} rows after 1 warm-up, "
                                  f"{'unmodified reference modules (baseline/_ref)' if kind == 'reference' else 'oracle port'} "
                                  f"on the host cores, {dt:.2f} s/step"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "host_issue_ms_per_step": host_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(B, world), launch="cuda_graph_replay" if use_graph else "eager",
                               **({"gradient_exchange": ("copy engines over NVLink peer memory, overlapped with backward per "
                                                         "gradient group (dp.PeerRegion)" if reducer.peer is not None else
                                                         "NCCL all-reduce after backward")} if reducer is not None and world > 1 else {})),
                "eager": eager_ms, "e2e": e2e, "gpu_launches": launches, "gpu_launches_per_step": launches_per_step, "clocks": clocks,
                "fwd_only": fwd, "shared_v": shared, "fused_glimpse_loop": fused_leg, "trainer_tail": tail, "gru": gru, "full_model": full, **extra,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that captured NCCL work keep the communicator busy at teardown (destroy_process_group() hangs behind
        # them, tools/nccl_graph_probe.py): finish the device work, agree that every rank is done, leave
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
