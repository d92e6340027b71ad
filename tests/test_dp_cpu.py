"""Data-parallel host logic on CPU: world-size-2 gloo processes, the bucketed overlapped gradient
all-reduce (cti_b200.dp.GradAllReducer) must reproduce the single-process gradient of the full batch;
row sharding keeps question groups together (reference src/MC/train.py:75-79)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(12, 40), torch.nn.ReLU(), torch.nn.Linear(40, 24), torch.nn.Tanh(),
                               torch.nn.Linear(24, 3))


def full_batch():
    g = torch.Generator().manual_seed(11)
    return torch.randn(16, 12, generator=g), torch.randn(16, 3, generator=g)


def _worker(rank, world, port, bucket_bytes, out):
    import cti_b200  # noqa: F401
    from cti_b200.dp import GradAllReducer, clip_flat_grads_, shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = make_model()
    x, y = full_batch()
    sl = shard_rows(x.shape[0], rank, world, group=4)
    red = GradAllReducer(model.parameters(), bucket_bytes=bucket_bytes)
    for step in range(2):                                   # second step: hooks re-arm, grads re-point at the buckets
        red.zero_grad()
        loss = ((model(x[sl]) - y[sl]) ** 2).sum()
        loss.backward()
        red.finish()
    # hook-free path used after a CUDA-graph replay: gradients already sit in p.grad
    red.set_hooks_enabled(False)
    for p in model.parameters():
        p.grad = None
    ((model(x[sl]) - y[sl]) ** 2).sum().backward()
    red.reduce_now()
    again = [p.grad.clone() for p in model.parameters()]
    red.set_hooks_enabled(True)
    red.zero_grad()
    ((model(x[sl]) - y[sl]) ** 2).sum().backward()
    red.finish()
    for a_, p in zip(again, model.parameters()):
        assert torch.allclose(a_, p.grad, rtol=1e-6, atol=1e-7)
    grads = [p.grad.clone() for p in model.parameters()]
    norm = clip_flat_grads_(red.flat_grads(), 0.25, denom=float(x.shape[0]))
    clipped = [p.grad.clone() for p in model.parameters()]
    if rank == 0:
        torch.save({"grads": grads, "clipped": clipped, "norm": norm, "n_buckets": len(red.buckets)}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [1 << 30, 2048])
def test_allreduced_gradient_equals_single_process(tmp_path, bucket_bytes):
    out = str(tmp_path / "g.pt")
    port = 29500 + (os.getpid() % 2000) + (1 if bucket_bytes < 1 << 20 else 0)
    mp.spawn(_worker, args=(2, port, bucket_bytes, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    model = make_model()
    x, y = full_batch()
    ((model(x) - y) ** 2).sum().backward()
    ref = [p.grad for p in model.parameters()]
    for g, r in zip(res["grads"], ref):
        assert torch.allclose(g, r, rtol=1e-5, atol=1e-6)
    if bucket_bytes < 1 << 20:
        assert res["n_buckets"] > 1
    # div by the global row count, then clip to 0.25 (reference src/MC/trainer.py:213-214, src/utils.py:323-328)
    total = torch.sqrt(sum((r / 16).pow(2).sum() for r in ref))
    assert torch.allclose(res["norm"], total, rtol=1e-5)
    coef = min(1.0, 0.25 / (total.item() + 1e-6))
    for c, r in zip(res["clipped"], ref):
        assert torch.allclose(c, r / 16 * coef, rtol=1e-4, atol=1e-7)


def test_shard_rows_keeps_groups_and_covers_everything():
    from cti_b200.dp import shard_rows
    for n, world, group in [(256, 8, 4), (20, 3, 4), (12, 5, 4), (7, 2, 1)]:
        seen = []
        for r in range(world):
            s = shard_rows(n, r, world, group)
            assert s.start % group == 0 and s.stop % group == 0
            seen += list(range(s.start, s.stop))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_rows(10, 0, 2, 4)


def test_unused_parameters_are_reduced_as_zeros():
    from cti_b200.dp import GradAllReducer
    model = make_model()
    extra = torch.nn.Parameter(torch.ones(5))
    red = GradAllReducer(list(model.parameters()) + [extra])
    x, y = full_batch()
    ((model(x) - y) ** 2).sum().backward()
    red.finish()
    assert extra.grad is not None and torch.all(extra.grad == 0)
    assert all(p.grad is not None for p in model.parameters())


def _worker_groups(rank, world, port, out):
    """Explicit buckets whose gradients are produced IN PLACE and reduced early (launch_bucket), next to ordinary ones:
    what prepack.bind_grad_buffers(..., groups=...) drives from its per-group backward nodes."""
    import cti_b200  # noqa: F401
    from cti_b200.dp import GradAllReducer, shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = make_model()
    x, y = full_batch()
    sl = shard_rows(x.shape[0], rank, world, group=4)
    ps = list(model.parameters())
    early = [[ps[4], ps[5]], [ps[2]]]                       # last layer first: the order backward finishes them
    with pytest.raises(RuntimeError, match="CUDA"):          # the peer-memory transport has no CPU path; raised on every rank
        GradAllReducer(ps, param_groups=early, transport="peer")
    red = GradAllReducer(ps, param_groups=early)
    assert [len(b.params) for b in red.buckets[:2]] == [2, 1] and red.buckets[0].params[0] is ps[4]
    red.mark_in_place(early[0] + early[1])
    views = {p: v for b in red.buckets for p, v in zip(b.params, b.views)}
    res = {}
    for mode in ("hooks", "hook_free"):
        red.set_hooks_enabled(mode == "hooks")
        for step in range(2):
            for p in ps:
                p.grad = None
            loss = ((model(x[sl]) - y[sl]) ** 2).sum()
            gs = torch.autograd.grad(loss, ps)
            # the "producers": in-place groups write their gradient into the bucket view and launch; the rest arrive as
            # ordinary gradient tensors
            pos = {id(p): i for i, p in enumerate(ps)}
            in_place = {id(q) for grp in early for q in grp}
            for gi, grp in enumerate(early):
                for p in grp:
                    views[p].copy_(gs[pos[id(p)]])
                    p.grad = views[p]
                red.launch_bucket(gi)
            for p, g_ in zip(ps, gs):
                if id(p) not in in_place:
                    p.grad = g_.clone()
            if mode == "hooks":
                # no autograd hooks fired (gradients were assigned by hand): finish() reduces what is still pending
                red.finish()
            else:
                red.reduce_now()
        res[mode] = [p.grad.clone() for p in ps]
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_in_place_buckets_launched_early_give_the_full_batch_gradient(tmp_path):
    out = str(tmp_path / "g.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_groups, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    model = make_model()
    x, y = full_batch()
    ((model(x) - y) ** 2).sum().backward()
    for mode in ("hooks", "hook_free"):
        for g, p in zip(res[mode], model.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6), mode
