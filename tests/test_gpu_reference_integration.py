"""The UNMODIFIED reference builders on the substituted modules, against the all-reference model, on the same GPU
(VERDICT r1 "next" item 8).

``baseline/_ref/src`` is a verbatim copy of /root/reference/src made by tools/install_reference.py (git-ignored; it
travels to the GPU box with the snapshot).  Per model the test
  1. builds the reference model with the reference's own classes (fp32, eager PyTorch on the GPU),
  2. calls ``cti_b200.install()`` and builds the SAME model with the UNCHANGED ``build_cti`` / ``build_ban`` -- the
     forward that runs is the reference's ``TanModel.forward`` / ``BanModel.forward`` / ``CTIModel.forward``,
  3. loads the reference's ``state_dict`` into it and compares class logits, attention maps, the chosen answer and one
     full trainer step (loss, backward, rescale + clip at 0.25, Adamax) between the two.
North_star tolerances: logits / attention <= 2e-2 max-abs, answer agreement >= 99.9 % (over >= 10 240 rows for the
multiple-choice model).
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
import ref_env  # noqa: E402

DEV = "cuda"
ABS_TOL = 2e-2

if ref_env.reference_root() is None:
    pytest.skip("the reference is not installed under baseline/_ref (tools/install_reference.py)", allow_module_level=True)


def _build_pair(kind, n_ans, hot_path_only=False):
    ref_env.import_reference()
    import src.MC.base_model as mc
    import src.FFOE.base_model as ff
    builder = {"mc_cti": mc.build_cti, "ffoe_ban": ff.build_ban, "ffoe_cti": ff.build_cti}[kind]
    args, ds = ref_env.fake_args_dataset(n_ans)
    torch.manual_seed(1204)
    ref = builder(args, ds)
    ref.classifier.main[2].inplace = False       # src/classifier.py:22: in-place dropout after ReLU trips torch-2 autograd
    cti_b200.install(hot_path_only=hot_path_only)
    try:
        new = builder(args, ds)
        assert type(new) is type(ref) and type(new).__module__.startswith("src.")   # the reference's own model class
    finally:
        cti_b200.uninstall()
    new.load_state_dict(ref.state_dict())
    return ref.to(DEV).eval(), new.to(DEV).eval(), args


def _inputs(rows, A, seed, clone=1):
    g = torch.Generator().manual_seed(seed)
    nq = rows // clone
    v = torch.relu(torch.randn(nq, 50, 2048, generator=g))
    nb = torch.randint(10, 51, (nq,), generator=g)
    v = v * (torch.arange(50)[None, :] < nb[:, None]).float()[:, :, None]
    v = v.repeat_interleave(clone, 0)
    q = torch.randint(0, 3000, (nq, 12), generator=g).repeat_interleave(clone, 0)
    a = torch.randint(0, 3001, (rows, A), generator=g) if A else None
    b = torch.rand(rows, 50, 6, generator=g)
    return v.to(DEV), b.to(DEV), q.to(DEV), None if a is None else a.to(DEV)


def _call(model, kind, v, b, q, a):
    if kind == "mc_cti":
        return model(v, b, q, a)
    if kind == "ffoe_ban":
        return model(v, b, q, None)
    return model(v, q, a), None


def _mc_answer(logits):
    """src/MC/trainer.py:292-299: the chosen candidate of each question = argmax over its 4 rows of softmax(logits)[:, 0]."""
    return torch.softmax(logits, 1)[:, 0].view(-1, 4).argmax(1)


@pytest.mark.parametrize("hot_path_only", [True, False])
def test_mc_model_unchanged_builder_logits_and_answers_over_10240_rows(hot_path_only):
    """hot_path_only=True: exactly the north_star's scope is substituted (FCNet / TCNet / TriAttention; the GRUs and
    the classifier stay the reference's fp32 modules) -> the chosen answer must agree on >= 99.9 % of the questions.
    hot_path_only=False: every drop-in, including the bf16 GRUs and classifier of SURVEY 8f rows 1-2.  The CPU
    emulation (oracle with bf16 rounding points, no kernel: tests/test_bf16_emulation_cpu.py) attributes ALL of the
    class-logit error of that configuration to those two components (hot path alone: 2e-5) and loses 0.2-0.4 % of the
    near-tied random-init answers to it; the kernels are held to 99.5 % there and the number is printed.
    The attention argmax (one of 3600 nearly equal weights at random init) flips on ~1.5 % of the maps under bf16
    operand rounding alone (same emulation), so it is held to 98 %, not 99.9 %."""
    ref, new, _ = _build_pair("mc_cti", 2, hot_path_only)
    agree = agree_p = total = total_p = 0
    worst_logit = worst_att = 0.0
    with torch.no_grad():
        for chunk in range(10):
            v, b, q, a = _inputs(1024, 6, 100 + chunk, clone=4)
            lr, ar = ref(v, b, q, a)
            ln, an = new(v, b, q, a)
            worst_logit = max(worst_logit, (ln - lr).abs().max().item())
            worst_att = max(worst_att, (an - ar).abs().max().item())
            agree += (_mc_answer(ln) == _mc_answer(lr)).sum().item()
            total += 256
            # argmax of the attention map per (row, glimpse)
            pr = ar.permute(0, 4, 1, 2, 3).reshape(1024 * 2, -1).argmax(1)
            pn = an.permute(0, 4, 1, 2, 3).reshape(1024 * 2, -1).argmax(1)
            agree_p += (pr == pn).sum().item()
            total_p += 2048
    print(f"\nMC model, 10240 rows: class-logit max-abs err {worst_logit:.3e}, attention {worst_att:.3e}, "
          f"answer agreement {agree}/{total}, attention-argmax agreement {agree_p}/{total_p}")
    assert worst_logit <= ABS_TOL and worst_att <= ABS_TOL
    assert agree / total >= (0.999 if hot_path_only else 0.995)
    assert agree_p / total_p >= 0.98


@pytest.mark.parametrize("kind,n_ans,A", [("ffoe_ban", 3129, 0), ("ffoe_cti", 1484, 3)])
def test_ffoe_models_unchanged_builders(kind, n_ans, A):
    ref, new, _ = _build_pair(kind, n_ans)
    agree = total = 0
    worst = worst_att = scale = 0.0
    with torch.no_grad():
        for chunk in range(4):
            v, b, q, a = _inputs(256, A, 300 + chunk)
            lr, ar = _call(ref, kind, v, b, q, a)
            ln, an = _call(new, kind, v, b, q, a)
            worst = max(worst, (ln - lr).abs().max().item())
            scale = max(scale, lr.abs().max().item())
            if ar is not None:
                worst_att = max(worst_att, (an - ar).abs().max().item())
            agree += (ln.argmax(1) == lr.argmax(1)).sum().item()
            total += 256
    print(f"\n{kind}: class-logit max-abs err {worst:.3e} (scale {scale:.2f}), attention {worst_att:.3e}, "
          f"argmax agreement {agree}/{total}")
    assert worst <= ABS_TOL * max(1.0, scale) and worst_att <= ABS_TOL
    # random-init class logits over 1484 / 3129 answers are nearly tied (spread ~1e-2): agreement is reported, and
    # bounded loosely; the 99.9 % criterion is asserted on the multiple-choice answers above
    assert agree / total >= 0.95


def _trainer_step(model, kind, batch, labels, opt, fused):
    """One step of the reference trainer (src/MC/trainer.py:160-256): loss / backward / flat-grad rescale + clip / Adamax."""
    v, b, q, a = batch
    for p in model.parameters():
        p.grad = None
    out, _ = _call(model, kind, v, b, q, a)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, labels, reduction="sum") / out.size(0)
    loss.backward()
    params = [p for p in model.parameters() if p.requires_grad]
    raw = [p.grad.detach().clone() for p in params]          # before the clip rescales p.grad in place
    if fused:
        opt.step(grad_denom=1.0)
    else:
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        norm = flat.norm()
        if norm > 0.25:
            flat.mul_(0.25 / (norm + 1e-6))                                     # src/utils.py:323-328
        o = 0
        for p in params:
            p.grad.copy_(flat[o:o + p.numel()].view_as(p.grad))
            o += p.numel()
        opt.step()
    return loss.item(), raw


def test_mc_model_one_trainer_step_matches_reference():
    ref, new, _ = _build_pair("mc_cti", 2)
    batch = _inputs(256, 6, 7, clone=4)
    labels = torch.zeros(256, 2, device=DEV)
    labels[torch.arange(256), torch.randint(0, 2, (256,), generator=torch.Generator().manual_seed(1)).to(DEV)] = 1.0
    p_ref = [p for p in ref.parameters() if p.requires_grad]
    p_new = [p for p in new.parameters() if p.requires_grad]
    before = [p.detach().clone() for p in p_ref]
    with torch.backends.cudnn.flags(enabled=False):       # cuDNN's GRU refuses backward in eval mode
        l_ref, g_ref = _trainer_step(ref, "mc_cti", batch, labels, torch.optim.Adamax(p_ref, lr=7e-4), fused=False)
    l_new, g_new = _trainer_step(new, "mc_cti", batch, labels, cti_b200.FusedClipAdamax(p_new, lr=7e-4, clip_norm=0.25),
                                 fused=True)
    assert abs(l_ref - l_new) <= 2e-3 * max(1.0, abs(l_ref)), (l_ref, l_new)
    # the whole flat gradient (what the clip and the optimizer see): north_star 3e-2, L2-relative
    num = sum((a_ - b_).pow(2).sum().item() for a_, b_ in zip(g_new, g_ref))
    den = sum(b_.pow(2).sum().item() for b_ in g_ref)
    print(f"\nMC trainer step: loss {l_ref:.5f} vs {l_new:.5f}; flat-gradient L2-rel err {(num / den) ** 0.5:.4f}")
    assert (num / den) ** 0.5 <= 3e-2
    # the update itself, from EQUAL gradients: rescale + clip at 0.25 + Adamax of the reference trainer
    # (src/MC/trainer.py:208-219,252-256) against FusedClipAdamax fed the reference's raw gradients.  (From each
    # model's OWN gradients the first Adamax step is lr * sign(g), which turns a 2.5 % gradient error on near-zero
    # entries into full-size sign flips -- a property of Adamax, not of the kernels; the gradient itself is held to
    # 3e-2 above.)
    new2 = [torch.nn.Parameter(p0.clone()) for p0 in before]
    for p, gr in zip(new2, g_ref):
        p.grad = gr.clone()
    cti_b200.FusedClipAdamax(new2, lr=7e-4, clip_norm=0.25).step(grad_denom=1.0)
    worst = max((a_.detach() - b_.detach()).abs().max().item() for a_, b_ in zip(new2, p_ref))
    print(f"parameter max-abs difference after one trainer step from equal gradients: {worst:.2e}")
    assert worst <= 2e-6


@pytest.mark.parametrize("h_out,k", [(None, 1), (40, 1), (20, 3)])
def test_bcnet_other_branches_against_the_reference_module(h_out, k):
    """BCNet.forward outside the h_out <= 14 single-launch case: the `h_out is None` branch (reference src/bc.py:42-47),
    the `h_net` branch (h_out > 32, :63-68) and a wide h_mat (15..32 maps), against the reference's own BCNet with the same
    state_dict, fp32 on the same GPU.  No shipped builder reaches these branches; parity only."""
    ref_env.import_reference()
    from src.bc import BCNet as RefBCNet
    torch.manual_seed(5)
    ref = RefBCNet(256, 128, 128, h_out, k=k).to(DEV).eval()
    new = cti_b200.BCNet(256, 128, 128, h_out, k=k).to(DEV).eval()
    assert list(new.state_dict().keys()) == list(ref.state_dict().keys())
    new.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(1)
    v = torch.relu(torch.randn(6, 20, 256, generator=g)).to(DEV)
    q0 = torch.tanh(torch.randn(6, 7, 128, generator=g)).to(DEV)
    qr, qn = q0.clone().requires_grad_(True), q0.clone().requires_grad_(True)
    out_r, out_n = ref(v, qr), new(v, qn)
    assert out_n.shape == out_r.shape
    scale = out_r.abs().max().item()
    assert (out_n - out_r).abs().max().item() <= 2e-2 * max(scale, 1.0)
    cot = torch.randn(out_r.shape, generator=g).to(DEV)
    (out_r * cot).sum().backward()
    (out_n * cot).sum().backward()
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-12)).item()
    # a residual-free ReLU net: bf16 operand rounding flips a few ReLU signs, 3-4 % on the input gradient in the CPU
    # emulation as well (tests/test_bf16_emulation_cpu.py); the flat parameter gradient behaves like the BAN path's
    assert rel(qn.grad, qr.grad) < 6e-2
    gr = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    gn = torch.cat([p.grad.reshape(-1) for p in new.parameters()])
    assert rel(gn, gr) < 5e-2                      # BAN at 256 rows: 3.7e-2, the CPU bf16 emulation 3.75e-2
