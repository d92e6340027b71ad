"""The snippet of INTEGRATION.md section 3 on the whole multiple-choice model (UNMODIFIED reference builder on the
substituted modules): gradient groups covering every weight-normed layer, gradients written in place into the reducer's
slab, equal to the ungrouped run.  One GPU (no exchange); `python tools/check_integration_groups.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cti_b200  # noqa: E402
import ref_env  # noqa: E402
from cti_b200.dp import GradAllReducer  # noqa: E402

ref_env.import_reference()
import src.MC.base_model as mc  # noqa: E402

args, ds = ref_env.fake_args_dataset(2)
torch.manual_seed(1204)
cti_b200.install()
try:
    model = mc.build_cti(args, ds)
finally:
    cti_b200.uninstall()
model = model.to("cuda").eval()
g = torch.Generator().manual_seed(5)
rows = 64
v = torch.relu(torch.randn(rows, 50, 2048, generator=g)).cuda()
b = torch.rand(rows, 50, 6, generator=g).cuda()
q = torch.randint(0, 3000, (rows, 12), generator=g).cuda()
a = torch.randint(0, 3001, (rows, 6), generator=g).cuda()
y = torch.rand(rows, 2, generator=g).cuda()


def step(reducer):
    for p in model.parameters():
        p.grad = None
    cti_b200.prepack(model)
    logits, _ = model(v, b, q, a)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, y).backward()
    if reducer is not None:
        reducer.reduce_now()
    return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


ref = step(None)
groups = [[model.classifier, model.t_net[1], model.q_prj[1], model.a_prj[1]],
          [model.t_net[0], model.q_prj[0], model.a_prj[0]], [model.v_att]]
params = [p for p in model.parameters() if p.requires_grad]
reducer = GradAllReducer(params, param_groups=cti_b200.weight_norm_param_groups(model, groups), transport="peer")
reducer.set_hooks_enabled(False)
cti_b200.bind_grad_buffers(model, reducer, groups=groups)
got = step(reducer)
worst = max(((got[n] - r).norm() / (r.norm() + 1e-20)).item() for n, r in ref.items())
lo, hi = reducer.slab.data_ptr(), reducer.slab.data_ptr() + 4 * reducer.slab.numel()
inside = sum(lo <= p.grad.data_ptr() < hi for p in params if p.grad is not None)
print(f"groups ok: {len(ref)} gradients, worst relative difference {worst:.2e}, {inside} of them live in the slab")
assert worst < 1e-3          # run-to-run noise of the float-atomic column sums (biases, embeddings): 8e-5 measured
