"""Drop-in boundary against the LIVE reference tree (build container only: /root/reference is not on the GPU
box).  After cti_b200.install() the unchanged reference builders construct their models out of the sm_100a
modules, with the reference's parameter names, shapes and order."""
import os
import sys
import types
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")),
                                reason="reference tree only exists in the build container")


def _shim_and_import():
    """torch._six / h5py stubs so src.utils imports on torch 2.x (nothing is patched in the
    reference's hot path)."""
    import collections
    import collections.abc
    warnings.filterwarnings("ignore")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    sys.modules.setdefault("torch._six", six)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    collections.Mapping, collections.Sequence = collections.abc.Mapping, collections.abc.Sequence


def _args_dataset(n_ans):
    ds = types.SimpleNamespace(dictionary=types.SimpleNamespace(ntoken=300), v_dim=2048, num_ans_candidates=n_ans)
    args = types.SimpleNamespace(op="c", num_hid=1024, gamma=2, h_mm=512, h_out=1, rank=32, k=1, activation="relu",
                                 dropout=0.5, use_counter=False, num_stacks=2)
    return args, ds


def test_reference_builders_construct_with_dropin_modules():
    _shim_and_import()
    import cti_b200
    import src.MC.base_model as mc
    import src.FFOE.base_model as ff
    args, ds = _args_dataset(3129)
    torch.manual_seed(1204)
    ref_mc = mc.build_cti(args, ds)
    ref_ban = ff.build_ban(args, ds)
    ref_ff = ff.build_cti(args, ds)
    cti_b200.install()
    try:
        torch.manual_seed(1204)
        new_mc = mc.build_cti(args, ds)
        new_ban = ff.build_ban(args, ds)
        new_ff = ff.build_cti(args, ds)
        # same RNG consumption at construction (including the reference's second a_tucker, src/tc.py:28): every
        # initial weight of all three models is bit-identical to the reference's under the same seed
        for ref, new in ((ref_mc, new_mc), (ref_ban, new_ban), (ref_ff, new_ff)):
            for (k, a_), (_, b_) in zip(ref.state_dict().items(), new.state_dict().items()):
                assert torch.equal(a_, b_), k
        assert isinstance(new_mc.v_att, cti_b200.TriAttention) and isinstance(new_mc.t_net[0], cti_b200.TCNet)
        assert isinstance(new_mc.q_prj[0], cti_b200.FCNet)
        assert isinstance(new_mc.classifier, cti_b200.SimpleClassifier)
        assert isinstance(new_ban.classifier, cti_b200.SimpleClassifier)
        assert isinstance(new_mc.q_emb, cti_b200.QuestionEmbedding) and isinstance(new_mc.ans_emb, cti_b200.QuestionEmbedding)
        assert isinstance(new_ban.v_att, cti_b200.BiAttention) and isinstance(new_ban.b_net[0], cti_b200.BCNet)
        for ref, new in ((ref_mc, new_mc), (ref_ban, new_ban), (ref_ff, new_ff)):
            rk = [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
            nk = [(k, tuple(v.shape)) for k, v in new.state_dict().items()]
            assert rk == nk
            assert [k for k, _ in ref.named_parameters()] == [k for k, _ in new.named_parameters()]
            new.load_state_dict(ref.state_dict())          # reference checkpoints load unchanged
    finally:
        cti_b200.uninstall()
    import src.tc
    assert src.tc.TCNet is not cti_b200.TCNet
