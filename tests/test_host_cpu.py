"""CPU-side checks of the product's host logic (no GPU, no compute calls):
the C-ABI library builds, loads and exports every symbol include/cti_sm100.h declares; the
drop-in modules have the reference's parameter names / shapes / order; the packed-core index map
equals the oracle's T_eff; gradient-bucket logic of the data-parallel reducer (gloo, world 2)."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cti_b200  # noqa: E402
from oracle import cti_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, os.path.join(ROOT, "iccv19_vqa-cti_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("cti_build", os.path.join(ROOT, "iccv19_vqa-cti_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = mod.build()                                  # no-op when up to date; nvcc cross-compiles without a GPU
    return ctypes.CDLL(path)


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cti_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cti_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = header_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cti_sm100.h but not exported"


def test_ctypes_signatures_cover_the_header():
    from cti_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_version_and_error_string(lib):
    lib.cti_version.restype = ctypes.c_int
    lib.cti_last_error.restype = ctypes.c_char_p
    assert lib.cti_version() >= 100
    assert isinstance(lib.cti_last_error(), bytes)


def test_argument_errors_do_not_launch(lib):
    # negative return = argument error, nothing launched, message recorded (no GPU needed)
    lib.cti_masked_softmax_fwd.restype = ctypes.c_int
    lib.cti_masked_softmax_fwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                           ctypes.c_void_p]
    rc = lib.cti_masked_softmax_fwd(None, None, 4, 0, None)
    assert rc < 0
    lib.cti_last_error.restype = ctypes.c_char_p
    assert b"softmax" in lib.cti_last_error()


def test_tri_attention_state_dict_matches_reference_keys(golden):
    g = golden["tri_d16"]
    c = g["cfg"]
    m = cti_b200.TriAttention(c["v_dim"], c["q_dim"], c["q_dim"], c["h_mm"], 1, c["rank"], c["G"], 1)
    ref = g["att_sd"]
    assert list(m.state_dict().keys()) == list(ref.keys())
    for k, t in m.state_dict().items():
        assert t.shape == ref[k].shape, k
    m.load_state_dict(ref)
    assert [k for k, _ in m.named_parameters()] == list(g["att_grads"].keys())
    pool = cti_b200.TCNet(c["v_dim"], c["q_dim"], c["q_dim"], c["h_mm"], 1, c["rank"], 1, k=c["k_pool"])
    assert list(pool.state_dict().keys()) == list(g["pool_sd"][0].keys())
    pool.load_state_dict(g["pool_sd"][0])


def test_bi_attention_state_dict_matches_reference_keys(golden):
    g = golden["bi_small"]
    c = g["cfg"]
    m = cti_b200.BiAttention(c["v_dim"], c["q_dim"], c["hid"], c["G"])
    assert list(m.state_dict().keys()) == list(g["att_sd"].keys())
    m.load_state_dict(g["att_sd"])
    assert [k for k, _ in m.named_parameters()] == list(g["att_grads"].keys())
    b = cti_b200.BCNet(c["v_dim"], c["q_dim"], c["hid"], None, k=1)
    assert list(b.state_dict().keys()) == list(g["pool_sd"][0].keys())


def test_fcnet_layout_matches_reference(golden):
    for name in ("fc_relu", "fc_lin", "fc_nodrop"):
        g = golden[name]
        m = cti_b200.FCNet([24, 40], g["act"], g["dropout"])
        assert list(m.state_dict().keys()) == list(g["sd"].keys())
        m.load_state_dict(g["sd"])


def test_fcnet_init_statistics():
    torch.manual_seed(0)
    m = cti_b200.FCNet([256, 128])
    lin = m.main[0]
    assert lin.weight_g.shape == () and torch.allclose(lin.weight_g, lin.weight_v.norm())
    assert lin.weight_v.abs().max() <= 1 / 16 + 1e-6 and lin.bias.abs().max() <= 1 / 16 + 1e-6


@pytest.mark.parametrize("R,d,G", [(2, 4, 1), (3, 4, 2), (2, 16, 2), (2, 16, 3)])
def test_packed_core_index_equals_oracle_teff(R, d, G):
    from cti_b200 import functions as F_
    torch.manual_seed(1)
    tg = torch.randn(1, R, d, d, d, G, 1)
    te = O.teff_from_tg(tg)                                          # (R,i,j,l,g)
    idx = F_.tpack_index(R, d, G, "cpu")
    tpack = tg.reshape(-1)[idx].view(R, d, d, G, d)                  # [r][l][i][g][j]
    assert torch.equal(tpack, te.permute(0, 3, 1, 4, 2))
    assert torch.equal(torch.sort(idx).values, torch.arange(idx.numel()))       # a permutation: grads scatter back 1:1


def test_modules_refuse_cpu_tensors():
    m = cti_b200.FCNet([16, 16]).eval()
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 16))
    t = cti_b200.TriAttention(64, 48, 48, 64, 1, 4, 2, 1).eval()
    with pytest.raises(RuntimeError):
        t(torch.randn(2, 5, 64), torch.randn(2, 3, 48), torch.randn(2, 2, 48))


def test_unsupported_domain_raises_like_the_reference():
    t = cti_b200.TriAttention(64, 48, 48, 64, 1, 4, 1, 1).eval()          # glimpse = 1
    with pytest.raises(RuntimeError):
        t(torch.randn(2, 5, 64), torch.randn(2, 3, 48), torch.randn(2, 2, 48))
    pool = cti_b200.TCNet(64, 48, 48, 512, 1, 32, 1, k=2)                 # h_dim * k >= 1024: no T_g / per-rank nets
    assert not hasattr(pool, "T_g") and not hasattr(pool, "v_net")
    with pytest.raises(RuntimeError):
        pool(torch.randn(2, 5, 64), torch.randn(2, 3, 48), torch.randn(2, 2, 48))


def test_split_heuristic_fills_waves():
    from cti_b200 import functions as F_
    for tiles in (1, 4, 16, 32, 64, 100, 148, 300):
        s = F_._pick_splits(tiles, 800)
        units = tiles * s
        assert 1 <= s <= 32                                                # at most 32 splits
        if 4 <= tiles < 148:                                               # (a single tile cannot fill 148 SMs with 32 splits)
            assert units / (-(-units // 148) * 148) >= 0.8
    assert F_._pick_splits(4, 4) == 1                                      # too few k-blocks to split


# --------------------------------------------------------------------------- #
# the header is a C interface; the rows next to the path keep the reference's parameter names
# --------------------------------------------------------------------------- #
def test_header_is_plain_c(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "cti_sm100.h"\nint main(void) { int (*f)(void) = cti_version; (void)f; return 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


NEXT = os.path.join(ROOT, "tests", "golden", "next_golden.pt")


def test_next_row_modules_keep_reference_state_dict_keys():
    import types
    g = torch.load(NEXT)
    args = types.SimpleNamespace(activation="relu", dropout=0.5)
    i, h, o = g["clf_mc"]["dims"]
    clf = cti_b200.SimpleClassifier(i, h, o, args)
    assert [(k, tuple(v.shape)) for k, v in clf.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in g["clf_mc"]["sd"].items()]
    din, hid = g["gru_small"]["dims"]
    qe = cti_b200.QuestionEmbedding(din, hid, 1, False, .0)
    assert [(k, tuple(v.shape)) for k, v in qe.state_dict().items()] == \
        [(k, tuple(v.shape)) for k, v in g["gru_small"]["sd"].items()]
    assert qe.init_hidden(3).shape == (1, 3, hid)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from mc_model import BanStudent, MCModel
    for cls, key in ((MCModel, "mc_model"), (BanStudent, "ban_model")):
        m = cls(**g[key]["args"])
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == \
            [(k, tuple(v.shape)) for k, v in g[key]["sd"].items()], key
        m.load_state_dict(g[key]["sd"])                     # the reference's checkpoint loads as is


def test_next_row_modules_refuse_cpu_and_unsupported_variants():
    qe = cti_b200.QuestionEmbedding(24, 32, 1, False, .0)
    with pytest.raises(RuntimeError, match="CUDA"):
        qe.forward_all(torch.randn(2, 3, 24))
    with pytest.raises(NotImplementedError):
        cti_b200.QuestionEmbedding(24, 32, 1, False, .0, rnn_type="LSTM")
    with pytest.raises(RuntimeError, match="CUDA"):
        cti_b200.FusedClipAdamax([torch.nn.Parameter(torch.zeros(4))])
    with pytest.raises(AssertionError):
        import types
        cti_b200.SimpleClassifier(8, 16, 2, types.SimpleNamespace(activation="gelu", dropout=0.1))


def test_reduce_now_leaves_gradients_on_the_bucket_views_and_remembers_sources():
    """Hook-free path of the reducer (what follows a CUDA-graph replay), single process."""
    from cti_b200.dp import GradAllReducer
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 3, 7)]
    red = GradAllReducer(ps, bucket_bytes=16)
    red.set_hooks_enabled(False)
    src = [torch.full((n,), float(i + 1)) for i, n in enumerate((5, 3, 7))]
    for p, s in zip(ps, src):
        p.grad = s
    red.reduce_now()
    assert red.slab is not None and red.slab.numel() == 8 + 4 + 8       # every view starts on a 16-byte boundary
    for p, s in zip(ps, src):
        assert torch.equal(p.grad, s) and p.grad.data_ptr() != s.data_ptr()          # now a view of the slab
    for s in src:                                                                    # a "replay" refills the same tensors
        s.mul_(3.0)
    red.reduce_now()
    for i, p in enumerate(ps):
        assert torch.equal(p.grad, torch.full_like(p.grad, 3.0 * (i + 1)))
    ps[1].grad = torch.full((3,), -1.0)                                              # an eager backward takes over
    red.reduce_now()
    assert torch.equal(ps[1].grad, torch.full((3,), -1.0))
