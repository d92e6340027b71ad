"""Import the unmodified reference (``src.*``) from baseline/_ref (GPU box) or /root/reference (build container).

Test / bench infrastructure, not product code.  The three stubs are the ones SURVEY.md appendix A lists -- they only
make ``src/utils.py`` importable on torch 2.x / without h5py; no line of the reference is changed.
"""
import collections
import collections.abc
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def reference_root():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "src")) and os.path.exists(os.path.join(c, "src", "tc.py")):
            return c
    return None


def import_reference():
    """Make ``import src.tc`` etc. resolve to the reference.  Returns its root, or None when it is not available."""
    root = reference_root()
    if root is None:
        return None
    warnings.filterwarnings("ignore", category=FutureWarning)
    warnings.filterwarnings("ignore", category=UserWarning)
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    sys.modules.setdefault("torch._six", six)                     # src/utils.py:18
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))      # src/dataset.py:13
    collections.Mapping, collections.Sequence = collections.abc.Mapping, collections.abc.Sequence   # src/utils.py:163
    return root


def fake_args_dataset(n_ans: int, ntoken: int = 3000, gamma: int = 2):
    """The (args, dataset) pair the reference builders read: no ``tfidf`` attribute."""
    ds = types.SimpleNamespace(dictionary=types.SimpleNamespace(ntoken=ntoken), v_dim=2048, num_ans_candidates=n_ans)
    args = types.SimpleNamespace(op="c", num_hid=1024, gamma=gamma, h_mm=512, h_out=1, rank=32, k=1, activation="relu",
                                 dropout=0.5, use_counter=False, num_stacks=2)
    return args, ds
