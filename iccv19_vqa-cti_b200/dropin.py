"""Substitute the accelerated classes into the reference's module namespace so that
``src/MC/base_model.py`` and ``src/FFOE/base_model.py`` build their models unchanged
(the builders import the classes by name from ``src.fc``, ``src.tc``,
``src.bc`` and ``src.attention``; reference src/MC/base_model.py:10-15)."""
from __future__ import annotations

import importlib
import sys

_SAVED = {}

_TARGETS = {
    "src.fc": ("FCNet",),
    "src.tc": ("TCNet",),
    "src.bc": ("BCNet",),
    "src.attention": ("BiAttention", "TriAttention"),
    "src.classifier": ("SimpleClassifier",),
    "src.language_model": ("QuestionEmbedding",),
    "src.loss_function": ("Distillation_Loss",),
}
# modules that did ``from src.x import Name`` and hold their own binding
_IMPORTERS = ("src.MC.base_model", "src.FFOE.base_model", "src.attention", "src.tc", "src.bc", "src.FFOE.train",
              "src.FFOE.main")


def install(hot_path_only: bool = False) -> None:
    """Patch the reference package (must be importable as ``src``) to use the sm_100a modules.
    hot_path_only: substitute only the classes the north_star names (FCNet, TCNet, TriAttention, BCNet, BiAttention) and
    leave the rows next to the path -- the GRU ``QuestionEmbedding`` and ``SimpleClassifier`` -- on the reference's own
    fp32 implementations."""
    from . import attention, bc, classifier, fc, language_model, loss_function, tc
    ours = {"FCNet": fc.FCNet, "TCNet": tc.TCNet, "BCNet": bc.BCNet, "BiAttention": attention.BiAttention,
            "TriAttention": attention.TriAttention}
    if not hot_path_only:
        ours.update({"SimpleClassifier": classifier.SimpleClassifier, "QuestionEmbedding": language_model.QuestionEmbedding,
                     "Distillation_Loss": loss_function.Distillation_Loss})
    for modname, names in _TARGETS.items():
        mod = importlib.import_module(modname)
        for n in names:
            if n not in ours:
                continue
            _SAVED.setdefault((modname, n), getattr(mod, n))
            setattr(mod, n, ours[n])
    for modname in _IMPORTERS:
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for n, cls in ours.items():
            if hasattr(mod, n) and getattr(mod, n) is not cls:
                _SAVED.setdefault((modname, n), getattr(mod, n))
                setattr(mod, n, cls)


def uninstall() -> None:
    for (modname, n), cls in list(_SAVED.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, n, cls)
    _SAVED.clear()
