"""Import shim: loads the package directory ``iccv19_vqa-cti_b200/`` (whose name is not a valid
Python identifier) under the module name ``cti_b200``.  ``import cti_b200`` with the repository
root on ``sys.path`` is the public entry point."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "iccv19_vqa-cti_b200")
_spec = importlib.util.spec_from_file_location("cti_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cti_b200"] = _mod
_spec.loader.exec_module(_mod)
