"""Import shim: registers the package directory `multi-modal-tracking_b200/` (not a valid Python
identifier) as the importable package `mmt_b200`.  `import mmt_b200` then `mmt_b200.<submodule>`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi-modal-tracking_b200")
_spec = importlib.util.spec_from_file_location(
    "mmt_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mmt_b200"] = _mod
_spec.loader.exec_module(_mod)
