"""Import shim: the product package lives in the directory `interactive-rate-tendons_b200/`
(the name the project layout prescribes), which is not a valid Python identifier.  `import
irt_b200` loads that directory as the package `irt_b200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "interactive-rate-tendons_b200")
_spec = importlib.util.spec_from_file_location(
    "irt_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["irt_b200"] = _mod
_spec.loader.exec_module(_mod)
