"""Import shim: the package directory is named ``cu-net_b200`` (not a valid identifier), so it is
loaded here under the module name ``cunet_b200``.  ``import cunet_b200`` is the supported spelling."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cu-net_b200")
_spec = importlib.util.spec_from_file_location(
    "cunet_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cunet_b200"] = _mod
_spec.loader.exec_module(_mod)
