"""Import shim: the package directory `periodicschurdecompositions.jl_b200/` has a dot in its
name and cannot be imported with a plain `import` statement, so this module loads it with
importlib and re-exports it as `psd_b200` (also registered under its literal name)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "periodicschurdecompositions.jl_b200")
_NAME = "periodicschurdecompositions_jl_b200"


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load()
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
capi = sys.modules[_NAME + ".capi"]
pkg = _pkg
__all__ = list(_pkg.__all__) + ["capi", "pkg"]
