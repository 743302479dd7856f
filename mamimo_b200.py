"""Import shim: the package directory is named ``dl-channel-estimation-mamimo_b200`` (hyphens are not
valid in a Python module name), so this module loads it under the importable name ``mamimo_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dl-channel-estimation-mamimo_b200")
_NAME = "_mamimo_b200_pkg"


def _load(name=_NAME):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_build_module():
    """The build helper alone (does not need the compiled library)."""
    name = _NAME + "_build"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load()
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
__all__ = list(_pkg.__all__)
