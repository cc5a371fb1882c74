"""Import helper: the package directory is ``ufemism2.0_b200`` (not a valid Python
identifier), so it is registered in ``sys.modules`` as ``ufemism2_0_b200``."""
import importlib.util
import os
import sys

_NAME = "ufemism2_0_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    pkg_dir = os.path.join(ROOT, "ufemism2.0_b200")
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
