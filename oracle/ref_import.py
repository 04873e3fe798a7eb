"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (container only).

Imports ``/root/reference/PyRayHF/library.py`` by file path with empty stub
modules for the third-party packages that are absent offline (``lmfit``,
``PyIRI``, ``PyIRI.sh_library``) and a stub ``PyRayHF`` package exposing
``logger`` (the reference does ``from PyRayHF import logger``, library.py:37).
Recipe from SURVEY.md section 8c.

The reference mount does not exist on the GPU box, so nothing under ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``tests/make_golden.py`` (fixture generation) and by CPU tests that are
skipped when the mount is absent.
"""
import importlib.util
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PYRAYHF_REFERENCE_ROOT", "/root/reference")
_LIB = os.path.join(REFERENCE_ROOT, "PyRayHF", "library.py")


def reference_available():
    return os.path.isfile(_LIB)


def load_reference_library():
    """Return the reference ``PyRayHF.library`` module object (cached)."""
    if "PyRayHF.library" in sys.modules and getattr(
            sys.modules["PyRayHF.library"], "__file__", "") == _LIB:
        return sys.modules["PyRayHF.library"]
    if not reference_available():
        raise FileNotFoundError(_LIB)
    for name in ("lmfit", "PyIRI", "PyIRI.sh_library"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["PyIRI"].sh_library = sys.modules["PyIRI.sh_library"]
    pkg = types.ModuleType("PyRayHF")
    pkg.__path__ = [os.path.dirname(_LIB)]
    pkg.logger = logging.getLogger("PyRayHF_logger")
    sys.modules["PyRayHF"] = pkg
    spec = importlib.util.spec_from_file_location("PyRayHF.library", _LIB)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["PyRayHF.library"] = mod
    spec.loader.exec_module(mod)
    pkg.library = mod
    return mod


def load_tutorial_fixture(which):
    """Unpickle ``Example_Input_{Day,Night}.p`` (numpy-only pickles)."""
    import pickle
    path = os.path.join(REFERENCE_ROOT, "docs", "tutorials",
                        "Example_Input_%s.p" % which)
    with open(path, "rb") as fh:
        return pickle.load(fh)
