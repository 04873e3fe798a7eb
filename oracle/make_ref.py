"""TEST/BENCH INFRASTRUCTURE ONLY -- recipe that puts the REAL reference next to the oracle for the GPU box.

The reference is pure Python (no sources to compile), so "building" it means: copy
``/root/reference/PyRayHF/library.py`` -- the file that holds ``vertical_forward_operator``
(library.py:459-509) and everything it calls -- unmodified into the git-ignored directory ``oracle/_ref/PyRayHF/``,
and write three stub packages beside it for the imports at the top of that file which are absent offline and
unused on this path (``lmfit``, ``PyIRI``, ``PyIRI.sh_library``; library.py:22-25) plus a ``PyRayHF/__init__.py``
that provides only ``logger`` (library.py:37; the real ``__init__`` needs installed-package metadata).  Recipe from
SURVEY.md section 8c / BASELINE.md section 3.  ``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never
enter the history) but not in ``.gpurunignore``, so it travels to the GPU box with the snapshot, where
``bench.py``'s CPU legs time it (``cpu_baseline.kind = "reference"``).

    python oracle/make_ref.py            # run by __graft_entry__.build() when /root/reference is mounted
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("PYRAYHF_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    os.path.join("PyRayHF", "__init__.py"):
        '"""Stub written by oracle/make_ref.py: the reference package\'s logger only (PyRayHF/__init__.py:4-6)."""\n'
        'import logging\n\nlogger = logging.getLogger("PyRayHF_logger")\n',
    os.path.join("lmfit", "__init__.py"):
        '"""Stub written by oracle/make_ref.py: lmfit is absent offline and unused by vertical_forward_operator."""\n',
    os.path.join("PyIRI", "__init__.py"):
        '"""Stub written by oracle/make_ref.py: PyIRI is absent offline and unused by vertical_forward_operator."""\n',
    os.path.join("PyIRI", "sh_library.py"):
        '"""Stub written by oracle/make_ref.py (library.py imports PyIRI.sh_library at module level)."""\n',
}


def make(force=False):
    """Returns the path of oracle/_ref when the reference copy is in place, else None."""
    src = os.path.join(REFERENCE_ROOT, "PyRayHF", "library.py")
    dst = os.path.join(REF_DIR, "PyRayHF", "library.py")
    if not os.path.isfile(src):
        return REF_DIR if os.path.isfile(dst) else None
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if force or not os.path.isfile(dst) or open(src, "rb").read() != open(dst, "rb").read():
        shutil.copyfile(src, dst)
    for rel, text in _STUBS.items():
        path = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    with open(os.path.join(REF_DIR, "README"), "w") as fh:
        fh.write("PyRayHF/library.py is an unmodified copy of %s made by oracle/make_ref.py; the other files are "
                 "stubs.  Git-ignored; do not edit.\n" % src)
    return REF_DIR


def load():
    """Import the reference's ``PyRayHF.library`` from oracle/_ref (None when the copy is absent)."""
    dst = os.path.join(REF_DIR, "PyRayHF", "library.py")
    if not os.path.isfile(dst):
        return None
    mod = sys.modules.get("PyRayHF.library")
    if mod is not None and os.path.abspath(getattr(mod, "__file__", "")) == dst:
        return mod
    for name in [n for n in sys.modules if n == "PyRayHF" or n.startswith("PyRayHF.")]:
        del sys.modules[name]
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return importlib.import_module("PyRayHF.library")


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
