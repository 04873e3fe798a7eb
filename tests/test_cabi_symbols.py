"""CPU: the C-ABI library loads and exports every symbol include/pyrayhf_b200.h declares.

No compute call is made (there is no GPU here); what is checked is the boundary itself and
that the product path fails loudly, with no CPU fallback, when no B200 is present.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from pyrayhf_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        entry.build()
    return _cabi.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "pyrayhf_b200.h")).read()
    declared = set(re.findall(r"\b(prhf_[a-z0-9_]+)\s*\(", header))
    declared.discard("prhf_ctx")
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None, name


def test_version_and_error_strings(lib):
    assert lib.prhf_version() == 100
    assert lib.prhf_error_string(2) == b"mode must be 'O' or 'X'"
    assert lib.prhf_error_string(0) == b"ok"


def test_sass_is_sm100a_only():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pyrayhf_b200
    from pyrayhf_b200 import synth
    den, bmag, bpsi, alt = synth.single_day_profile()
    with pytest.raises(_cabi.PrhfError) as ei:
        pyrayhf_b200.vertical_forward_operator(np.array([2.0]), den, bmag, bpsi, alt, 'X', 50)
    assert ei.value.code == _cabi.ERR_NO_DEVICE
    # argument errors that the reference raises before any arithmetic still surface
    with pytest.raises(ValueError, match="mode must be 'O' or 'X'"):
        pyrayhf_b200.vertical_forward_operator(np.array([2.0]), den, bmag, bpsi, alt, 'x', 50)
    # the reference needs an ndarray for freq (freq * 1e6, f.size): lists and Python scalars fail the same way
    with pytest.raises(TypeError, match="can't multiply sequence"):
        pyrayhf_b200.vertical_forward_operator([2.0, 3.0], den, bmag, bpsi, alt, 'X', 50)
    with pytest.raises(AttributeError, match="has no attribute 'size'"):
        pyrayhf_b200.vertical_forward_operator(2.0, den, bmag, bpsi, alt, 'X', 50)


def test_product_does_not_import_oracle():
    """The shipped package never imports, loads or links anything under oracle/."""
    pkg = os.path.join(ROOT, "pyrayhf_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|libvfo_oracle|vfo_oracle|oracle/_build|oracle\.scalar",
                     re.MULTILINE)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(text), fn
