"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/vfo_oracle_scalar.c.

``variant=0`` literal float64 restatement, ``variant=1`` long-double "truth"
(see the C file header).  Only tests/, smoke() and bench.py's CPU legs use it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvfo_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    src = os.path.join(_HERE, "vfo_oracle_scalar.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.vfo_oracle_multiplier.argtypes = [ctypes.c_int, _dp]
        L.vfo_oracle_multiplier.restype = None
        L.vfo_oracle_profile.argtypes = [_dp, ctypes.c_int, _dp, _dp, _dp, _dp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int,
                                         ctypes.c_int, _dp, _dp]
        L.vfo_oracle_profile.restype = ctypes.c_int
        L.vfo_oracle_batch.argtypes = [_dp, ctypes.c_int, ctypes.c_int64, _dp, _dp, _dp, _dp,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, _dp, _ip]
        L.vfo_oracle_batch.restype = ctypes.c_int
        L.vfo_oracle_mup.argtypes = [_dp, _dp, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp]
        L.vfo_oracle_mup.restype = None
        L.vfo_oracle_pairwise_sum.argtypes = [_dp, ctypes.c_long]
        L.vfo_oracle_pairwise_sum.restype = ctypes.c_double
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _mode(mode):
    if mode not in ('O', 'X'):
        raise ValueError("mode must be 'O' or 'X'")
    return 0 if mode == 'O' else 1


def multiplier(n_points):
    m = np.empty(max(n_points, 0))
    lib().vfo_oracle_multiplier(n_points, _p(m))
    return m


def vertical_forward_operator(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                              variant=0, multiplier=None, n_threads=1, return_hc=False):
    """Same contract as library.py:459-509 (errors included)."""
    freq, den, bmag, bpsi, alt = map(_c, (freq, den, bmag, bpsi, alt))
    freq = freq.reshape(-1)
    m = _c(multiplier) if multiplier is not None else None
    vh = np.empty(freq.size)
    hc = np.empty(freq.size)
    st = lib().vfo_oracle_profile(_p(freq), freq.size, _p(den), _p(bmag), _p(bpsi), _p(alt),
                                  alt.size, _mode(mode), n_points, _p(m), variant, n_threads,
                                  _p(vh), _p(hc))
    if st == 1:
        raise ValueError("Density must be non-negative")
    if st == 2:
        raise IndexError("index -1 is out of bounds for axis 1 with size 0")
    return (vh, hc) if return_hc else vh


def vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                                      variant=0, multiplier=None, n_threads=0):
    """[P, A] profiles -> ([P, F] virtual heights, [P] status).  Failed profiles are NaN rows."""
    freq, den, bmag, bpsi, alt = map(_c, (freq, den, bmag, bpsi, alt))
    n_prof, n_alt = den.shape
    n_freq = freq.shape[-1]
    m = _c(multiplier) if multiplier is not None else None
    vh = np.empty((n_prof, n_freq))
    st = np.zeros(n_prof, dtype=np.int32)
    lib().vfo_oracle_batch(_p(freq), n_freq, n_freq if freq.ndim == 2 else 0, _p(den), _p(bmag),
                           _p(bpsi), _p(alt), n_alt if alt.ndim == 2 else 0, n_prof, n_alt,
                           _mode(mode), n_points, _p(m), variant, n_threads, _p(vh),
                           st.ctypes.data_as(_ip))
    return vh, st


def mup(X, Y, psi, mode='O', variant=0):
    X, Y, psi = map(_c, (X, Y, psi))
    out = np.empty(X.size)
    lib().vfo_oracle_mup(_p(X), _p(Y), _p(psi), X.size, _mode(mode), variant, _p(out))
    return out
