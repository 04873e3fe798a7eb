"""ctypes binding of ``pyrayhf_b200/csrc/libpyrayhf_b200.so`` (include/pyrayhf_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C pyrayhf_b200/csrc``.
A missing library or a missing sm_100 device is a hard error: this package has no CPU path.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRHF_LIB_PATH") or os.path.join(_HERE, "csrc", "libpyrayhf_b200.so")   # env: developer variants

OK = 0
ERR_INVALID_ARG = 1
ERR_BAD_MODE = 2
ERR_CUDA = 3
ERR_NALT_TOO_LARGE = 4
ERR_NO_DEVICE = 5
FLAG_LITERAL = 1

# every symbol include/pyrayhf_b200.h declares
EXPORTED_SYMBOLS = (
    "prhf_version", "prhf_error_string", "prhf_last_cuda_error", "prhf_ctx_create",
    "prhf_ctx_destroy", "prhf_max_n_alt", "prhf_grid_multiplier_f64", "prhf_vfo_f64",
    "prhf_vfo_host_f64", "prhf_vfo_stream_f64", "prhf_host_register", "prhf_host_unregister",
    "prhf_mu_mup_f64", "prhf_measure_fp64_peak", "prhf_launch_count",
    "prhf_selftest_math", "prhf_kernel_timing", "prhf_residual_f64", "prhf_argmin_f64",
    "prhf_den2freq_f64", "prhf_find_x_f64", "prhf_find_y_f64", "prhf_smooth_grid_f64",
    "prhf_regrid_f64", "prhf_find_vh_f64", "prhf_synth_profiles_f64",
    "prhf_snell_f64", "prhf_snell_fan_f64",
)

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_u = ctypes.c_uint

_lib = None
_lock = threading.Lock()

try:                                   # optional buffer-protocol caller (pyrayhf_b200/csrc/_fastcall.c)
    from pyrayhf_b200 import _prhf_fast as _fast
except ImportError:                    # same C-ABI call through ctypes instead
    _fast = None


class PrhfError(RuntimeError):
    """A C-ABI call failed (status code in ``.code``)."""

    def __init__(self, code, text):
        super().__init__("pyrayhf_b200: %s (code %d)" % (text, code))
        self.code = code


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pyrayhf_b200: CUDA extension not built (%s missing). Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C pyrayhf_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.prhf_version.restype = _i
        L.prhf_error_string.argtypes = [_i]
        L.prhf_error_string.restype = ctypes.c_char_p
        L.prhf_last_cuda_error.argtypes = [_vp, ctypes.POINTER(ctypes.c_char_p)]
        L.prhf_last_cuda_error.restype = _i
        L.prhf_ctx_create.argtypes = [_i, ctypes.POINTER(_vp)]
        L.prhf_ctx_create.restype = _i
        L.prhf_ctx_destroy.argtypes = [_vp]
        L.prhf_ctx_destroy.restype = None
        L.prhf_max_n_alt.argtypes = [_vp]
        L.prhf_max_n_alt.restype = _i
        L.prhf_grid_multiplier_f64.argtypes = [_vp, _i, _vp, _vp]
        L.prhf_grid_multiplier_f64.restype = _i
        vfo_args = [_vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _u, _vp, _vp]
        L.prhf_vfo_f64.argtypes = vfo_args + [_vp]
        L.prhf_vfo_f64.restype = _i
        L.prhf_vfo_host_f64.argtypes = vfo_args
        L.prhf_vfo_host_f64.restype = _i
        L.prhf_vfo_stream_f64.argtypes = vfo_args[:14] + [_i64, _vp, _i64, _vp, _vp, _i]
        L.prhf_vfo_stream_f64.restype = _i
        L.prhf_host_register.argtypes = [_vp, ctypes.c_size_t]
        L.prhf_host_register.restype = _i
        L.prhf_host_unregister.argtypes = [_vp]
        L.prhf_host_unregister.restype = _i
        L.prhf_mu_mup_f64.argtypes = [_vp, _vp, _vp, _vp, _i64, _i, _i, _u, _vp, _vp, _vp]
        L.prhf_mu_mup_f64.restype = _i
        L.prhf_measure_fp64_peak.argtypes = [_vp, ctypes.POINTER(ctypes.c_double)]
        L.prhf_measure_fp64_peak.restype = _i
        L.prhf_selftest_math.argtypes = [_vp, ctypes.POINTER(ctypes.c_double)]
        L.prhf_selftest_math.restype = _i
        L.prhf_kernel_timing.argtypes = [_vp, _i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(_i)]
        L.prhf_kernel_timing.restype = _i
        L.prhf_residual_f64.argtypes = [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]
        L.prhf_residual_f64.restype = _i
        L.prhf_argmin_f64.argtypes = [_vp, _vp, _i64, _vp, _vp]
        L.prhf_argmin_f64.restype = _i
        _d = ctypes.c_double
        L.prhf_den2freq_f64.argtypes = [_vp, _vp, _i64, _vp, _vp, _vp]
        L.prhf_den2freq_f64.restype = _i
        L.prhf_find_x_f64.argtypes = [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp]
        L.prhf_find_x_f64.restype = _i
        L.prhf_find_y_f64.argtypes = [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp]
        L.prhf_find_y_f64.restype = _i
        L.prhf_smooth_grid_f64.argtypes = [_vp, _d, _d, _i, _d, _vp, _vp]
        L.prhf_smooth_grid_f64.restype = _i
        L.prhf_regrid_f64.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp]
        L.prhf_regrid_f64.restype = _i
        L.prhf_find_vh_f64.argtypes = [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _d, _i, _u, _vp, _vp]
        L.prhf_find_vh_f64.restype = _i
        L.prhf_synth_profiles_f64.argtypes = [_vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp]
        L.prhf_synth_profiles_f64.restype = _i
        L.prhf_snell_f64.argtypes = [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _u, _d, _d, _i, _d,
                                     _vp, _vp, _vp, _i, _vp, _vp]
        L.prhf_snell_f64.restype = _i
        L.prhf_snell_fan_f64.argtypes = [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _u, _d, _d, _i, _d,
                                         _vp, _vp, _vp, _i, _vp, _vp]
        L.prhf_snell_fan_f64.restype = _i
        L.prhf_launch_count.argtypes = [_vp]
        L.prhf_launch_count.restype = _i64
        _lib = L
    return _lib


class Context:
    """Owns one ``prhf_ctx`` (one CUDA device; one stream at a time)."""

    def __init__(self, device=-1):
        L = load()
        h = _vp()
        rc = L.prhf_ctx_create(int(device), ctypes.byref(h))
        if rc != OK:
            raise PrhfError(rc, L.prhf_error_string(rc).decode())
        self._h = h
        self._L = L
        self.device = device
        self.vfo_host = lambda *a, _f=L.prhf_vfo_host_f64, _h=h: _f(_h, *a)
        self.fast = _fast
        self.fn_addr = ctypes.cast(L.prhf_vfo_host_f64, ctypes.c_void_p).value
        self.ctx_addr = h.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.prhf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc == OK:
            return
        text = self._L.prhf_error_string(rc).decode()
        if rc == ERR_BAD_MODE:
            raise ValueError("mode must be 'O' or 'X'")       # library.py:396
        if rc == ERR_CUDA:
            msg = ctypes.c_char_p()
            code = self._L.prhf_last_cuda_error(self._h, ctypes.byref(msg))
            text += ": %s (cudaError %d)" % ((msg.value or b"?").decode(), code)
        raise PrhfError(rc, text)

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._L

    def max_n_alt(self):
        return self._L.prhf_max_n_alt(self._h)

    def launch_count(self):
        return self._L.prhf_launch_count(self._h)

    def selftest_math(self):
        out = (ctypes.c_double * 6)()
        self.check(self._L.prhf_selftest_math(self._h, out))
        return list(out)

    def kernel_timing(self, enable):
        """(rows_kernel_ms, tile_kernel_ms, launch_pairs) accumulated since the last call; sets the mode."""
        a, b, n = ctypes.c_double(), ctypes.c_double(), _i()
        self.check(self._L.prhf_kernel_timing(self._h, int(bool(enable)), ctypes.byref(a), ctypes.byref(b),
                                              ctypes.byref(n)))
        return a.value, b.value, n.value

    def measure_fp64_peak(self):
        out = ctypes.c_double()
        self.check(self._L.prhf_measure_fp64_peak(self._h, ctypes.byref(out)))
        return out.value


_contexts = {}
_ctx_lock = threading.Lock()


def context(device=-1):
    """Process-wide context per (device, host thread)."""
    key = (int(device), threading.get_ident())
    ctx = _contexts.get(key)
    if ctx is None:
        with _ctx_lock:
            ctx = _contexts.get(key)
            if ctx is None:
                ctx = Context(device)
                _contexts[key] = ctx
    return ctx


def host_register(addr, nbytes):
    """Page-lock a host range this process owns (cudaHostRegister, portable)."""
    L = load()
    rc = L.prhf_host_register(_vp(addr), int(nbytes))
    if rc != OK:
        raise PrhfError(rc, "cudaHostRegister failed for %d bytes" % nbytes)


def host_unregister(addr):
    load().prhf_host_unregister(_vp(addr))
