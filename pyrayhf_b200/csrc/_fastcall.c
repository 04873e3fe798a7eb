/*
 * CPython shim that calls prhf_vfo_host_f64 (include/pyrayhf_b200.h) on numpy buffers without the
 * ~2 us-per-argument cost of ctypes / ndarray.ctypes.  It contains no arithmetic: it extracts the data
 * pointers through the buffer protocol and forwards them to the C ABI through a function pointer that
 * pyrayhf_b200/_cabi.py obtains from the loaded libpyrayhf_b200.so.  Optional: without it the package
 * makes the same call through ctypes.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef int (*vfo_host_fn)(void *, const double *, int, int64_t, const double *, const double *, const double *,
                           const double *, int64_t, int64_t, int, int, int, unsigned, double *, int *);

/* 1-D C-contiguous buffer of 8-byte floats ('d'); returns 0 on success */
static int get_f64(PyObject *o, Py_buffer *v, int writable) {
  if (PyObject_GetBuffer(o, v, (writable ? PyBUF_WRITABLE : 0) | PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
    PyErr_Clear();
    return -1;
  }
  if (v->ndim != 1 || v->itemsize != 8 || !v->format || v->format[0] != 'd' || v->format[1] != 0) {
    PyBuffer_Release(v);
    return -1;
  }
  return 0;
}

/* vfo_host(fn_addr, ctx_addr, freq, den, bmag, bpsi, alt, mode, n_points, flags, vh_out, status_i32)
 * -> rc of prhf_vfo_host_f64, or -1 when an argument is not a contiguous float64 vector (caller converts). */
static PyObject *fast_vfo_host(PyObject *self, PyObject *const *args, Py_ssize_t nargs) {
  (void)self;
  if (nargs != 12) {
    PyErr_SetString(PyExc_TypeError, "vfo_host expects 12 arguments");
    return NULL;
  }
  vfo_host_fn fn = (vfo_host_fn)PyLong_AsVoidPtr(args[0]);
  void *ctx = PyLong_AsVoidPtr(args[1]);
  const long mode = PyLong_AsLong(args[7]);
  const long n_points = PyLong_AsLong(args[8]);
  const unsigned long flags = PyLong_AsUnsignedLong(args[9]);
  if (PyErr_Occurred()) return NULL;
  Py_buffer b[6];
  int got = 0;
  long rc = -1;
  for (; got < 5; ++got)
    if (get_f64(args[2 + got], &b[got], 0) != 0) goto done;
  if (get_f64(args[10], &b[5], 1) != 0) goto done;
  got = 6;
  {
    Py_buffer st;
    if (PyObject_GetBuffer(args[11], &st, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) {
      PyErr_Clear();
      goto done;
    }
    const Py_ssize_t n_freq = b[0].shape[0], n_alt = b[1].shape[0];
    if (st.len >= 4 && b[2].shape[0] == n_alt && b[3].shape[0] == n_alt && b[4].shape[0] == n_alt &&
        b[5].shape[0] == n_freq && n_freq > 0 && n_alt > 0) {
      Py_BEGIN_ALLOW_THREADS
      rc = fn(ctx, (const double *)b[0].buf, (int)n_freq, 0, (const double *)b[1].buf, (const double *)b[2].buf,
              (const double *)b[3].buf, (const double *)b[4].buf, 0, 1, (int)n_alt, (int)mode, (int)n_points,
              (unsigned)flags, (double *)b[5].buf, (int *)st.buf);
      Py_END_ALLOW_THREADS
    }
    PyBuffer_Release(&st);
  }
done:
  for (int k = 0; k < got; ++k) PyBuffer_Release(&b[k]);
  return PyLong_FromLong(rc);
}

static PyMethodDef methods[] = {
    {"vfo_host", (PyCFunction)(void (*)(void))fast_vfo_host, METH_FASTCALL,
     "forward numpy float64 vectors to prhf_vfo_host_f64"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_prhf_fast", NULL, -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__prhf_fast(void) { return PyModule_Create(&moddef); }
