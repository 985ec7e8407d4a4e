/*
 * fastcall.c -- CPython binding of the ONE C-ABI call an unmodified model makes once per GRU
 * iteration (smoe_corr_lookup, include/smoe_corr.h).  ctypes spends ~2 us converting the eight
 * arguments of that call; the kernel it launches runs ~3 us, so on the eager (no CUDA graph) path
 * the host was the bottleneck (tools/eager_overhead.py).  This module does the same call with
 * METH_FASTCALL argument parsing (~0.2 us).  It holds no state besides the function pointer it is
 * given at start-up (the address ctypes resolved in the already-loaded libsmoecorr.so), links
 * against nothing but libpython, and is optional: without it the package calls through ctypes.
 *
 *   _smoe_fastcall.bind(addr_of_smoe_corr_lookup)
 *   rc = _smoe_fastcall.lookup(pyr_addr, coords_ptr, coords_batch_stride, radius, out_ptr,
 *                              out_dtype, stream)          # coord_scale = 1
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef int (*lookup_fn_t)(const void* pyr, const float* coords, int64_t coords_batch_stride,
                           float coord_scale, int radius, void* out, int out_dtype, void* stream);
static lookup_fn_t g_lookup = NULL;

static PyObject* fc_bind(PyObject* self, PyObject* arg) {
  (void)self;
  void* p = PyLong_AsVoidPtr(arg);
  if (p == NULL && PyErr_Occurred()) return NULL;
  g_lookup = (lookup_fn_t)p;
  Py_RETURN_NONE;
}

static PyObject* fc_lookup(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  (void)self;
  if (nargs != 7) {
    PyErr_SetString(PyExc_TypeError, "lookup(pyr, coords, bstride, radius, out, out_dtype, stream)");
    return NULL;
  }
  if (g_lookup == NULL) {
    PyErr_SetString(PyExc_RuntimeError, "_smoe_fastcall.bind() has not been called");
    return NULL;
  }
  const void* pyr = PyLong_AsVoidPtr(args[0]);
  const float* coords = (const float*)PyLong_AsVoidPtr(args[1]);
  const long long bstride = PyLong_AsLongLong(args[2]);
  const long radius = PyLong_AsLong(args[3]);
  void* out = PyLong_AsVoidPtr(args[4]);
  const long out_dtype = PyLong_AsLong(args[5]);
  void* stream = PyLong_AsVoidPtr(args[6]);
  if (PyErr_Occurred()) return NULL;
  const int rc = g_lookup(pyr, coords, (int64_t)bstride, 1.0f, (int)radius, out, (int)out_dtype, stream);
  return PyLong_FromLong(rc);
}

static PyMethodDef fc_methods[] = {
    {"bind", (PyCFunction)fc_bind, METH_O, "bind(address of smoe_corr_lookup)"},
    {"lookup", (PyCFunction)(void (*)(void))fc_lookup, METH_FASTCALL,
     "lookup(pyr_addr, coords_ptr, coords_batch_stride, radius, out_ptr, out_dtype, stream) -> rc"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef fc_module = {PyModuleDef_HEAD_INIT, "_smoe_fastcall",
                                       "fast binding of smoe_corr_lookup", -1, fc_methods,
                                       NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__smoe_fastcall(void) { return PyModule_Create(&fc_module); }
