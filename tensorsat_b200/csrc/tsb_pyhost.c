/* tsb_pyhost.c — host-side helper of the Python binding (CPython C API; not part of the C-ABI library).
 *
 * TreeContraction.contract_from_host packs the leaves of a diagram (about a thousand arrays of 2..8 bytes for a CNF
 * network: FinRel tensors, tensorsat/lang/fin_rel.py:328-340) into ONE pinned staging buffer before the single H2D copy.
 * Done with bytes.join that costs ~70 us per contraction, more than the device spends on a twentieth of the plan;
 * here it is one pass over the sequence (numpy arrays through the numpy C API, anything else through the buffer
 * protocol), with no intermediate Python objects.
 *
 *   pack_into(leaves, dst) -> int   bytes written, or -1 if a leaf is not a contiguous buffer / dst is too small
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>

static PyObject* pack_into(PyObject* self, PyObject* args) {
  PyObject *leaves, *dst_obj;
  if (!PyArg_ParseTuple(args, "OO", &leaves, &dst_obj)) return NULL;
  Py_buffer dst;
  if (PyObject_GetBuffer(dst_obj, &dst, PyBUF_WRITABLE | PyBUF_SIMPLE) != 0) return NULL;
  PyObject* seq = PySequence_Fast(leaves, "leaves must be a sequence");
  if (!seq) { PyBuffer_Release(&dst); return NULL; }
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
  PyObject** items = PySequence_Fast_ITEMS(seq);
  char* out = (char*)dst.buf;
  Py_ssize_t used = 0;
  for (Py_ssize_t i = 0; i < n; i++) {
    if (PyArray_Check(items[i])) {   /* numpy arrays directly: the buffer protocol costs ~70 ns per array */
      PyArrayObject* a = (PyArrayObject*)items[i];
      const Py_ssize_t len = (Py_ssize_t)PyArray_NBYTES(a);
      if (!PyArray_IS_C_CONTIGUOUS(a) || used + len > dst.len) { used = -1; break; }
      memcpy(out + used, PyArray_DATA(a), (size_t)len);
      used += len;
      continue;
    }
    Py_buffer v;
    if (PyObject_GetBuffer(items[i], &v, PyBUF_SIMPLE) != 0) {   /* not a buffer, or not C-contiguous */
      PyErr_Clear();
      used = -1;
      break;
    }
    if (used + v.len > dst.len) { PyBuffer_Release(&v); used = -1; break; }
    memcpy(out + used, v.buf, (size_t)v.len);
    used += v.len;
    PyBuffer_Release(&v);
  }
  Py_DECREF(seq);
  PyBuffer_Release(&dst);
  return PyLong_FromSsize_t(used);
}

static PyMethodDef methods[] = {{"pack_into", pack_into, METH_VARARGS, "pack_into(leaves, dst) -> bytes written or -1"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_tsb_host", NULL, -1, methods};
PyMODINIT_FUNC PyInit__tsb_host(void) {
  import_array();
  return PyModule_Create(&module);
}
