/*
 * _qcut_pyhelper -- CPython-side walk of a SamplerV2 result (binding plumbing, no arithmetic).
 *
 * reconstruct_expectation_values() receives O(10^5) PubResult objects; reading
 * result[idx].data.observable_measurements.array / .qpd_measurements.array and the address and
 * shape of every array from Python costs microseconds per PUB (reference: the same attribute walk
 * at cutting_reconstruction.py:152-155).  This helper does the walk through the C API and the
 * buffer protocol and fills flat tables for qcut_stage_h2d.  It is loaded with ctypes.PyDLL (GIL
 * held) and resolves the Python C-API symbols from the running interpreter.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* Returns the number of PUBs collected (hi - lo), or -1 with a Python exception set.
 * flags[j] != 0 marks a PUB whose arrays need the slow path (wrong dtype, not C-contiguous, ndim != 2,
 * or observable rows != qpd rows); its table entries are then undefined.
 * Every array object is appended to `keep` (obs, qpd, obs, qpd, ...) so the addresses stay valid. */
int64_t qcut_py_collect(PyObject* result, int64_t lo, int64_t hi, uint64_t* obs_ptr, uint64_t* qpd_ptr,
                        int64_t* shots, int64_t* nb_obs, int64_t* nb_qpd, uint8_t* flags, PyObject* keep) {
  static PyObject *s_data = NULL, *s_obs = NULL, *s_qpd = NULL, *s_array = NULL;
  if (!s_data) {
    s_data = PyUnicode_InternFromString("data");
    s_obs = PyUnicode_InternFromString("observable_measurements");
    s_qpd = PyUnicode_InternFromString("qpd_measurements");
    s_array = PyUnicode_InternFromString("array");
    if (!s_data || !s_obs || !s_qpd || !s_array) return -1;
  }
  for (int64_t idx = lo; idx < hi; ++idx) {
    const int64_t j = idx - lo;
    PyObject* key = PyLong_FromLongLong(idx);
    if (!key) return -1;
    PyObject* pub = PyObject_GetItem(result, key);
    Py_DECREF(key);
    if (!pub) return -1;
    PyObject* data = PyObject_GetAttr(pub, s_data);
    Py_DECREF(pub);
    if (!data) return -1;
    PyObject* regs[2] = {PyObject_GetAttr(data, s_obs), NULL};
    if (regs[0]) regs[1] = PyObject_GetAttr(data, s_qpd);
    Py_DECREF(data);
    if (!regs[0] || !regs[1]) { Py_XDECREF(regs[0]); Py_XDECREF(regs[1]); return -1; }
    PyObject* arr[2] = {PyObject_GetAttr(regs[0], s_array), NULL};
    if (arr[0]) arr[1] = PyObject_GetAttr(regs[1], s_array);
    Py_DECREF(regs[0]);
    Py_DECREF(regs[1]);
    if (!arr[0] || !arr[1]) { Py_XDECREF(arr[0]); Py_XDECREF(arr[1]); return -1; }
    flags[j] = 0;
    int64_t rows[2] = {0, 0}, cols[2] = {0, 0};
    uint64_t ptr[2] = {0, 0};
    for (int r = 0; r < 2; ++r) {
      Py_buffer view;
      if (PyObject_GetBuffer(arr[r], &view, PyBUF_STRIDES | PyBUF_FORMAT) != 0) {
        PyErr_Clear();
        flags[j] = 1;
        continue;
      }
      const int ok = view.ndim == 2 && view.itemsize == 1 && view.format && view.format[0] == 'B' && view.format[1] == 0 &&
                     view.strides[1] == 1 && (view.shape[0] <= 1 || view.strides[0] == view.shape[1]);
      if (ok) {
        ptr[r] = (uint64_t)(uintptr_t)view.buf;
        rows[r] = view.shape[0];
        cols[r] = view.shape[1];
      } else {
        flags[j] = 1;
      }
      PyBuffer_Release(&view);
    }
    if (!flags[j] && rows[0] != rows[1]) flags[j] = 1;
    obs_ptr[j] = ptr[0];
    qpd_ptr[j] = ptr[1];
    shots[j] = rows[1];
    nb_obs[j] = cols[0];
    nb_qpd[j] = cols[1];
    const int rc = PyList_Append(keep, arr[0]) || PyList_Append(keep, arr[1]);
    Py_DECREF(arr[0]);
    Py_DECREF(arr[1]);
    if (rc) return -1;
  }
  return hi - lo;
}
