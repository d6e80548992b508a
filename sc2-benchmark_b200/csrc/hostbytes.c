/* hostbytes.c -- CPython helper for the reference contract's `list[bytes]` (sc2bench/models/layer.py:507: encode returns one
 * `bytes` object per image).  Splitting a staging buffer into 256 bytes objects, or gathering 256 bytes objects into one, is
 * a 12 MB memcpy per batch; done in a Python loop it holds the GIL for milliseconds, and with one host thread per batch in
 * flight every kernel launch of the other threads then waits for it.  Here the objects are created under the GIL (cheap)
 * and the copies run with the GIL released.
 *
 *   split(buffer, offsets) -> list[bytes]      buffer: contiguous bytes-like; offsets: contiguous int64[n + 1]
 *   join(strings, buffer, offsets) -> total    strings: list/tuple of bytes; buffer: writable; offsets: writable int64[n + 1]
 *                                              (-1: buffer too small; raises TypeError / ValueError on malformed input)
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static PyObject *hb_split(PyObject *self, PyObject *args) {
    Py_buffer buf, offs;
    if (!PyArg_ParseTuple(args, "y*y*", &buf, &offs)) return NULL;
    PyObject *list = NULL;
    char **dst = NULL;
    const int64_t *o = (const int64_t *)offs.buf;
    const Py_ssize_t n = offs.len / 8 - 1;
    if (n < 0 || (offs.len & 7)) {
        PyErr_SetString(PyExc_ValueError, "offsets must be int64[n + 1]");
        goto done;
    }
    for (Py_ssize_t i = 0; i < n; ++i)
        if (o[i] < 0 || o[i + 1] < o[i] || o[i + 1] > buf.len) {
            PyErr_SetString(PyExc_ValueError, "offsets out of range");
            goto done;
        }
    list = PyList_New(n);
    dst = (char **)malloc(sizeof(char *) * (size_t)(n > 0 ? n : 1));
    if (!list || !dst) {
        Py_CLEAR(list);
        PyErr_NoMemory();
        goto done;
    }
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *b = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(o[i + 1] - o[i]));
        if (!b) {
            Py_CLEAR(list);
            goto done;
        }
        dst[i] = PyBytes_AS_STRING(b);
        PyList_SET_ITEM(list, i, b);
    }
    Py_BEGIN_ALLOW_THREADS
    for (Py_ssize_t i = 0; i < n; ++i) memcpy(dst[i], (const char *)buf.buf + o[i], (size_t)(o[i + 1] - o[i]));
    Py_END_ALLOW_THREADS
done:
    free(dst);
    PyBuffer_Release(&buf);
    PyBuffer_Release(&offs);
    return list;
}

static PyObject *hb_join(PyObject *self, PyObject *args) {
    PyObject *seq;
    Py_buffer buf, offs;
    if (!PyArg_ParseTuple(args, "Ow*w*", &seq, &buf, &offs)) return NULL;
    PyObject *result = NULL, *fast = NULL;
    const char **src = NULL;
    int64_t *o = (int64_t *)offs.buf;
    fast = PySequence_Fast(seq, "strings must be a list or tuple of bytes");
    if (!fast) goto done;
    {
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
        if (offs.len < (n + 1) * 8) {
            PyErr_SetString(PyExc_ValueError, "offsets buffer too small");
            goto done;
        }
        src = (const char **)malloc(sizeof(char *) * (size_t)(n > 0 ? n : 1));
        if (!src) {
            PyErr_NoMemory();
            goto done;
        }
        int64_t total = 0;
        o[0] = 0;
        for (Py_ssize_t i = 0; i < n; ++i) {
            PyObject *b = PySequence_Fast_GET_ITEM(fast, i);
            if (!PyBytes_Check(b)) {
                PyErr_SetString(PyExc_TypeError, "every string must be a bytes object");
                goto done;
            }
            src[i] = PyBytes_AS_STRING(b);
            total += (int64_t)PyBytes_GET_SIZE(b);
            o[i + 1] = total;
        }
        if (total > buf.len) {
            result = PyLong_FromLong(-1);
            goto done;
        }
        Py_BEGIN_ALLOW_THREADS
        for (Py_ssize_t i = 0; i < n; ++i) memcpy((char *)buf.buf + o[i], src[i], (size_t)(o[i + 1] - o[i]));
        Py_END_ALLOW_THREADS
        result = PyLong_FromLongLong(total);
    }
done:
    free((void *)src);
    Py_XDECREF(fast);
    PyBuffer_Release(&buf);
    PyBuffer_Release(&offs);
    return result;
}

static PyMethodDef hb_methods[] = {
    {"split", hb_split, METH_VARARGS, "split(buffer, offsets) -> list[bytes] (copies with the GIL released)"},
    {"join", hb_join, METH_VARARGS, "join(strings, buffer, offsets) -> total bytes, -1 if buffer is too small"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef hb_module = {PyModuleDef_HEAD_INIT, "_sc2_hostbytes", "list[bytes] <-> staging buffer, GIL released", -1, hb_methods};

PyMODINIT_FUNC PyInit__sc2_hostbytes(void) { return PyModule_Create(&hb_module); }
