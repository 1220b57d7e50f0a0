// sa_marshal.cpp -- CPython extension: list[FieldElement] <-> packed 16-byte
// little-endian limbs (the boundary marshalling of SURVEY.md section 2.1, K9).
//
// The reference's value type is algebra.FieldElement(value: int, field: Field)
// (code/algebra.py:14-17).  `unpack` builds instances WITHOUT calling __init__ and
// sets `value` then `field`, the same attribute order __init__ uses, so pickles of
// the results (code/ip.py:18-22, the Fiat-Shamir transcript) are byte-identical
// to pickles of elements the reference itself created.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -I<python include> sa_marshal.cpp -o sa_marshal<ext suffix>
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <cstdint>
#include <cstring>

static PyObject *s_value = nullptr;
static PyObject *s_field = nullptr;

static int long_to_16(PyObject *v, unsigned char *dst) {
    if (!PyLong_Check(v)) {
        PyErr_SetString(PyExc_TypeError, "field element value must be an int");
        return -1;
    }
#if PY_VERSION_HEX >= 0x030D0000
    return _PyLong_AsByteArray((PyLongObject *)v, dst, 16, 1, 0, 1);
#else
    return _PyLong_AsByteArray((PyLongObject *)v, dst, 16, 1, 0);
#endif
}

// pack(seq) -> bytearray; items are ints or objects with an int attribute `value`
static PyObject *sa_pack(PyObject *, PyObject *arg) {
    PyObject *seq = PySequence_Fast(arg, "pack() needs a sequence");
    if (!seq) return nullptr;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject *out = PyByteArray_FromStringAndSize(nullptr, n * 16);
    if (!out) {
        Py_DECREF(seq);
        return nullptr;
    }
    unsigned char *dst = (unsigned char *)PyByteArray_AS_STRING(out);
    PyObject **items = PySequence_Fast_ITEMS(seq);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *it = items[i];
        int rc;
        if (PyLong_Check(it)) {
            rc = long_to_16(it, dst + 16 * i);
        } else {
            PyObject *v = PyObject_GetAttr(it, s_value);
            if (!v) {
                rc = -1;
            } else {
                rc = long_to_16(v, dst + 16 * i);
                Py_DECREF(v);
            }
        }
        if (rc < 0) {
            Py_DECREF(out);
            Py_DECREF(seq);
            return nullptr;
        }
    }
    Py_DECREF(seq);
    return out;
}

// unpack_ints(buffer) -> list[int]
static PyObject *sa_unpack_ints(PyObject *, PyObject *arg) {
    Py_buffer buf;
    if (PyObject_GetBuffer(arg, &buf, PyBUF_SIMPLE) < 0) return nullptr;
    const Py_ssize_t n = buf.len / 16;
    PyObject *out = PyList_New(n);
    // a million new instances would trigger the cyclic collector again and again, and every pass walks
    // all of them (40 % of the call at 2^20): pause it while the list is filled
    const int gc_was_on = PyGC_Disable();
    if (out) {
        const unsigned char *src = (const unsigned char *)buf.buf;
        for (Py_ssize_t i = 0; i < n; i++) {
            PyObject *v = _PyLong_FromByteArray(src + 16 * i, 16, 1, 0);
            if (!v) {
                Py_CLEAR(out);
                break;
            }
            PyList_SET_ITEM(out, i, v);
        }
    }
    if (gc_was_on) PyGC_Enable();
    PyBuffer_Release(&buf);
    return out;
}

// unpack(buffer, field, cls) -> list[cls] with .value, .field set (no __init__ call)
static PyObject *sa_unpack(PyObject *, PyObject *args) {
    PyObject *bufobj, *field, *cls;
    if (!PyArg_ParseTuple(args, "OOO", &bufobj, &field, &cls)) return nullptr;
    if (!PyType_Check(cls)) {
        PyErr_SetString(PyExc_TypeError, "unpack(): third argument must be the FieldElement class");
        return nullptr;
    }
    PyTypeObject *tp = (PyTypeObject *)cls;
    Py_buffer buf;
    if (PyObject_GetBuffer(bufobj, &buf, PyBUF_SIMPLE) < 0) return nullptr;
    const Py_ssize_t n = buf.len / 16;
    PyObject *out = PyList_New(n);
    // a million new instances would trigger the cyclic collector again and again, and every pass walks
    // all of them (40 % of the call at 2^20): pause it while the list is filled
    const int gc_was_on = PyGC_Disable();
    if (out) {
        const unsigned char *src = (const unsigned char *)buf.buf;
        for (Py_ssize_t i = 0; i < n; i++) {
            PyObject *v = _PyLong_FromByteArray(src + 16 * i, 16, 1, 0);
            PyObject *obj = v ? tp->tp_alloc(tp, 0) : nullptr;
            if (!obj || PyObject_SetAttr(obj, s_value, v) < 0 || PyObject_SetAttr(obj, s_field, field) < 0) {
                Py_XDECREF(v);
                Py_XDECREF(obj);
                Py_CLEAR(out);
                break;
            }
            Py_DECREF(v);
            PyList_SET_ITEM(out, i, obj);
        }
    }
    if (gc_was_on) PyGC_Enable();
    PyBuffer_Release(&buf);
    return out;
}

static PyMethodDef methods[] = {
    {"pack", sa_pack, METH_O, "pack(seq of int | FieldElement) -> bytearray (16 bytes LE per element)"},
    {"unpack", sa_unpack, METH_VARARGS, "unpack(buffer, field, FieldElement) -> list[FieldElement]"},
    {"unpack_ints", sa_unpack_ints, METH_O, "unpack_ints(buffer) -> list[int]"},
    {nullptr, nullptr, 0, nullptr}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "sa_marshal",
                                       "list[FieldElement] <-> packed limbs", -1, methods};

PyMODINIT_FUNC PyInit_sa_marshal(void) {
    s_value = PyUnicode_InternFromString("value");
    s_field = PyUnicode_InternFromString("field");
    return PyModule_Create(&moduledef);
}
