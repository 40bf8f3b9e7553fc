/* pyhost.c -- host-side glue for the PYTHON host only (not part of the C ABI in include/asb200.h, which stays free of
 * Python types): walks comparelist2's records -- [id, SEQ, tag, idx] lists, amplicon_sorter.py:551-561 -- once, in C,
 * and hands back what host.process_list needs of them: the idx keys and, for every SEQ, pointer and length of its
 * bytes.  The reads then go to the GPU straight from the str objects (asb_upload_reads_scattered): no 100 MB
 * "".join, no second copy.  Loaded with ctypes.PyDLL (the GIL is held); CPython's public API only.
 * Returns 0 = ok, 1 = something this fast path does not handle (a record that is not a list / tuple of >= 4 items,
 * a SEQ that is not an ASCII str, an idx that is not an int64) -- the caller then takes the generic Python path. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

int asbpy_collect_records(PyObject *recs, int64_t *keys, const char **ptrs, uint32_t *lens)
{
    if (!recs || !PyList_Check(recs)) return 1;
    const Py_ssize_t n = PyList_GET_SIZE(recs);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *r = PyList_GET_ITEM(recs, i), *seq, *idx;
        if (PyList_Check(r) && PyList_GET_SIZE(r) >= 4) { seq = PyList_GET_ITEM(r, 1); idx = PyList_GET_ITEM(r, 3); }
        else if (PyTuple_Check(r) && PyTuple_GET_SIZE(r) >= 4) { seq = PyTuple_GET_ITEM(r, 1); idx = PyTuple_GET_ITEM(r, 3); }
        else return 1;
        if (!PyUnicode_Check(seq) || !PyLong_Check(idx)) return 1;
        Py_ssize_t sz = 0;
        const char *p = PyUnicode_AsUTF8AndSize(seq, &sz);
        if (!p) { PyErr_Clear(); return 1; }
        if (sz != PyUnicode_GET_LENGTH(seq) || sz > 0x7FFFFF00) return 1; /* one byte per character: ASCII */
        int overflow = 0;
        const long long v = PyLong_AsLongLongAndOverflow(idx, &overflow);
        if (overflow || (v == -1 && PyErr_Occurred())) { PyErr_Clear(); return 1; }
        keys[i] = (int64_t)v;
        ptrs[i] = p;
        lens[i] = (uint32_t)sz;
    }
    return 0;
}
