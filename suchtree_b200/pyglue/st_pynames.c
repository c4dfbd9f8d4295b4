/* Name -> id glue for the by-name entry points (SuchTree.distances_by_name,
 * MuchTree.pyx:945-979; quartet_topologies_by_name, :1378-1422).
 *
 * The reference walks the list of name tuples in Python: two isinstance checks, two `in`
 * tests and two dict lookups per pair.  With the distances themselves taking ~1 ms per 10^6
 * pairs on the GPU, that walk IS the call; this is the same walk as one C loop over the
 * list (borrowed references, the dict's cached string hashes), writing int64 ids straight
 * into the array the CUDA path reads.
 *
 * Loaded with ctypes.PyDLL (the GIL stays held; arguments are py_object).  Returns -1 when
 * every row was converted, else the index of the first row it could not convert (not a
 * tuple/list of `width` str, or a name that is not a leaf): the caller then runs the
 * reference's own loop, which raises the reference's own error for that row.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ST_PY_API __attribute__((visibility("default")))
#else
#define ST_PY_API
#endif

ST_PY_API Py_ssize_t st_py_names_to_ids(PyObject *rows, PyObject *leaves, int64_t *out, Py_ssize_t width) {
    if (!PyList_CheckExact(rows) || !PyDict_CheckExact(leaves) || !out || width < 1) return 0;
    const Py_ssize_t n = PyList_GET_SIZE(rows);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *row = PyList_GET_ITEM(rows, i);
        PyObject **items;
        if (PyTuple_CheckExact(row) && PyTuple_GET_SIZE(row) == width) {
            items = &PyTuple_GET_ITEM(row, 0);
        } else if (PyList_CheckExact(row) && PyList_GET_SIZE(row) == width) {
            items = &PyList_GET_ITEM(row, 0);
        } else {
            return i;
        }
        for (Py_ssize_t k = 0; k < width; ++k) {
            PyObject *name = items[k];
            if (!PyUnicode_CheckExact(name)) return i;
            PyObject *id = PyDict_GetItemWithError(leaves, name); /* borrowed */
            if (!id) {
                PyErr_Clear();
                return i;
            }
            const long long v = PyLong_AsLongLong(id);
            if (v == -1 && PyErr_Occurred()) {
                PyErr_Clear();
                return i;
            }
            out[i * width + k] = (int64_t)v;
        }
    }
    return -1;
}

ST_PY_API int st_py_glue_version(void) { return 1; }
