/* Host helper of the drop-in boundary (CPython extension, no CUDA): pack a list / tuple of equal-length str objects
 * into a contiguous uint8[N, L] buffer of residue CHARACTER codes in one pass.
 *
 * This is the host half of what replaces the reference's per-character Python loop (flexs/utils/sequence_utils.py:32-47,
 * called per sequence at keras_model.py:53-56 and :70-73): the characters go to the GPU as bytes and
 * flexs_encode_dev maps them to residue indices there.  The pure-Python route ("".join + encode + a length check)
 * costs ~0.22 s per million 100-mers; this pass copies each string's Latin-1 buffer directly (~0.03 s).
 *
 * pack(sequences, out) -> width
 *   sequences: list or tuple of str, all of one length (ValueError otherwise: "all sequences must have the same length"),
 *              only code points <= 255 (ValueError otherwise, worded like str.index's failure as the Python path is)
 *   out:       writable C-contiguous buffer of exactly len(sequences) * width bytes
 *
 * pack_bits(sequences, alphabet, out, threads) -> width
 *   the same walk, but each character is mapped through the alphabet (first occurrence wins, like str.index) and
 *   stored as ceil(log2 A) bits, rows padded to whole bytes: the wire format of flexs_model_score_host_packed
 *   (include/flexs_b200.h).  The per-row work runs on `threads` POSIX threads with the GIL released (the strings are
 *   kept alive by references taken first).  A character outside the alphabet raises ValueError naming the sequence
 *   and position, as str.index does in the reference (sequence_utils.py:46).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static PyObject *pack(PyObject *self, PyObject *args) {
    PyObject *seqs;
    Py_buffer out;
    (void)self;
    if (!PyArg_ParseTuple(args, "Ow*", &seqs, &out)) return NULL;
    PyObject *fast = PySequence_Fast(seqs, "sequences must be a list or tuple of str");
    if (!fast) { PyBuffer_Release(&out); return NULL; }
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject **items = PySequence_Fast_ITEMS(fast);
    Py_ssize_t width = -1;
    unsigned char *dst = (unsigned char *)out.buf;
    PyObject *result = NULL;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *s = items[i];
        if (!PyUnicode_Check(s)) { PyErr_SetString(PyExc_TypeError, "sequences must be str"); goto done; }
        const Py_ssize_t len = PyUnicode_GET_LENGTH(s);
        if (width < 0) {
            width = len;
            if (!PyBuffer_IsContiguous(&out, 'C') || out.len != n * width) {
                PyErr_SetString(PyExc_ValueError, "all sequences must have the same length");
                goto done;
            }
        } else if (len != width) {
            PyErr_SetString(PyExc_ValueError, "all sequences must have the same length");
            goto done;
        }
        if (PyUnicode_KIND(s) != PyUnicode_1BYTE_KIND) {  /* compact strings are 1-byte exactly when every code point <= 255 */
            PyErr_SetString(PyExc_ValueError, "substring not found: non latin-1 character in sequence");
            goto done;
        }
        memcpy(dst + i * width, PyUnicode_1BYTE_DATA(s), (size_t)width);
    }
    result = PyLong_FromSsize_t(width < 0 ? 0 : width);
done:
    Py_DECREF(fast);
    PyBuffer_Release(&out);
    return result;
}

typedef struct {
    PyObject **items;
    Py_ssize_t begin, end, width, row_bytes;
    int bits;
    const unsigned char *lut;
    unsigned char *dst;
    int err;                      /* 0 ok, 1 not a str, 2 ragged, 3 not latin-1, 4 character outside the alphabet */
    Py_ssize_t bad_row, bad_col;  /* first offending sequence (and position, for err 4) in this range */
} pack_job;

/* Runs WITHOUT the GIL: only reads immutable fields of str objects that the caller's list keeps alive. */
static void *pack_rows(void *arg) {
    pack_job *j = (pack_job *)arg;
    j->err = 0; j->bad_row = -1; j->bad_col = -1;
    for (Py_ssize_t r = j->begin; r < j->end; ++r) {
        PyObject *s = j->items[r];
        if (r + 8 < j->end) {   /* header + characters of a compact str: up to 40 + width bytes */
            const char *nx = (const char *)j->items[r + 8];
            __builtin_prefetch(nx);
            __builtin_prefetch(nx + 64);
            if (j->width > 88) __builtin_prefetch(nx + 128);
        }
        int err = 0;
        if (!PyUnicode_Check(s)) err = 1;
        else if (PyUnicode_GET_LENGTH(s) != j->width) err = 2;
        else if (PyUnicode_KIND(s) != PyUnicode_1BYTE_KIND) err = 3;
        if (err) {
            if (!j->err) { j->err = err; j->bad_row = r; }
            continue;
        }
        const unsigned char *src = PyUnicode_1BYTE_DATA(s);
        const unsigned char *lut = j->lut;
        unsigned char *out = j->dst + r * j->row_bytes;
        const Py_ssize_t width = j->width;
        unsigned bad = 0;   /* OR of the codes: a character outside the alphabet (0xFF) sets bit 7; looked at once per row */
        if (j->bits == 2) {
            /* DNA / RNA: four residues per output byte, no shifts by a variable, no branch per character */
            Py_ssize_t i = 0;
            for (; i + 4 <= width; i += 4) {
                const unsigned c0 = lut[src[i]], c1 = lut[src[i + 1]], c2 = lut[src[i + 2]], c3 = lut[src[i + 3]];
                bad |= c0 | c1 | c2 | c3;
                *out++ = (unsigned char)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
            }
            if (i < width) {
                unsigned byte = 0;
                for (int sh = 0; i < width; ++i, sh += 2) {
                    const unsigned c = lut[src[i]];
                    bad |= c;
                    byte |= (c & 3u) << sh;
                }
                *out++ = (unsigned char)byte;
            }
        } else {
            uint64_t acc = 0;
            int have = 0;
            const int bits = j->bits;
            for (Py_ssize_t i = 0; i < width; ++i) {
                const unsigned code = lut[src[i]];
                bad |= code;
                acc |= (uint64_t)(code & 0x7Fu) << have;
                have += bits;
                if (have >= 32) { memcpy(out, &acc, 4); out += 4; acc >>= 32; have -= 32; }
            }
            while (have > 0) { *out++ = (unsigned char)acc; acc >>= 8; have -= 8; }
        }
        if ((bad & 0x80u) && !j->err) {   /* the row's bytes are meaningless now; the caller raises ValueError */
            Py_ssize_t i = 0;
            while (i < width && lut[src[i]] != 0xFF) ++i;
            j->err = 4; j->bad_row = r; j->bad_col = i;
        }
    }
    return NULL;
}

static PyObject *pack_bits(PyObject *self, PyObject *args) {
    PyObject *seqs;
    Py_buffer out, alpha;
    int threads = 1;
    (void)self;
    if (!PyArg_ParseTuple(args, "Oy*w*|i", &seqs, &alpha, &out, &threads)) return NULL;
    PyObject *fast = PySequence_Fast(seqs, "sequences must be a list or tuple of str");
    if (!fast) { PyBuffer_Release(&out); PyBuffer_Release(&alpha); return NULL; }
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject **items = PySequence_Fast_ITEMS(fast);
    PyObject *result = NULL;
    unsigned char lut[256];
    memset(lut, 0xFF, sizeof lut);
    /* codes stay below 128: bit 7 of a looked-up code marks a character outside the alphabet (pack_rows ORs the codes) */
    if (alpha.len < 2 || alpha.len > 127) { PyErr_SetString(PyExc_ValueError, "alphabet must have 2..127 characters"); goto done; }
    for (Py_ssize_t i = alpha.len - 1; i >= 0; --i) lut[((const unsigned char *)alpha.buf)[i]] = (unsigned char)i;
    int bits = 1;
    while ((1 << bits) < alpha.len) ++bits;
    Py_ssize_t width = 0;
    if (n > 0) {
        if (!PyUnicode_Check(items[0])) { PyErr_SetString(PyExc_TypeError, "sequences must be str"); goto done; }
        width = PyUnicode_GET_LENGTH(items[0]);
    }
    const Py_ssize_t row_bytes = (width * bits + 7) / 8;
    if (!PyBuffer_IsContiguous(&out, 'C') || out.len != n * row_bytes) {
        PyErr_SetString(PyExc_ValueError, "out must be a C-contiguous buffer of len(sequences) * ceil(width * bits / 8) bytes");
        goto done;
    }
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n * width < (Py_ssize_t)1 << 18) threads = 1;  /* thread start-up costs more than the packing */
    pack_job jobs[64];
    pthread_t tids[64];
    int started[64];
    /* The str objects are reached through the caller's list, which must not be mutated while this call runs (the usual
     * contract of an extension that releases the GIL over borrowed data); touching a million heap objects is a million
     * cache misses, so that walk is what the threads parallelise. */
    Py_BEGIN_ALLOW_THREADS
    for (int t = 0; t < threads; ++t) {
        jobs[t].items = items; jobs[t].width = width; jobs[t].row_bytes = row_bytes; jobs[t].bits = bits;
        jobs[t].lut = lut; jobs[t].dst = (unsigned char *)out.buf;
        jobs[t].begin = n * t / threads; jobs[t].end = n * (t + 1) / threads;
        started[t] = (t < threads - 1) && pthread_create(&tids[t], NULL, pack_rows, &jobs[t]) == 0;
        if (!started[t]) pack_rows(&jobs[t]);  /* the last range (or one whose thread could not start) runs here */
    }
    for (int t = 0; t < threads; ++t)
        if (started[t]) pthread_join(tids[t], NULL);
    Py_END_ALLOW_THREADS
    for (int t = 0; t < threads; ++t) {
        if (!jobs[t].err) continue;
        const Py_ssize_t r = jobs[t].bad_row, c = jobs[t].bad_col;
        if (jobs[t].err == 1) PyErr_SetString(PyExc_TypeError, "sequences must be str");
        else if (jobs[t].err == 2) PyErr_SetString(PyExc_ValueError, "all sequences must have the same length");
        else if (jobs[t].err == 3) PyErr_SetString(PyExc_ValueError, "substring not found: non latin-1 character in sequence");
        else {
            PyObject *ch = PyUnicode_FromOrdinal(PyUnicode_1BYTE_DATA(items[r])[c]);
            PyErr_Format(PyExc_ValueError, "substring not found: character %R of sequence %zd (position %zd) is not in the alphabet",
                         ch ? ch : Py_None, r, c);
            Py_XDECREF(ch);
        }
        goto done;  /* ranges are ordered: the first range with an error holds the first offending sequence */
    }
    result = PyLong_FromSsize_t(width);
done:
    Py_DECREF(fast);
    PyBuffer_Release(&out);
    PyBuffer_Release(&alpha);
    return result;
}

static PyMethodDef methods[] = {
    {"pack_bits", pack_bits, METH_VARARGS,
     "pack_bits(sequences, alphabet, out, threads=1) -> width: map equal-length str objects through the alphabet and store "
     "ceil(log2 A) bits per residue"},
    {"pack", pack, METH_VARARGS, "pack(sequences, out) -> width: copy equal-length Latin-1 str objects into a uint8 buffer"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_packstr", "host string packing for flexs_b200", -1, methods,
                                       NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__packstr(void) { return PyModule_Create(&moduledef); }
