"""ctypes front-end to oracle/liboracle.so (the CPU restatement of the reference).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (gossamer_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
FASTA, FASTQ, LINE = 0, 1, 2
MODE_GRAPH, MODE_KMERSET, MODE_FORWARD = 0, 1, 2


class OracleError(RuntimeError):
    pass


class OracleParseError(OracleError):
    pass


class _Input(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_uint64), ("format", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64),
                ("n_kept", C.c_uint64), ("t_extract", C.c_double), ("t_sort", C.c_double), ("t_emit", C.c_double)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_fs_new.restype = C.c_void_p
        L.orc_fs_free.argtypes = [C.c_void_p]
        L.orc_fs_count.argtypes = [C.c_void_p]
        L.orc_fs_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_fs_name.restype = C.c_char_p
        L.orc_fs_size.argtypes = [C.c_void_p, C.c_int]
        L.orc_fs_size.restype = C.c_uint64
        L.orc_fs_data.argtypes = [C.c_void_p, C.c_int]
        L.orc_fs_data.restype = C.c_void_p
        L.orc_fs_put.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64]
        for f in ("orc_extract", "orc_frame", "orc_count", "orc_read_graph", "orc_read_kmer_set"):
            getattr(L, f).restype = C.c_int64
        for f in ("orc_fnv_hash", "orc_sparse_d"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_uint64] * (2 if f == "orc_fnv_hash" else 3)
        _lib = L
    return _lib


def _u64p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _inputs(inputs):
    """inputs: list of (bytes-like, format).  Returns (ctypes array, keepalive)."""
    arr = (_Input * len(inputs))()
    keep = []
    for i, (data, fmt) in enumerate(inputs):
        if isinstance(data, np.ndarray):
            buf = np.ascontiguousarray(data, dtype=np.uint8)
            keep.append(buf)
            arr[i].data = buf.ctypes.data
            arr[i].size = buf.size
        else:
            b = bytes(data)
            cb = C.create_string_buffer(b, len(b)) if len(b) else C.create_string_buffer(1)
            keep.append(cb)
            arr[i].data = C.cast(cb, C.c_void_p).value
            arr[i].size = len(b)
        arr[i].format = fmt
    return arr, keep


def _check(rc, err):
    if rc == -2:
        raise OracleParseError(err.value.decode())
    if rc < 0:
        raise OracleError(err.value.decode())
    return rc


class MemFS:
    """In-memory file set (name -> bytes) living on the C++ side."""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_fs_new())

    def __del__(self):
        try:
            lib().orc_fs_free(self.h)
        except Exception:
            pass

    def files(self):
        L = lib()
        out = {}
        for i in range(L.orc_fs_count(self.h)):
            n = L.orc_fs_size(self.h, i)
            out[L.orc_fs_name(self.h, i).decode()] = C.string_at(L.orc_fs_data(self.h, i), n) if n else b""
        return out

    def put(self, name, data):
        data = bytes(data)
        lib().orc_fs_put(self.h, name.encode(), data, len(data))


def build_graph(inputs, k, min_count=1, threads=1, base="graph"):
    arr, keep = _inputs(inputs)
    fs, st, err = MemFS(), Stats(), C.create_string_buffer(512)
    _check(lib().orc_build_graph(arr, len(inputs), k, C.c_uint64(min_count), threads, base.encode(), fs.h, C.byref(st), err, 512), err)
    return fs, st


def build_kmer_set(inputs, k, threads=1, base="kset"):
    arr, keep = _inputs(inputs)
    fs, st, err = MemFS(), Stats(), C.create_string_buffer(512)
    _check(lib().orc_build_kmer_set(arr, len(inputs), k, threads, base.encode(), fs.h, C.byref(st), err, 512), err)
    return fs, st


def extract(inputs, w, mode):
    """Window keys in stream order -> (lo u64[], hi u64[], n_reads)."""
    arr, keep = _inputs(inputs)
    err = C.create_string_buffer(512)
    nr = C.c_uint64()
    n = _check(lib().orc_extract(arr, len(inputs), w, mode, None, None, C.c_uint64(0), C.byref(nr), err, 512), err)
    lo = np.zeros(n, np.uint64)
    hi = np.zeros(n, np.uint64)
    _check(lib().orc_extract(arr, len(inputs), w, mode, _u64p(lo), _u64p(hi), C.c_uint64(n), C.byref(nr), err, 512), err)
    return lo, hi, nr.value


def frame(inputs):
    arr, keep = _inputs(inputs)
    err = C.create_string_buffer(512)
    nr = C.c_uint64()
    n = _check(lib().orc_frame(arr, len(inputs), None, C.c_uint64(0), C.byref(nr), err, 512), err)
    buf = C.create_string_buffer(max(n, 1))
    _check(lib().orc_frame(arr, len(inputs), buf, C.c_uint64(n), C.byref(nr), err, 512), err)
    reads = buf.raw[:n].decode("latin-1").split("\n")[:-1] if n else []
    return reads


def count(lo, hi, key_bits, min_count=1, threads=1):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = None if hi is None else np.ascontiguousarray(hi, np.uint64)
    n = lo.size
    olo, ohi, oc = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    m = lib().orc_count(_u64p(lo), _u64p(hi), C.c_uint64(n), key_bits, C.c_uint64(min_count), threads, _u64p(olo), _u64p(ohi), _u64p(oc))
    return olo[:m].copy(), ohi[:m].copy(), oc[:m].copy()


def write_graph(lo, hi, counts, k, m_est=None, base="graph"):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    counts = np.ascontiguousarray(counts, np.uint64)
    fs, err = MemFS(), C.create_string_buffer(512)
    m_est = lo.size if m_est is None else m_est
    _check(lib().orc_write_graph(_u64p(lo), _u64p(hi), _u64p(counts), C.c_uint64(lo.size), k, C.c_uint64(m_est), base.encode(), fs.h, err, 512), err)
    return fs


def write_kmer_set(lo, hi, k, m_est=None, base="kset"):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    fs, err = MemFS(), C.create_string_buffer(512)
    m_est = lo.size if m_est is None else m_est
    _check(lib().orc_write_kmer_set(_u64p(lo), _u64p(hi), C.c_uint64(lo.size), k, C.c_uint64(m_est), base.encode(), fs.h, err, 512), err)
    return fs


def write_sparse_array(lo, hi, n_universe, m_est, base="sa"):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    fs, err = MemFS(), C.create_string_buffer(512)
    _check(lib().orc_write_sparse_array(_u64p(lo), _u64p(hi), C.c_uint64(lo.size), C.c_uint64(n_universe & (2**64 - 1)),
                                        C.c_uint64(n_universe >> 64), C.c_uint64(m_est), base.encode(), fs.h, err, 512), err)
    return fs


def write_dense_select(pos, invert, name="ds"):
    pos = np.ascontiguousarray(pos, np.uint64)
    fs, err = MemFS(), C.create_string_buffer(512)
    _check(lib().orc_write_dense_select(_u64p(pos), C.c_uint64(pos.size), int(invert), name.encode(), fs.h, err, 512), err)
    return fs


def write_vba(counts, m_est=None, base="vba"):
    counts = np.ascontiguousarray(counts, np.uint32)
    fs, err = MemFS(), C.create_string_buffer(512)
    m_est = counts.size if m_est is None else m_est
    _check(lib().orc_write_vba(_u64p(counts), C.c_uint64(counts.size), C.c_uint64(m_est), base.encode(), fs.h, err, 512), err)
    return fs


def _as_fs(files):
    if isinstance(files, MemFS):
        return files
    fs = MemFS()
    for name, data in files.items():
        fs.put(name, data)
    return fs


def read_graph(files, base="graph", exercise_select=True):
    """Decode a Graph through the restated reference readers -> (k, lo, hi, counts u32, hist_total)."""
    fs = _as_fs(files)
    err = C.create_string_buffer(512)
    k, ht = C.c_uint64(), C.c_uint64()
    n = _check(lib().orc_read_graph(fs.h, base.encode(), None, None, None, C.c_uint64(0), C.byref(k), C.byref(ht), 0, err, 512), err)
    lo, hi, cn = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    _check(lib().orc_read_graph(fs.h, base.encode(), _u64p(lo), _u64p(hi), _u64p(cn), C.c_uint64(n), C.byref(k), C.byref(ht),
                                int(exercise_select), err, 512), err)
    return k.value, lo, hi, cn, ht.value


def read_kmer_set(files, base="kset", exercise_select=True):
    fs = _as_fs(files)
    err = C.create_string_buffer(512)
    k, cnt = C.c_uint64(), C.c_uint64()
    n = _check(lib().orc_read_kmer_set(fs.h, base.encode(), None, None, C.c_uint64(0), C.byref(k), C.byref(cnt), 0, err, 512), err)
    lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    _check(lib().orc_read_kmer_set(fs.h, base.encode(), _u64p(lo), _u64p(hi), C.c_uint64(n), C.byref(k), C.byref(cnt),
                                   int(exercise_select), err, 512), err)
    return k.value, cnt.value, lo, hi


def dense_select_eval(files, bitmap_name, ds_name, invert, n):
    fs = _as_fs(files)
    err = C.create_string_buffer(512)
    out = np.zeros(n, np.uint64)
    _check(lib().orc_dense_select_eval(fs.h, bitmap_name.encode(), ds_name.encode(), int(invert), C.c_uint64(n), _u64p(out), err, 512), err)
    return out


def reverse_complement(x, k):
    lo, hi = C.c_uint64(), C.c_uint64()
    lib().orc_reverse_complement(C.c_uint64(x & (2**64 - 1)), C.c_uint64(x >> 64), k, C.byref(lo), C.byref(hi))
    return lo.value | (hi.value << 64)


def normalize(x, k):
    lo, hi = C.c_uint64(), C.c_uint64()
    lib().orc_normalize(C.c_uint64(x & (2**64 - 1)), C.c_uint64(x >> 64), k, C.byref(lo), C.byref(hi))
    return lo.value | (hi.value << 64)


def fnv_hash(x):
    return lib().orc_fnv_hash(C.c_uint64(x & (2**64 - 1)), C.c_uint64(x >> 64))


def sparse_d(n_universe, m):
    return lib().orc_sparse_d(C.c_uint64(n_universe & (2**64 - 1)), C.c_uint64(n_universe >> 64), C.c_uint64(m))


def kmer_to_string(x, k):
    buf = C.create_string_buffer(k + 1)
    lib().orc_kmer_to_string(C.c_uint64(x & (2**64 - 1)), C.c_uint64(x >> 64), k, buf)
    return buf.value.decode()


def merge_and_annotate(files, lhs, rhs, out):
    """Restated merge-and-annotate-kmer-sets: -> (files of `out` incl. .lhs-bits/.rhs-bits, (n_lhs, n_rhs, n_common, n_out))."""
    fs = _as_fs(files)
    err = C.create_string_buffer(512)
    st = (C.c_uint64 * 4)()
    _check(lib().orc_merge_and_annotate(fs.h, lhs.encode(), rhs.encode(), out.encode(), st, err, 512), err)
    return {n: v for n, v in fs.files().items() if n.startswith(out + ".")}, tuple(int(x) for x in st)


def compute_near_kmers(files, base):
    """Restated compute-near-kmers: -> ({base.lhs-bits, base.rhs-bits}, number of gray k-mers)."""
    fs = _as_fs(files)
    err = C.create_string_buffer(512)
    lib().orc_compute_near_kmers.restype = C.c_int64
    gray = _check(lib().orc_compute_near_kmers(fs.h, base.encode(), err, 512), err)
    out = fs.files()
    return {n: out[n] for n in (base + ".lhs-bits", base + ".rhs-bits")}, int(gray)
