"""ctypes front-end to oracle/_ref/libgossref.so: the REAL reference writers/readers
(SparseArray / DenseSelect / WordyBitVector / IntegerArray / VariableByteArray / Graph / KmerSet),
compiled unmodified from /root/reference/src against the Boost shim in oracle/ref/shim.
Test infrastructure only.  `available()` is False when the library has not been built (it can only
be built where /root/reference exists; the prebuilt .so travels to the GPU box)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libgossref.so")
_lib = None


def available():
    if os.path.exists(PATH):
        return True
    if os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref")])
        except Exception:
            return False
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.ref_store_new.restype = C.c_void_p
        L.ref_store_free.argtypes = [C.c_void_p]
        L.ref_store_put.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64]
        L.ref_store_list.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int]
        L.ref_store_name.argtypes = [C.c_void_p, C.c_int]
        L.ref_store_name.restype = C.c_char_p
        L.ref_store_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_store_size.restype = C.c_uint64
        L.ref_store_data.argtypes = [C.c_void_p, C.c_int]
        L.ref_store_data.restype = C.c_void_p
        L.ref_read_graph.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


LOW_SUFFIXES = ["", ".upr", ".lwr", ".upr.upr", ".upr.lwr", ".lwr.upr", ".lwr.lwr"]


def sparse_array_names(base):
    return [base + ".header", base + ".high-bits", base + "-d0", base + "-d1"] + [base + ".low-bits" + s for s in LOW_SUFFIXES]


def graph_names(base):
    n = [base + ".header", base + "-counts-hist.txt", base + "-counts.ord0", base + "-counts.ord1", base + "-counts.ord2"]
    n += sparse_array_names(base + "-edges") + sparse_array_names(base + "-counts.ord1p") + sparse_array_names(base + "-counts.ord2p")
    return n


def kmer_set_names(base):
    return [base + ".header"] + sparse_array_names(base + ".kmers")


class Store:
    def __init__(self):
        self.h = C.c_void_p(lib().ref_store_new())

    def __del__(self):
        try:
            lib().ref_store_free(self.h)
        except Exception:
            pass

    def put_all(self, files):
        for name, data in files.items():
            data = bytes(data)
            lib().ref_store_put(self.h, name.encode(), data, len(data))

    def files(self, candidates):
        arr = (C.c_char_p * len(candidates))(*[c.encode() for c in candidates])
        n = lib().ref_store_list(self.h, arr, len(candidates))
        out = {}
        for i in range(n):
            sz = lib().ref_store_size(self.h, i)
            out[lib().ref_store_name(self.h, i).decode()] = C.string_at(lib().ref_store_data(self.h, i), sz) if sz else b""
        return out


def _check(rc, err):
    if rc < 0:
        raise RuntimeError("reference: " + err.value.decode())
    return rc


def write_graph(lo, hi, counts, k, m_est=None, base="graph"):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    counts = np.ascontiguousarray(counts, np.uint64)
    st, err = Store(), C.create_string_buffer(512)
    _check(lib().ref_write_graph(st.h, _p(lo), _p(hi), _p(counts), C.c_uint64(lo.size), k,
                                 C.c_uint64(lo.size if m_est is None else m_est), base.encode(), err, 512), err)
    return st.files(graph_names(base))


def write_kmer_set(lo, hi, k, m_est=None, base="kset"):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    st, err = Store(), C.create_string_buffer(512)
    _check(lib().ref_write_kmer_set(st.h, _p(lo), _p(hi), C.c_uint64(lo.size), k, C.c_uint64(lo.size if m_est is None else m_est),
                                    base.encode(), err, 512), err)
    return st.files(kmer_set_names(base))


def write_sparse_array(lo, hi, universe, m_est, base="sa"):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    st, err = Store(), C.create_string_buffer(512)
    _check(lib().ref_write_sparse_array(st.h, _p(lo), _p(hi), C.c_uint64(lo.size), C.c_uint64(universe & (2**64 - 1)),
                                        C.c_uint64(universe >> 64), C.c_uint64(m_est), base.encode(), err, 512), err)
    return st.files(sparse_array_names(base))


def _names(lst):
    arr = (C.c_char_p * max(1, len(lst)))(*[x.encode() for x in lst])
    return arr, len(lst)


def build_graph(inputs, k, threads=2, log_slots=20, base="graph"):
    """The reference's own GossCmdBuildGraph (parsers, k-merising, BackyardHash, sort, builders) over
    in-memory files.  inputs: list of (bytes, format) with format 0 fasta / 1 fastq / 2 line."""
    st, err = Store(), C.create_string_buffer(512)
    names = {0: [], 1: [], 2: []}
    for i, (data, fmt) in enumerate(inputs):
        nm = f"in{i}." + {0: "fa", 1: "fq", 2: "txt"}[fmt]
        st.put_all({nm: bytes(data)})
        names[fmt].append(nm)
    fa, nfa = _names(names[0])
    fq, nfq = _names(names[1])
    ln, nln = _names(names[2])
    _check(lib().ref_build_graph(st.h, k, C.c_uint64(log_slots), C.c_uint64(1 << log_slots), C.c_uint64(threads), base.encode(),
                                 fa, nfa, fq, nfq, ln, nln, err, 512), err)
    return st, st.files(graph_names(base))


def build_kmer_set(inputs, k, threads=2, log_slots=20, base="kset"):
    st, err = Store(), C.create_string_buffer(512)
    names = {0: [], 1: [], 2: []}
    for i, (data, fmt) in enumerate(inputs):
        nm = f"in{i}." + {0: "fa", 1: "fq", 2: "txt"}[fmt]
        st.put_all({nm: bytes(data)})
        names[fmt].append(nm)
    fa, nfa = _names(names[0])
    fq, nfq = _names(names[1])
    ln, nln = _names(names[2])
    _check(lib().ref_build_kmer_set(st.h, k, C.c_uint64(log_slots), C.c_uint64(1 << log_slots), C.c_uint64(threads), base.encode(),
                                    fa, nfa, fq, nfq, ln, nln, err, 512), err)
    return st, st.files(kmer_set_names(base))


def trim_graph(store, src, dst, c):
    """The reference's own `trim-graph -C c` on a graph already in `store`."""
    err = C.create_string_buffer(512)
    _check(lib().ref_trim_graph(store.h, src.encode(), dst.encode(), C.c_uint64(c), err, 512), err)
    return store.files(graph_names(dst))


def read_graph(files, base="graph"):
    """Open a file set with the reference's own Graph::open / select / rank / multiplicity."""
    st, err = Store(), C.create_string_buffer(512)
    st.put_all(files)
    k = C.c_uint64()
    n = _check(lib().ref_read_graph(st.h, base.encode(), None, None, None, C.c_uint64(0), C.byref(k), err, 512), err)
    lo, hi, cn = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    _check(lib().ref_read_graph(st.h, base.encode(), _p(lo), _p(hi), _p(cn), C.c_uint64(n), C.byref(k), err, 512), err)
    return k.value, lo, hi, cn


def merge_graphs(store, ins, out, max_merge=8, kmer_sets=False):
    """The reference's own merge-graphs / merge-kmer-sets on file sets already in `store`."""
    err = C.create_string_buffer(1024)
    arr, n = _names(list(ins))
    _check(lib().ref_merge_graphs(store.h, arr, n, C.c_uint64(max_merge), out.encode(), 1 if kmer_sets else 0, err, 1024), err)
    return store.files(kmer_set_names(out) if kmer_sets else graph_names(out))


def dump_graph(store, src, out_file="dump.txt"):
    err = C.create_string_buffer(512)
    _check(lib().ref_dump_graph(store.h, src.encode(), out_file.encode(), err, 512), err)
    return store.files([out_file])[out_file]


def restore_graph(store, text, out, in_file="restore.txt"):
    err = C.create_string_buffer(512)
    store.put_all({in_file: bytes(text)})
    _check(lib().ref_restore_graph(store.h, in_file.encode(), out.encode(), err, 512), err)
    return store.files(graph_names(out))


def annotated_names(base):
    return kmer_set_names(base) + [base + ".lhs-bits", base + ".rhs-bits"]


def merge_and_annotate(store, lhs, rhs, out):
    """The reference's own merge-and-annotate-kmer-sets (xenome index, step 3) on kmer sets already in `store`."""
    err = C.create_string_buffer(512)
    _check(lib().ref_merge_and_annotate(store.h, lhs.encode(), rhs.encode(), out.encode(), err, 512), err)
    return store.files(annotated_names(out))


def compute_near_kmers(store, base, threads=2):
    """The reference's own compute-near-kmers (xenome index, step 4): rewrites base.lhs-bits / base.rhs-bits in `store`."""
    err = C.create_string_buffer(512)
    _check(lib().ref_compute_near_kmers(store.h, base.encode(), C.c_uint64(threads), err, 512), err)
    return store.files([base + ".lhs-bits", base + ".rhs-bits"])
