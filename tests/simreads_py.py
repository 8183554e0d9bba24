"""ctypes front-end to tools/libsimreads.so (seeded synthetic reads)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "tools", "libsimreads.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools"), "libsimreads.so"])
        L = C.CDLL(path)
        L.sim_fastq_bytes.restype = C.c_uint64
        L.sim_fastq_bytes.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.sim_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
        L.sim_reads_fastq.restype = C.c_uint64
        L.sim_reads_fastq.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def genome(G, seed=42):
    g = np.empty(G, np.uint8)
    lib().sim_genome(G, seed, g.ctypes.data)
    return g


def reads_fastq(genome_arr, read_len, n_reads, err=0.0, seed=43, first_idx=0, out=None):
    """FASTQ text as a numpy uint8 array (optionally written into `out`, e.g. a pinned buffer)."""
    n = lib().sim_fastq_bytes(n_reads, read_len, first_idx)
    if out is None:
        out = np.empty(n, np.uint8)
    assert out.size >= n
    w = lib().sim_reads_fastq(genome_arr.ctypes.data, genome_arr.size, read_len, n_reads, err, seed, first_idx, out.ctypes.data)
    assert w == n
    return out[:n]


def fastq_bytes(n_reads, read_len, first_idx=0):
    return lib().sim_fastq_bytes(n_reads, read_len, first_idx)
