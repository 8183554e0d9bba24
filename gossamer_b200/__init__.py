"""gossamer_b200 -- B200-native `goss build-graph` / `build-kmer-set`.

Python is only a thin ctypes binding over the C ABI in include/gossamer_b200.h (the product is
libgossamer_b200.so + the C++ `goss` host).  There is no CPU path: importing works anywhere
(so that the symbol table can be checked), but every compute call needs a B200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgossamer_b200.so")

GRAPH, KMERSET = 0, 1
FASTA, FASTQ, LINE = 0, 1, 2
LAST_OF_FILE = 1
BLOCK_ASYNC = 2
ABI_VERSION = 1
NCCL_ID_BYTES = 128

STATUS_NAMES = {0: "GSB_OK", -1: "GSB_EINVAL", -2: "GSB_EPARSE", -3: "GSB_EIO", -4: "GSB_ENOMEM",
                -5: "GSB_ECUDA", -6: "GSB_ENCCL", -7: "GSB_ERANGE"}


class GossamerError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


class ParseError(GossamerError):
    pass


_LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("kind", C.c_int32), ("k", C.c_int32), ("device", C.c_int32),
                ("min_count", C.c_uint64), ("max_batch_keys", C.c_uint64), ("log", _LOG_FN), ("log_user", C.c_void_p)]


class Counts(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64), ("n_kept", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_h2d", "ms_scan", "ms_extract", "ms_sort", "ms_reduce", "ms_merge", "ms_emit",
                                          "ms_d2h", "ms_exchange", "ms_sort_sweeps", "ms_all_to_all", "ms_unfold")] + \
               [(n, C.c_uint64) for n in ("exchange_bytes_sent", "exchange_peer_memory", "bytes_in", "bytes_out", "n_symbols", "sort_key_bytes", "sort_passes",
                                          "sort_passes_model", "n_batches", "kernel_launches", "hbm_peak_bytes", "n_sorted_keys", "device_allocs")] + \
               [(n, C.c_double) for n in ("ms_exchange_agree", "ms_exchange_survivors", "ms_exchange_publish")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_OPEN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p))
_PWRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64)
_CLOSE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


class Sink(C.Structure):
    _fields_ = [("user", C.c_void_p), ("open", _OPEN_FN), ("pwrite", _PWRITE_FN), ("close", _CLOSE_FN)]


_SIZE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64))
_PREAD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64)


class Source(C.Structure):
    _fields_ = [("user", C.c_void_p), ("size", _SIZE_FN), ("pread", _PREAD_FN)]


class GraphInfo(C.Structure):
    _fields_ = [("version", C.c_uint64), ("k", C.c_uint64), ("flags", C.c_uint64), ("n_items", C.c_uint64)]


EXPORTS = ["gsb_kmerset_merge_annotate", "gsb_kmerset_near_kmers", "gsb_graph_peek", "gsb_graph_load", "gsb_graph_load_pairs", "gsb_graph_finish", "gsb_graph_dump", "gsb_create", "gsb_destroy", "gsb_last_error", "gsb_push_block", "gsb_push_device_block", "gsb_finish_counting",
           "gsb_emit", "gsb_timer_begin", "gsb_timer_end", "gsb_host_alloc", "gsb_host_free", "gsb_host_bind_near_device", "gsb_get_stats", "gsb_reset", "gsb_comm_make_id", "gsb_comm_attach", "gsb_gather_to_root", "gsb_plan_splitters", "gsb_samples_per_rank",
           "gsb_debug_copy_counts", "gsb_debug_sort_keys", "gsb_debug_sort_bench", "gsb_debug_set_tuning", "gsb_debug_set_partition", "gsb_debug_set_pairsort", "gsb_debug_plan", "gsb_debug_emit_sparse_array", "gsb_debug_emit_graph",
           "gsb_debug_extract"]

_lib = None


def lib():
    """Load libgossamer_b200.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `make -C gossamer_b200` (or __graft_entry__.build()); "
                              "gossamer_b200 has no CPU or pure-Python fallback")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.gsb_last_error.restype = C.c_char_p
        L.gsb_last_error.argtypes = [C.c_void_p]
        L.gsb_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.gsb_destroy.argtypes = [C.c_void_p]
        L.gsb_push_block.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32]
        L.gsb_push_device_block.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32]
        L.gsb_finish_counting.argtypes = [C.c_void_p, C.POINTER(Counts)]
        L.gsb_emit.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Sink)]
        L.gsb_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.gsb_timer_begin.argtypes = [C.c_void_p]
        L.gsb_timer_end.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.gsb_reset.argtypes = [C.c_void_p]
        L.gsb_comm_make_id.argtypes = [C.c_void_p]
        L.gsb_comm_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.gsb_gather_to_root.argtypes = [C.c_void_p]
        L.gsb_graph_peek.argtypes = [C.c_char_p, C.POINTER(Source), C.c_int, C.POINTER(GraphInfo), C.c_char_p, C.c_size_t]
        L.gsb_graph_load.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Source)]
        L.gsb_graph_load_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.gsb_graph_finish.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(Counts)]
        L.gsb_graph_dump.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Sink)]
        L.gsb_kmerset_merge_annotate.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(Source), C.c_char_p, C.POINTER(Sink), C.POINTER(C.c_uint64)]
        L.gsb_kmerset_near_kmers.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Source), C.POINTER(Sink), C.POINTER(C.c_uint64)]
        L.gsb_debug_copy_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.gsb_debug_copy_counts.restype = C.c_int64
        L.gsb_debug_sort_keys.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.gsb_debug_sort_keys.restype = C.c_int64
        L.gsb_debug_emit_sparse_array.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                                  C.c_char_p, C.POINTER(Sink)]
        L.gsb_debug_emit_graph.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.POINTER(Sink)]
        L.gsb_debug_extract.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                        C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
        L.gsb_debug_extract.restype = C.c_int64
        _lib = L
    return _lib


class MemorySink:
    """In-memory file store behind the gsb_sink callbacks (the StringFileFactory analogue)."""

    def __init__(self, keep_data=True):
        self.files = {}
        self.sizes = {}
        self.segments = {}          # name -> [(offset, length)]: the pieces this process wrote (multi-GPU emission)
        self.keep_data = keep_data
        self._handles = {}
        self._next = 1

        def _open(user, name, size_hint, out):
            # size_hint is the size of the whole file; a file may be opened again for a further piece, and
            # bytes that nobody writes are zero
            h = self._next
            self._next += 1
            nm = name.decode()
            self._handles[h] = nm
            self.sizes[nm] = max(self.sizes.get(nm, 0), size_hint)
            self.segments.setdefault(nm, [])
            if self.keep_data:
                buf = self.files.setdefault(nm, bytearray())
                if len(buf) < size_hint:
                    buf.extend(b"\0" * (size_hint - len(buf)))
            out[0] = h
            return 0

        def _pwrite(user, handle, offset, data, length):
            nm = self._handles[handle]
            self.sizes[nm] = max(self.sizes[nm], offset + length)
            self.segments[nm].append((offset, length))
            if self.keep_data:
                buf = self.files[nm]
                if len(buf) < offset + length:
                    buf.extend(b"\0" * (offset + length - len(buf)))
                buf[offset:offset + length] = C.string_at(data, length)
            return 0

        def _close(user, handle):
            self._handles.pop(handle, None)
            return 0

        self._cbs = (_OPEN_FN(_open), _PWRITE_FN(_pwrite), _CLOSE_FN(_close))
        self.c = Sink(None, *self._cbs)

    def as_bytes(self):
        return {k: bytes(v) for k, v in self.files.items()}


class MemorySource:
    """{name: bytes} behind the gsb_source callbacks (the StringFileFactory analogue for reading)."""

    def __init__(self, files):
        self.files = {k: bytes(v) for k, v in files.items()}

        def _size(user, name, out):
            b = self.files.get(name.decode())
            if b is None:
                return 1
            out[0] = len(b)
            return 0

        def _pread(user, name, offset, dst, length):
            b = self.files.get(name.decode())
            if b is None or offset + length > len(b):
                return 1
            C.memmove(dst, b[offset:offset + length], length)
            return 0

        self._cbs = (_SIZE_FN(_size), _PREAD_FN(_pread))
        self.c = Source(None, *self._cbs)


class DirectorySource:
    """Real files (names are paths)."""

    def __init__(self):
        def _size(user, name, out):
            try:
                out[0] = os.path.getsize(name.decode())
                return 0
            except OSError:
                return 1

        def _pread(user, name, offset, dst, length):
            try:
                with open(name.decode(), "rb") as f:
                    f.seek(offset)
                    b = f.read(length)
                if len(b) != length:
                    return 1
                C.memmove(dst, b, length)
                return 0
            except OSError:
                return 1

        self._cbs = (_SIZE_FN(_size), _PREAD_FN(_pread))
        self.c = Source(None, *self._cbs)


class DirectorySink:
    """Writes the file set under a directory prefix, like PhysicalFileFactory."""

    def __init__(self):
        self._handles = {}
        self._next = 1

        def _open(user, name, size_hint, out):
            # no truncation: with several GPUs every rank opens the same file and writes its own pieces
            h = self._next
            self._next += 1
            fd = os.open(name.decode(), os.O_CREAT | os.O_RDWR, 0o644)
            os.ftruncate(fd, size_hint)
            self._handles[h] = os.fdopen(fd, "r+b")
            out[0] = h
            return 0

        def _pwrite(user, handle, offset, data, length):
            f = self._handles[handle]
            f.seek(offset)
            f.write(C.string_at(data, length))
            return 0

        def _close(user, handle):
            self._handles.pop(handle).close()
            return 0

        self._cbs = (_OPEN_FN(_open), _PWRITE_FN(_pwrite), _CLOSE_FN(_close))
        self.c = Sink(None, *self._cbs)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Builder:
    """One counting context: push raw text blocks, finish, emit.  Mirrors GossCmdBuildGraph /
    GossCmdBuildKmerSet (src/GossCmdBuildGraph.cc:270-426, src/GossCmdBuildKmerSet.tcc:212-332)."""

    def __init__(self, kind, k, min_count=1, device=0, max_batch_keys=0, log=None):
        self._log_cb = _LOG_FN(lambda u, sev, msg: log(sev, msg.decode())) if log else _LOG_FN()
        cfg = Config(ABI_VERSION, kind, k, device, min_count, max_batch_keys, self._log_cb, None)
        self.h = C.c_void_p()
        rc = lib().gsb_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise GossamerError(rc, lib().gsb_last_error(None).decode())
        self.kind, self.k = kind, k

    def close(self):
        if self.h:
            lib().gsb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = lib().gsb_last_error(self.h).decode()
            raise (ParseError if rc == -2 else GossamerError)(rc, msg)

    def push(self, data, fmt, last=True):
        """data: bytes / bytearray / numpy uint8 array in host memory."""
        if isinstance(data, np.ndarray):
            buf = np.ascontiguousarray(data, dtype=np.uint8)
            self._check(lib().gsb_push_block(self.h, buf.ctypes.data, buf.size, fmt, LAST_OF_FILE if last else 0))
        else:
            b = bytes(data)
            self._check(lib().gsb_push_block(self.h, b, len(b), fmt, LAST_OF_FILE if last else 0))

    def push_pointer(self, host_ptr, nbytes, fmt, last=True, overlap=False):
        """overlap=True (GSB_BLOCK_ASYNC): the copy of this block overlaps the device work of the previous one; the
        memory must be page-locked and stay unchanged until the next push / finish returns."""
        flags = (LAST_OF_FILE if last else 0) | (BLOCK_ASYNC if overlap else 0)
        self._check(lib().gsb_push_block(self.h, C.c_void_p(host_ptr), nbytes, fmt, flags))

    def push_device(self, device_ptr, nbytes, fmt, last=True):
        self._check(lib().gsb_push_device_block(self.h, C.c_void_p(device_ptr), nbytes, fmt, LAST_OF_FILE if last else 0))

    def finish(self):
        c = Counts()
        self._check(lib().gsb_finish_counting(self.h, C.byref(c)))
        return c

    def emit(self, prefix, sink):
        """sink=None builds the files on the device and drops them (device-only timing)."""
        self._check(lib().gsb_emit(self.h, prefix.encode(), C.byref(sink.c) if sink is not None else None))

    def timer_begin(self):
        self._check(lib().gsb_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_double()
        self._check(lib().gsb_timer_end(self.h, C.byref(ms)))
        return ms.value

    def stats(self):
        s = Stats()
        self._check(lib().gsb_get_stats(self.h, C.byref(s)))
        return s

    def reset(self):
        self._check(lib().gsb_reset(self.h))

    # ---- existing file sets: trim / merge / dump / restore ----
    def load(self, prefix, source):
        """Decode the file set `prefix` on the device and merge it into the run (counts of equal keys are summed)."""
        self._check(lib().gsb_graph_load(self.h, prefix.encode(), C.byref(source.c)))

    def load_pairs(self, lo, hi, counts):
        lo = np.ascontiguousarray(lo, np.uint64)
        hi = None if hi is None else np.ascontiguousarray(hi, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint64)
        self._check(lib().gsb_graph_load_pairs(self.h, _ptr(lo), _ptr(hi), _ptr(counts), lo.size))

    def finish_loaded(self, cutoff=0, m_est=0):
        c = Counts()
        self._check(lib().gsb_graph_finish(self.h, cutoff, m_est, C.byref(c)))
        return c

    def dump(self, name, sink):
        self._check(lib().gsb_graph_dump(self.h, name.encode(), C.byref(sink.c)))

    def attach(self, nccl_id, n_ranks, rank):
        self._check(lib().gsb_comm_attach(self.h, nccl_id, n_ranks, rank))

    def gather_to_root(self):
        self._check(lib().gsb_gather_to_root(self.h))

    def counts_arrays(self):
        """(lo, hi, counts) of this rank's reduced run (test hook)."""
        m = lib().gsb_debug_copy_counts(self.h, None, None, None, 0)
        if m < 0:
            self._check(int(m))
        lo, hi, cn = np.zeros(m, np.uint64), np.zeros(m, np.uint64), np.zeros(m, np.uint64)
        lib().gsb_debug_copy_counts(self.h, _ptr(lo), _ptr(hi), _ptr(cn), m)
        return lo, hi, cn


def samples_per_rank():
    f = lib().gsb_samples_per_rank
    f.restype = C.c_uint32
    return int(f())


def plan_splitters(sample_lo, sample_hi, n_ranks):
    """Host-only: pooled samples -> (lo, hi) arrays of the n_ranks-1 splitters (no GPU needed)."""
    lo = np.ascontiguousarray(sample_lo, np.uint64)
    hi = np.ascontiguousarray(sample_hi, np.uint64)
    inter = np.empty(2 * lo.size, np.uint64)
    inter[0::2], inter[1::2] = lo, hi
    out = np.zeros(2 * (n_ranks - 1), np.uint64)
    f = lib().gsb_plan_splitters
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    rc = f(_ptr(inter), lo.size, n_ranks, _ptr(out))
    if rc != 0:
        raise GossamerError(rc, lib().gsb_last_error(None).decode())
    return out[0::2].copy(), out[1::2].copy()


def make_nccl_id():
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    rc = lib().gsb_comm_make_id(buf)
    if rc != 0:
        raise GossamerError(rc, lib().gsb_last_error(None).decode())
    return buf.raw


def graph_peek(prefix, source, kind=GRAPH):
    info, err = GraphInfo(), C.create_string_buffer(512)
    rc = lib().gsb_graph_peek(prefix.encode(), C.byref(source.c), kind, C.byref(info), err, 512)
    if rc != 0:
        raise GossamerError(rc, err.value.decode())
    return info


def trim_graph(files, src_prefix, dst_prefix, cutoff, device=0):
    """`goss trim-graph -C cutoff` on a file set held in memory (src/GossCmdTrimGraph.cc:27-127).  Returns {name: bytes}."""
    source = MemorySource(files)
    info = graph_peek(src_prefix, source)
    b = Builder(GRAPH, int(info.k), device=device)
    try:
        b.load(src_prefix, source)
        b.finish_loaded(cutoff=cutoff, m_est=0)            # the reference passes the exact kept count
        sink = MemorySink()
        b.emit(dst_prefix, sink)
        return sink.as_bytes()
    finally:
        b.close()


def merge_file_sets(files, prefixes, dst_prefix, kind=GRAPH, max_merge=8, device=0):
    """`goss merge-graphs` / `merge-kmer-sets` (src/GossCmdMerge.tcc:148-296), including its grouping into temporary
    file sets when there are more than max_merge inputs (the size estimate of the last merge depends on it)."""
    files = dict(files)
    todo = [(p, False) for p in prefixes]
    tmp_n = 0

    def merge(ins, out):
        source = MemorySource(files)
        infos = [graph_peek(p, source, kind) for p in ins]
        for p, i in zip(ins, infos):
            if i.k != infos[0].k:
                raise GossamerError(-1, "all graphs involved in a merge must have the same kmer-size.\n"
                                        f"{ins[0]} has k={infos[0].k}.\n{p} has k={i.k}.\n")
        tot = sum(int(i.n_items) for i in infos)
        b = Builder(kind, int(infos[0].k), device=device)
        try:
            for p in ins:
                b.load(p, source)
            b.finish_loaded(cutoff=0, m_est=tot)
            sink = MemorySink()
            b.emit(out, sink)
            return sink.as_bytes()
        finally:
            b.close()

    while len(todo) > max_merge:
        group, todo = todo[:max_merge], todo[max_merge:]
        out = f"__tmp{tmp_n}"
        tmp_n += 1
        files.update(merge([p for p, _ in group], out))
        todo.append((out, True))
    return merge([p for p, _ in todo], dst_prefix)


def merge_and_annotate_kmer_sets(files, lhs_prefix, rhs_prefix, dst_prefix, device=0):
    """`goss merge-and-annotate-kmer-sets` (xenome index, step 3; src/GossCmdMergeAndAnnotateKmerSets.cc:27-207) on file sets
    held in memory.  Returns ({name: bytes} of dst_prefix incl. .lhs-bits / .rhs-bits, (n_lhs, n_rhs, n_common, n_out))."""
    source = MemorySource(files)
    info = graph_peek(lhs_prefix, source, KMERSET)
    b = Builder(KMERSET, int(info.k), device=device)
    try:
        sink = MemorySink()
        st = (C.c_uint64 * 4)()
        b._check(lib().gsb_kmerset_merge_annotate(b.h, lhs_prefix.encode(), rhs_prefix.encode(), C.byref(source.c), dst_prefix.encode(), C.byref(sink.c), st))
        return sink.as_bytes(), tuple(int(x) for x in st)
    finally:
        b.close()


def compute_near_kmers(files, prefix, device=0):
    """`goss compute-near-kmers` (xenome index, step 4; src/GossCmdComputeNearKmers.cc:57-225).  Returns ({prefix.lhs-bits,
    prefix.rhs-bits: bytes}, number of gray k-mers)."""
    source = MemorySource(files)
    info = graph_peek(prefix, source, KMERSET)
    b = Builder(KMERSET, int(info.k), device=device)
    try:
        sink = MemorySink()
        gray = C.c_uint64()
        b._check(lib().gsb_kmerset_near_kmers(b.h, prefix.encode(), C.byref(source.c), C.byref(sink.c), C.byref(gray)))
        return sink.as_bytes(), int(gray.value)
    finally:
        b.close()


def build_graph(inputs, k, min_count=1, prefix="graph", sink=None, device=0):
    """inputs: list of (bytes-like, format).  Returns (sink, Counts, Stats)."""
    sink = sink or MemorySink()
    b = Builder(GRAPH, k, min_count=min_count, device=device)
    try:
        order = {LINE: 0, FASTA: 1, FASTQ: 2}      # src/GossCmdBuildGraph.cc:284-300
        for data, fmt in sorted(inputs, key=lambda x: order[x[1]]):
            b.push(data, fmt)
        counts = b.finish()
        b.emit(prefix, sink)
        return sink, counts, b.stats()
    finally:
        b.close()


def build_kmer_set(inputs, k, prefix="kset", sink=None, device=0):
    sink = sink or MemorySink()
    b = Builder(KMERSET, k, device=device)
    try:
        order = {LINE: 0, FASTA: 1, FASTQ: 2}
        for data, fmt in sorted(inputs, key=lambda x: order[x[1]]):
            b.push(data, fmt)
        counts = b.finish()
        b.emit(prefix, sink)
        return sink, counts, b.stats()
    finally:
        b.close()


# ---- component-level debug entry points (tests) --------------------------------------------------

def debug_sort_keys(lo, hi, key_bits, device=0):
    lo = np.ascontiguousarray(lo, np.uint64).copy()
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64).copy()
    rc = lib().gsb_debug_sort_keys(device, _ptr(lo), _ptr(hi), lo.size, key_bits)
    if rc < 0:
        raise GossamerError(int(rc), lib().gsb_last_error(None).decode())
    return lo, hi, int(rc)


LEGACY_COUNTING = 1 << 16
SAMPLED_SURVIVORS = 1 << 17      # multi-GPU: re-partition the survivors with sampled splitters (the fallback of the pair-sort exchange)


def debug_set_tuning(tuning_id):
    """Test / profiling switches (see include/gossamer_b200.h); LEGACY_COUNTING selects the LSD-sort counting path."""
    lib().gsb_debug_set_tuning(int(tuning_id))


def debug_set_partition(max_slots=0, total_bits=0):
    """Test-only: force the bucket geometry of the partition counting (0 = default)."""
    lib().gsb_debug_set_partition(int(max_slots), int(total_bits))


def host_bind_near_device(device):
    """Bind this process to the CPUs of the NUMA node the GPU hangs off (call before allocating pinned buffers)."""
    lib().gsb_host_bind_near_device(int(device))


def debug_plan(what, key_bytes, key_bits, n, first_bits=0):
    """Host-only: (levels, total bits, slots or capacity, [bits per pass]) of the counting (what=0), a streamed build (1), the pair sort (2)."""
    out = (C.c_uint32 * 11)()
    f = lib().gsb_debug_plan
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_uint32)]
    rc = f(what, key_bytes, key_bits, n, first_bits, out)
    if rc != 0:
        raise GossamerError(rc, "gsb_debug_plan")
    return int(out[0]), int(out[1]), int(out[2]), [int(x) for x in out[3:3 + int(out[0])]]


def debug_set_pairsort(cap=0, bits=-1):
    """Test-only: force the bucket capacity / partition bits of the pair sort (0 / -1 = default)."""
    lib().gsb_debug_set_pairsort(int(cap), int(bits))


def debug_sort_bench(n, key_bits, iters=3, tuning=0, device=0):
    """Random keys generated and sorted on the device -> (mean sweep ms, mean sort ms, sweeps per sort)."""
    sw, tot, nsw = C.c_double(), C.c_double(), C.c_int()
    f = lib().gsb_debug_sort_bench
    f.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    rc = f(device, n, key_bits, iters, tuning, C.byref(sw), C.byref(tot), C.byref(nsw))
    if rc != 0:
        raise GossamerError(rc, lib().gsb_last_error(None).decode())
    return sw.value, tot.value, nsw.value


def debug_emit_sparse_array(lo, hi, universe, m_est, base="sa", device=0):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = None if hi is None else np.ascontiguousarray(hi, np.uint64)
    sink = MemorySink()
    rc = lib().gsb_debug_emit_sparse_array(device, _ptr(lo), _ptr(hi), lo.size, universe & (2**64 - 1), universe >> 64, m_est,
                                           base.encode(), C.byref(sink.c))
    if rc != 0:
        raise GossamerError(rc, lib().gsb_last_error(None).decode())
    return sink.as_bytes()


def debug_emit_graph(lo, hi, counts, k, prefix="graph", device=0):
    lo = np.ascontiguousarray(lo, np.uint64)
    hi = np.ascontiguousarray(hi if hi is not None else np.zeros_like(lo), np.uint64)
    counts = np.ascontiguousarray(counts, np.uint64)
    sink = MemorySink()
    rc = lib().gsb_debug_emit_graph(device, _ptr(lo), _ptr(hi), _ptr(counts), lo.size, k, prefix.encode(), C.byref(sink.c))
    if rc != 0:
        raise GossamerError(rc, lib().gsb_last_error(None).decode())
    return sink.as_bytes()


def debug_extract(text, fmt, kind, k, device=0):
    text = bytes(text)
    err = C.create_string_buffer(512)
    nr = C.c_uint64()
    n = lib().gsb_debug_extract(device, text, len(text), fmt, kind, k, None, None, 0, C.byref(nr), err, 512)
    if n < 0:
        raise (ParseError if n == -2 else GossamerError)(int(n), err.value.decode())
    lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    n2 = lib().gsb_debug_extract(device, text, len(text), fmt, kind, k, _ptr(lo), _ptr(hi), n, C.byref(nr), err, 512)
    assert n2 == n
    return lo, hi, nr.value
