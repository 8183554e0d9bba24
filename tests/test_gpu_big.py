"""Large configurations (BASELINE configs 2-4 at or near full size), run only with GSB_BIG=1 on the
GPU box because they take minutes and tens of GB.  Parity at these sizes goes through
size-independent properties: instance counts in closed form, single-batch == forced multi-batch
(byte-identical files), header / bitmap / plane consistency, symmetry of a sample."""
import ctypes as C
import hashlib
import os
import struct
import time

import numpy as np
import pytest

import gossamer_b200 as G
import oracle_py as O
import simreads_py as S

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("GSB_BIG") != "1", reason="set GSB_BIG=1 to run the large configurations")]


class HashSink:
    """gsb_sink that keeps sha256 + size per file (and the bytes of small files)."""

    def __init__(self, keep_below=1 << 20):
        self.hash, self.size, self.small = {}, {}, {}
        self._h, self._next, self._keep = {}, 1, keep_below

        def _open(user, name, size_hint, out):
            h = self._next
            self._next += 1
            nm = name.decode()
            self._h[h] = nm
            self.hash[nm] = hashlib.sha256()
            self.size[nm] = 0
            self.small[nm] = bytearray()
            out[0] = h
            return 0

        def _pwrite(user, handle, offset, data, length):
            nm = self._h[handle]
            assert offset == self.size[nm]
            b = C.string_at(data, length)
            self.hash[nm].update(b)
            self.size[nm] += length
            if self.size[nm] <= self._keep:
                self.small[nm] += b
            return 0

        def _close(user, handle):
            self._h.pop(handle, None)
            return 0

        self._cbs = (G._OPEN_FN(_open), G._PWRITE_FN(_pwrite), G._CLOSE_FN(_close))
        self.c = G.Sink(None, *self._cbs)

    def digest(self):
        return {n: (self.size[n], self.hash[n].hexdigest()) for n in self.hash}


def _build(kind, k, genome, read_len, n_reads, err, chunk_reads, min_count=1, max_batch_keys=0):
    t_start = time.perf_counter()
    b = G.Builder(kind, k, min_count=min_count, max_batch_keys=max_batch_keys)
    done = 0
    i = 0
    while done < n_reads:
        n = min(chunk_reads, n_reads - done)
        text = S.reads_fastq(genome, read_len, n, err=err, seed=1000 + i, first_idx=done)
        b.push(text, G.FASTQ)
        done += n
        i += 1
    counts = b.finish()
    sink = HashSink()
    b.emit("out", sink)
    st = b.stats()
    b.close()
    d = st.as_dict()
    print(f"[big] kind={kind} k={k} m={min_count} reads={n_reads} x {read_len} max_batch_keys={max_batch_keys}: "
          f"{counts.n_instances} instances, {counts.n_distinct} distinct, {counts.n_kept} kept, {st.n_batches} batches, "
          f"{time.perf_counter() - t_start:.1f} s wall (read simulation included); device ms: "
          + ", ".join(f"{key[3:]} {d[key]:.1f}" for key in ("ms_scan", "ms_extract", "ms_sort", "ms_reduce", "ms_merge", "ms_unfold", "ms_emit"))
          + f"; hbm peak {st.hbm_peak_bytes / 1e9:.1f} GB; bytes out {st.bytes_out}", flush=True)
    return counts, st, sink


def test_config2_full_size_single_vs_multi_batch():
    """BASELINE configs[1]: k=31, m=2, 5 Mbp genome, 1,666,667 x 150 bp, e=1 % -> 396,666,746 instances."""
    g = S.genome(5_000_000, 42)
    c1, st1, s1 = _build(G.GRAPH, 31, g, 150, 1_666_667, 0.01, 600_000, min_count=2)
    assert c1.n_instances == 1_666_667 * 119 * 2 == 396_666_746
    assert c1.n_kept < c1.n_distinct < c1.n_instances and st1.n_batches == 1
    c2, st2, s2 = _build(G.GRAPH, 31, g, 150, 1_666_667, 0.01, 600_000, min_count=2, max_batch_keys=60_000_000)     # 198.3 M folded keys: four batches
    assert st2.n_batches >= 3
    assert (c2.n_instances, c2.n_distinct, c2.n_kept) == (c1.n_instances, c1.n_distinct, c1.n_kept)
    assert s1.digest() == s2.digest()                                     # merging batches changes no byte
    # header / plane consistency
    d = s1.digest()
    assert struct.unpack("<3Q", bytes(s1.small["out.header"])) == (2011101014, 31, 0)
    ver, D, qD = struct.unpack("<3Q", bytes(s1.small["out-edges.header"])[:24])
    count = struct.unpack("<Q", bytes(s1.small["out-edges.header"])[56:64])[0]
    assert ver == 2012030501 and count == c1.n_kept and qD == 8 * ((D + 7) // 8)
    assert d["out-counts.ord0"][0] == c1.n_kept
    low = sum(sz for n, (sz, _) in d.items() if n.startswith("out-edges.low-bits"))
    assert low == c1.n_kept * qD // 8
    hist = dict(tuple(map(int, l.split(b"\t"))) for l in bytes(s1.small["out-counts-hist.txt"]).splitlines())
    assert sum(hist.values()) == c1.n_kept and min(hist) >= 2
    assert sum(k * v for k, v in hist.items()) <= c1.n_instances


def test_config3_quarter_kmer_set_batched():
    """BASELINE configs[2] at one quarter: build-kmer-set k=25, 25 M x 100 bp reads (2.5 Gbases) from a 100 Mbp genome."""
    g = S.genome(100_000_000, 42)
    n_reads = 25_000_000
    c1, st1, s1 = _build(G.KMERSET, 25, g, 100, n_reads, 0.01, 2_000_000)
    assert c1.n_instances == n_reads * 76
    c2, st2, s2 = _build(G.KMERSET, 25, g, 100, n_reads, 0.01, 2_000_000, max_batch_keys=700_000_000)
    assert st2.n_batches >= 3 and c2.n_kept == c1.n_kept
    assert s1.digest() == s2.digest()
    assert struct.unpack("<3Q", bytes(s1.small["out.header"])) == (2011101701, 25, c1.n_kept)


def test_config4_scaled_128bit_keys():
    """BASELINE configs[3] scaled to 5 M reads: k=55 (112-bit keys in 16-byte words), 150 bp reads, 100 Mbp genome."""
    g = S.genome(100_000_000, 42)
    n_reads = 5_000_000
    c1, st1, s1 = _build(G.GRAPH, 55, g, 150, n_reads, 0.01, 1_000_000)
    assert c1.n_instances == n_reads * 95 * 2 and st1.sort_key_bytes == 16
    c2, st2, s2 = _build(G.GRAPH, 55, g, 150, n_reads, 0.01, 1_000_000, max_batch_keys=150_000_000)                 # 475 M folded keys: four batches
    assert st2.n_batches >= 3 and s1.digest() == s2.digest()


@pytest.mark.skipif(os.environ.get("GSB_BIG_FULL") != "1", reason="set GSB_BIG_FULL=1 as well: 20 GB of reads, several minutes")
def test_config3_full_size_kmer_set():
    """BASELINE configs[2] in full: build-kmer-set k=25 on 10 Gbases (100 M x 100 bp reads, 100 Mbp genome, e=1 %):
    7.6 G instances = 61 GB of keys, more than one batch on a 180 GB GPU."""
    g = S.genome(100_000_000, 42)
    n_reads = 100_000_000
    c, st, s = _build(G.KMERSET, 25, g, 100, n_reads, 0.01, 4_000_000)
    assert c.n_instances == n_reads * 76 == 7_600_000_000
    assert st.n_batches >= 2 and 0 < c.n_kept < c.n_instances
    d = s.digest()
    assert struct.unpack("<3Q", bytes(s.small["out.header"])) == (2011101701, 25, c.n_kept)
    ver, D, qD = struct.unpack("<3Q", bytes(s.small["out.kmers.header"])[:24])
    assert ver == 2012030501 and struct.unpack("<Q", bytes(s.small["out.kmers.header"])[56:64])[0] == c.n_kept
    low = sum(sz for n, (sz, _) in d.items() if n.startswith("out.kmers.low-bits"))
    assert low == c.n_kept * qD // 8
    print("c3 full:", c.n_instances, "instances,", c.n_kept, "k-mers,", st.n_batches, "batches;", st.as_dict())
