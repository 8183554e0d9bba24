"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, byte for byte.

Every test here needs a B200 (`-m gpu`).  The oracle is the checker only.
"""
import struct

import os

import numpy as np
import pytest

import gossamer_b200 as G
import oracle_py as O
import simreads_py as S

pytestmark = pytest.mark.gpu


def _diff(a, b):
    """Human-readable difference of two {name: bytes} file sets."""
    msgs = []
    for n in sorted(set(a) | set(b)):
        if n not in a:
            msgs.append(f"{n}: missing from GPU output")
        elif n not in b:
            msgs.append(f"{n}: unexpected file")
        elif a[n] != b[n]:
            x, y = a[n], b[n]
            first = next((i for i in range(min(len(x), len(y))) if x[i] != y[i]), min(len(x), len(y)))
            msgs.append(f"{n}: sizes {len(x)} vs {len(y)}, first difference at byte {first}")
    return msgs


def _keys(lo, hi):
    return [int(l) | (int(h) << 64) for l, h in zip(lo, hi)]


def _fastq(reads):
    return "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)).encode()


def _random_reads(seed, glen, n, rlen, err=0.0):
    g = S.genome(glen, seed)
    return bytes(S.reads_fastq(g, rlen, n, err=err, seed=seed + 1))


# ---- sort ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,bits", [(0, 52), (1, 52), (1000, 52), (4096, 64), (4097, 64), (100_000, 64), (1_000_003, 52),
                                    (300_000, 112), (70_001, 126), (5000, 66)])
def test_radix_sort_matches_numpy(n, bits):
    rng = np.random.default_rng(n + bits)
    lo = rng.integers(0, 2**64, n, dtype=np.uint64)
    hi = rng.integers(0, 2**64, n, dtype=np.uint64)
    if bits <= 64:
        lo &= np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(2**64 - 1)
        hi[:] = 0
    else:
        hi &= np.uint64((1 << (bits - 64)) - 1)
    # make duplicates likely
    if n > 10:
        lo[n // 2:] = lo[: n - n // 2]
        hi[n // 2:] = hi[: n - n // 2]
    slo, shi, passes = G.debug_sort_keys(lo, hi, bits)
    order = np.lexsort((lo, hi))
    assert np.array_equal(slo, lo[order]) and np.array_equal(shi, hi[order])
    assert passes <= (bits + 7) // 8


def test_radix_sort_skewed_digits():
    # low-entropy keys (poly-A like): most digits constant -> passes are skipped, result still sorted
    rng = np.random.default_rng(3)
    lo = (rng.integers(0, 4, 200_000, dtype=np.uint64) << np.uint64(40)) | rng.integers(0, 3, 200_000, dtype=np.uint64)
    slo, shi, passes = G.debug_sort_keys(lo, None, 64)
    assert np.array_equal(slo, np.sort(lo)) and passes <= 2


# ---- extraction --------------------------------------------------------------------------------------
def _fold(keys):
    """The reference's stream is x, rc(x), x', rc(x'), ... (src/ReverseComplementAdapter.hh:34-55); the device
    extracts ONE key per window, the smaller of the pair (strand folding, csrc/fold.cu)."""
    assert len(keys) % 2 == 0
    return [min(a, b) for a, b in zip(keys[0::2], keys[1::2])]


CASES = [
    (b">\nAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n", G.FASTA, 27),
    (b">\nNACTTTTGATGCAATGTCAAATTCTCCNCGTCATTCGCAACTGAATACAAGNGAATTTGGAAGGAGAATNTGGTA\n", G.FASTA, 15),
    (b">1\nTTTT\n>2\nTTTTATGTACTATTATCTTATTTCTAAATATTAACTATAGTATCCCCTGGCGTTAATACAGCTCTAGAAATC\n", G.FASTA, 14),
    (b">r\r\nACGTACGTACGTAAACCCGGGTTT\r\nACGTACGTTTTTGGGGCCCCAAAA\r\n", G.FASTA, 7),
    (b">a\nACGTAC\nGTACGT\n\nAAACCCGGGTTT\n>b\n>c\nacgtnACGTAGGATCCAGGATTACCA", G.FASTA, 5),
    (b"@r1\nACGTACGTAGGCT\n+\nIIIIIIIIIIIII\n@r2\nGGCCAATTGGCCAA\n+r2\nJJJJJJJJJJJJJJ\n", G.FASTQ, 5),
    (b"@r1\nACGTAC\nGTAGGCT\n+\nIIIIII\nIIIIIII\n@r2\nGGCCAATTGGCCAA\n+\nJJJJJJJJJJJJJJ\n", G.FASTQ, 5),
    (b"@r1\nACGTACGTAGGCT\n+\n@IIIIIIIIIII+\n@r2\nGGCCAATTGGCCAA\n+\n+JJJJJJJJJJJJJ", G.FASTQ, 5),
    (b"@r\r\nACGTACGTAGGCT\r\n+\r\nIIIIIIIIIIIII\r\n", G.FASTQ, 5),
    (b"@r\nACGTACGTAGGCT\n+\nIIIIIIIIIIIII\n\n", G.FASTQ, 5),
    (b"ACGTACGTAGGCTAGGA\n\nGGNNACGTAGGCTAGACCA\nAC", G.LINE, 5),
    (b"", G.LINE, 5),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("kind", [G.GRAPH, G.KMERSET])
def test_extract_small_cases(case, kind):
    text, fmt, k = CASES[case]
    w = k + 1 if kind == G.GRAPH else k
    olo, ohi, oreads = O.extract([(text, fmt)], w, O.MODE_GRAPH if kind == G.GRAPH else O.MODE_KMERSET)
    glo, ghi, greads = G.debug_extract(text, fmt, kind, k)
    assert greads == oreads
    want = _keys(olo, ohi)
    if kind == G.GRAPH:
        want = _fold(want)
    assert sorted(_keys(glo, ghi)) == sorted(want)


@pytest.mark.parametrize("k", [15, 25, 31, 32, 33, 55, 62])
def test_extract_random_reads_all_key_widths(k):
    text = _random_reads(100 + k, 20_000, 3000, 100, err=0.02)
    # sprinkle Ns and lower case
    arr = bytearray(text)
    rng = np.random.default_rng(k)
    for p in rng.integers(0, len(arr), 300):
        if arr[p] in b"ACGT":
            arr[p] = ord("N") if p % 3 == 0 else arr[p] | 0x20
    text = bytes(arr)
    olo, ohi, _ = O.extract([(text, G.FASTQ)], k + 1, O.MODE_GRAPH)
    glo, ghi, _ = G.debug_extract(text, G.FASTQ, G.GRAPH, k)
    assert sorted(_keys(glo, ghi)) == sorted(_fold(_keys(olo, ohi)))
    if k <= 63:
        olo, ohi, _ = O.extract([(text, G.FASTQ)], k, O.MODE_KMERSET)
        glo, ghi, _ = G.debug_extract(text, G.FASTQ, G.KMERSET, k)
        assert sorted(_keys(glo, ghi)) == sorted(_keys(olo, ohi))


@pytest.mark.parametrize("text,msg", [
    (b"r1\nACGT\n+\nIIII\n", "expected '@' at beginning of line 1"),
    (b"@r1\nACGT\n", "expected sequence data or quality header at line 3"),
    (b"@r1\nACGT\n@r2\n", "expected '+' at beginning of line 3"),
    (b"@r1\nACGT\n+r2\nIIII\n", "quality title does not match sequence title at line 3"),
    (b"@r1\nACGT\n+\nIII\n", "length mistmatch between sequence and quality data just before line 5"),
    (b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n+\nIIIII\n", "length mistmatch between sequence and quality data just before line 9"),
])
def test_fastq_errors_match_reference_text(text, msg):
    with pytest.raises(G.ParseError) as e:
        G.debug_extract(text, G.FASTQ, G.GRAPH, 3)
    assert e.value.message == msg
    with pytest.raises(O.OracleParseError) as oe:
        O.frame([(text, O.FASTQ)])
    assert str(oe.value) == msg


def _irregular_fastq(n_rec, seed, deceptive=False, bad_at=None):
    """Thousands of lines of FASTQ that is NOT in the four-line layout: wrapped sequence and quality lines, empty reads,
    quality lines that start with '@' / '+'; deceptive: every quality block contains three well-formed fake records."""
    rng = np.random.default_rng(seed)
    g = bytes(S.genome(30_000, seed))
    out = []
    for i in range(n_rec):
        if deceptive:
            a = int(rng.integers(0, len(g) - 33))
            seq = g[a:a + 33]
            qual_lines = [b"@f", b"ACGT", b"+", b"IIII"] * 3                           # 33 characters that read like three records
            rec = b"@d%d\n" % i + seq[:17] + b"\n" + seq[17:] + b"\n+\n" + b"\n".join(qual_lines) + b"\n"
        else:
            n = int(rng.integers(0, 5)) * 37 if i % 13 == 0 else int(rng.integers(60, 140))   # now and then an empty read
            a = int(rng.integers(0, len(g) - 200))
            seq = g[a:a + n]
            width = [60, 1000, 25][i % 3]
            qual = bytearray(b"I" * n)
            if n > 61 and i % 5 == 0:
                qual[60 if width == 60 else 25] = ord("@")
            if n > 51 and i % 7 == 0:
                qual[50] = ord("+")
            sl = [seq[j:j + width] for j in range(0, n, width)] or [b""]
            ql = [bytes(qual[j:j + width]) for j in range(0, n, width)] or [b""]
            if bad_at is not None and i == bad_at:
                ql[-1] = ql[-1] + b"I"                                                   # quality one longer than the sequence
            rec = b"@q%d extra\n" % i + b"\n".join(sl) + b"\n+" + (b"q%d extra" % i if i % 4 == 0 else b"") + b"\n" + b"\n".join(ql) + b"\n"
        out.append(rec)
    return b"".join(out)


@pytest.mark.parametrize("kind,n_rec", [("mixed", 9000), ("deceptive", 6000), ("mixed", 700)])
def test_irregular_fastq_is_framed_in_parallel_and_exactly(kind, n_rec):
    # blocks of more than 2048 lines outside the four-line layout go through the speculate-and-verify framing (ingest.cu);
    # the deceptive input makes its anchors fail verification, so the one-thread machine takes over: same reads either way
    text = _irregular_fastq(n_rec, 3 if kind == "mixed" else 4, deceptive=kind == "deceptive")
    n_want = len(O.frame([(text, O.FASTQ)]))
    lo, hi, n_reads = G.debug_extract(text, G.FASTQ, G.GRAPH, 15)
    olo, ohi, on = O.extract([(text, O.FASTQ)], 16, O.MODE_GRAPH)
    assert n_reads == n_want == on == n_rec
    assert sorted(_keys(lo, hi)) == sorted(_fold(_keys(olo, ohi)))


def test_irregular_fastq_error_deep_inside_has_the_reference_text():
    text = _irregular_fastq(9000, 5, bad_at=7000)
    with pytest.raises(O.OracleParseError) as oe:
        O.frame([(text, O.FASTQ)])
    with pytest.raises(G.ParseError) as e:
        G.debug_extract(text, G.FASTQ, G.GRAPH, 15)
    assert e.value.message == str(oe.value) and "length mistmatch" in e.value.message


def test_fasta_error_matches_reference_text():
    with pytest.raises(G.ParseError) as e:
        G.debug_extract(b"ACGT\n", G.FASTA, G.GRAPH, 3)
    assert e.value.message == "expected '>' at beginning of line 0"


# ---- emission ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits,m", [(20, 100), (32, 3000), (52, 20000), (52, 200_000), (64, 9000), (72, 500), (100, 700),
                                    (112, 30000), (126, 300), (56, 2), (52, 8192), (52, 8193), (52, 16384), (30, 0)])
def test_sparse_array_files_bit_exact(bits, m):
    rng = np.random.default_rng(bits * 1000 + m)
    vals = sorted({int.from_bytes(rng.bytes(16), "little") & ((1 << bits) - 1) for _ in range(m)})
    lo = np.array([v & (2**64 - 1) for v in vals], np.uint64)
    hi = np.array([v >> 64 for v in vals], np.uint64)
    want = O.write_sparse_array(lo, hi, 1 << bits, len(vals), base="x").files()
    got = G.debug_emit_sparse_array(lo, hi if bits > 64 else None, 1 << bits, len(vals), base="x")
    assert not _diff(got, want)


@pytest.mark.parametrize("density", [0.5, 0.02, 0.002, 0.00005])
def test_sparse_array_all_select_block_classes(density):
    # clustered positions force small / intermediate / large select blocks in d0 and d1
    rng = np.random.default_rng(int(1 / density))
    n = 60_000
    gaps = rng.geometric(density, n).astype(np.uint64)
    gaps[::977] += np.uint64(1 << 26)                      # occasional huge jumps
    vals = np.cumsum(gaps)
    bits = int(vals[-1]).bit_length() + 1
    want = O.write_sparse_array(vals, None, 1 << bits, n // 50, base="x").files()   # small m_est -> small D -> wide spans
    got = G.debug_emit_sparse_array(vals, None, 1 << bits, n // 50, base="x")
    assert not _diff(got, want)
    for name in ("x-d0", "x-d1"):
        hdr = struct.unpack("<16Q", want[name][:128])
        assert hdr[8] == hdr[10] + hdr[12] + hdr[14]


def test_graph_files_bit_exact_with_wide_counts():
    rng = np.random.default_rng(5)
    n = 50_000
    counts = rng.integers(1, 200, n).astype(np.uint64)
    counts[rng.integers(0, n, 900)] = rng.integers(256, 65536, 900)
    counts[rng.integers(0, n, 60)] = rng.integers(65536, 2**32, 60)
    counts[7] = 2**32 + 5
    edges = np.sort(rng.choice(1 << 40, n, replace=False)).astype(np.uint64)
    want = O.write_graph(edges, None, counts, k=20).files()
    got = G.debug_emit_graph(edges, None, counts, k=20)
    assert not _diff(got, want)


# ---- whole command ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", range(len(CASES) - 1))
def test_build_graph_small_cases_bit_exact(case):
    text, fmt, k = CASES[case]
    want, ost = O.build_graph([(text, fmt)], k)
    sink, counts, stats = G.build_graph([(text, fmt)], k)
    assert not _diff(sink.as_bytes(), want.files())
    assert (counts.n_reads, counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_reads, ost.n_instances, ost.n_distinct, ost.n_kept)


def test_reference_known_answers_on_gpu():
    # src/testGossCmdBuildGraph.cc:115-179
    sink, counts, _ = G.build_graph([(CASES[0][0], G.FASTA)], 27)
    k, lo, hi, cn, total = O.read_graph(sink.as_bytes())
    assert counts.n_kept == 2 and total == 2 and _keys(lo, hi) == [0, (1 << 56) - 1] and list(cn) == [1, 1]
    sink, counts, _ = G.build_graph([(CASES[1][0], G.FASTA)], 15)
    assert counts.n_kept == 42 and O.read_graph(sink.as_bytes())[4] == 42


@pytest.mark.parametrize("k,min_count,err", [(25, 1, 0.0), (31, 1, 0.01), (31, 2, 0.01), (55, 1, 0.01), (55, 3, 0.02), (15, 1, 0.0)])
def test_build_graph_random_reads_bit_exact(k, min_count, err):
    text = _random_reads(7 * k + min_count, 50_000, 20_000, 100, err=err)
    want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=min_count, threads=4)
    sink, counts, stats = G.build_graph([(text, G.FASTQ)], k, min_count=min_count)
    assert not _diff(sink.as_bytes(), want.files())
    assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)
    # files open through the restated reference readers and satisfy lint-graph's symmetry invariant
    kk, lo, hi, cn, total = O.read_graph(sink.as_bytes(), exercise_select=(k == 25))
    assert kk == k and total == counts.n_kept
    got = dict(zip(_keys(lo, hi), map(int, cn)))
    for e in list(got)[:2000]:
        assert got[O.reverse_complement(e, k + 1)] == got[e]


@pytest.mark.parametrize("k", [3, 4, 5, 7, 9])
@pytest.mark.parametrize("min_count", [1, 2, 3, 4, 5])
def test_self_complementary_edges_and_min_count(k, min_count):
    # tiny k: a large share of the (k+1)-mers are their own reverse complement when k+1 is even; each of
    # their windows counts twice in the reference, which the folded counting must reproduce (also in the filter)
    text = _random_reads(1000 * k + min_count, 300, 400, 30, err=0.05)
    want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=min_count)
    sink, counts, _ = G.build_graph([(text, G.FASTQ)], k, min_count=min_count)
    assert not _diff(sink.as_bytes(), want.files())
    assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)
    # the same through several sort+reduce+merge rounds (doubling and filter applied after the last merge)
    b = G.Builder(G.GRAPH, k, min_count=min_count, max_batch_keys=3000)
    recs = text.split(b"\n@r")
    chunks = [b"\n@r".join(recs[i:i + 100]) for i in range(0, len(recs), 100)]
    for i, ch in enumerate(chunks):
        b.push((b"" if i == 0 else b"@r") + ch + (b"\n" if i + 1 < len(chunks) else b""), G.FASTQ, last=True)
    counts2 = b.finish()
    st = b.stats()
    sink2 = G.MemorySink()
    b.emit("graph", sink2)
    b.close()
    assert st.n_batches > 1
    assert not _diff(sink2.as_bytes(), want.files())
    assert (counts2.n_instances, counts2.n_distinct, counts2.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)


@pytest.mark.parametrize("k,min_count", [(31, 2), (25, 1), (55, 2), (40, 1)])
@pytest.mark.parametrize("max_slots,total_bits", [(0, 0), (256, 3), (1024, 0), (0, 20), (64, 12), (4096, 9)])
def test_partition_counting_any_bucket_geometry(k, min_count, max_slots, total_bits):
    # counting by partitioning (csrc/partition.cu): forcing few partition bits / small tables makes buckets overflow (they
    # then take the full-sort path), forcing many bits gives three passes over tiny buckets; whatever happens the files
    # must not change
    text = _random_reads(11 * k + min_count, 30_000, 12_000, 100, err=0.01)
    want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=min_count, threads=4)
    try:
        G.debug_set_partition(max_slots, total_bits)
        sink, counts, stats = G.build_graph([(text, G.FASTQ)], k, min_count=min_count)
    finally:
        G.debug_set_partition(0, 0)
    assert not _diff(sink.as_bytes(), want.files())
    assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)


@pytest.mark.parametrize("k,min_count,cap,bits", [(31, 2, 0, -1), (31, 1, 64, -1), (55, 1, 64, -1), (25, 1, 0, 3), (55, 2, 32, 6), (31, 1, 4, 1), (25, 2, 4, 8), (27, 1, 0, 0),
                                                  (62, 1, 16, -1), (15, 1, 0, 22)])
def test_pair_sort_geometry_does_not_matter(k, min_count, cap, bits):
    # the survivors are ordered by most-significant-digit passes + a shared-memory sort per bucket; forced capacities and
    # pass widths exercise one pass / two passes / three passes, buckets that overflow (radix-sorted on their own), the
    # whole-array fallback (more than 64 overflowing buckets: cap 4 with 8 bits) and more bits than the key has
    text = _random_reads(13 * k + min_count, 30_000, 12_000, 100, err=0.01)
    want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=min_count, threads=4)
    try:
        G.debug_set_pairsort(cap, bits)
        sink, counts, stats = G.build_graph([(text, G.FASTQ)], k, min_count=min_count)
    finally:
        G.debug_set_pairsort(0, -1)
    assert not _diff(sink.as_bytes(), want.files())
    assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)


def test_pair_sort_skewed_keys():
    # low-complexity reads: most distinct edges share a long prefix (A..A / T..T), so the top bits of the real key are far
    # from uniform -- oversize buckets take the radix sort, the rest the shared-memory sort; same files either way
    rng = np.random.default_rng(5)
    recs = []
    for i in range(6000):
        tail = "".join("ACGT"[x] for x in rng.integers(0, 4, 24))
        head = "A" * int(rng.integers(40, 76))
        recs.append(f"@s{i}\n{head}{tail}\n+\n{'I' * (len(head) + 24)}\n")
    text = "".join(recs).encode() + _random_reads(99, 20_000, 3_000, 100, err=0.01)
    for k in (31, 55):
        want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=1, threads=4)
        try:
            G.debug_set_pairsort(256, -1)
            sink, counts, stats = G.build_graph([(text, G.FASTQ)], k, min_count=1)
        finally:
            G.debug_set_pairsort(0, -1)
        assert not _diff(sink.as_bytes(), want.files())
        assert counts.n_kept == ost.n_kept


@pytest.mark.parametrize("k,min_count", [(31, 2), (25, 1), (55, 2)])
def test_legacy_lsd_counting_still_matches(k, min_count):
    # the round-1 counting path (full LSD sort of the raw keys + run-length reduce) stays as the overflow path of the
    # partition counting and as a cross-check: same files
    text = _random_reads(11 * k + min_count, 30_000, 12_000, 100, err=0.01)
    want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=min_count, threads=4)
    try:
        G.debug_set_tuning(G.LEGACY_COUNTING)
        sink2, counts2, stats2 = G.build_graph([(text, G.FASTQ)], k, min_count=min_count)
    finally:
        G.debug_set_tuning(0)
    assert not _diff(sink2.as_bytes(), want.files())
    assert (counts2.n_instances, counts2.n_distinct, counts2.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)
    assert stats2.sort_passes == stats2.sort_passes_model


def test_partition_counting_heavy_repeats():
    # one k-mer repeated 200,000 times (poly-A reads) next to ordinary reads: its bucket cannot fit a shared-memory table
    # and must come back through the overflow path with the exact count
    reads = _random_reads(77, 20_000, 3_000, 100, err=0.01)
    poly = b"".join(b"@p%d\n%s\n+\n%s\n" % (i, b"A" * 100, b"I" * 100) for i in range(3000))
    text = reads + poly
    for k, m in ((25, 1), (31, 2)):
        want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=m, threads=4)
        sink, counts, _ = G.build_graph([(text, G.FASTQ)], k, min_count=m)
        assert not _diff(sink.as_bytes(), want.files())
        assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)


@pytest.mark.parametrize("k", [25, 32, 40, 63])
def test_build_kmer_set_bit_exact(k):
    g = S.genome(60_000, seed=k)
    fasta = (">g1\n" + "\n".join(bytes(g[i:i + 60]).decode() for i in range(0, 30_000, 60)) + "\n>g2 desc\n" +
             "\n".join(bytes(g[i:i + 71]).decode() for i in range(30_000, 60_000, 71)) + "\n").encode()
    want, ost = O.build_kmer_set([(fasta, O.FASTA)], k, threads=2)
    sink, counts, _ = G.build_kmer_set([(fasta, G.FASTA)], k)
    assert not _diff(sink.as_bytes(), want.files())
    assert counts.n_instances == ost.n_instances and counts.n_kept == ost.n_kept


def test_fasta_split_into_blocks_matches_whole_file():
    g = S.genome(200_000, seed=9)
    lines = [">chr1"] + [bytes(g[i:i + 60]).decode() for i in range(0, 200_000, 60)]
    text = ("\n".join(lines) + "\n").encode()
    want, _ = O.build_graph([(text, O.FASTA)], 27)
    b = G.Builder(G.GRAPH, 27)
    # split at line boundaries into 7 blocks; windows straddle the block edges
    cuts = [0] + [text.index(b"\n", len(text) * i // 7) + 1 for i in range(1, 7)] + [len(text)]
    for i in range(7):
        b.push(text[cuts[i]:cuts[i + 1]], G.FASTA, last=(i == 6))
    b.finish()
    sink = G.MemorySink()
    b.emit("graph", sink)
    b.close()
    assert not _diff(sink.as_bytes(), want.files())


def test_overlapped_block_pushes_match_one_block():
    # GSB_BLOCK_ASYNC: the copy of block i+1 overlaps the device work of block i; same files, errors surface one call later
    import torch
    text = _random_reads(21, 40_000, 30_000, 100, err=0.01)
    want, ost = O.build_graph([(text, O.FASTQ)], 31, min_count=2, threads=4)
    host = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    cuts = [0] + [text.index(b"\n@r", len(text) * i // 5) + 1 for i in range(1, 5)] + [len(text)]
    b = G.Builder(G.GRAPH, 31, min_count=2)
    for rep in range(2):
        for i in range(5):
            b.push_pointer(host.data_ptr() + cuts[i], cuts[i + 1] - cuts[i], G.FASTQ, last=True, overlap=True)
        counts = b.finish()
        sink = G.MemorySink()
        b.emit("graph", sink)
        assert not _diff(sink.as_bytes(), want.files())
        assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)
        b.reset()
    bad = torch.frombuffer(bytearray(b"@r1\nACGT\n+\nIII\n"), dtype=torch.uint8).pin_memory()
    b.push_pointer(bad.data_ptr(), bad.numel(), G.FASTQ, last=True, overlap=True)      # accepted: not looked at yet
    with pytest.raises(G.ParseError):
        b.finish()
    b.close()


def test_multiple_inputs_and_formats():
    a = _random_reads(1, 30_000, 5000, 80)
    g = S.genome(30_000, seed=1)
    fa = (">x\n" + bytes(g[:5000]).decode() + "\n").encode()
    ln = b"\n".join(bytes(g[i:i + 90]) for i in range(0, 20_000, 45)) + b"\n"
    inputs = [(a, G.FASTQ), (fa, G.FASTA), (ln, G.LINE)]
    want, _ = O.build_graph(inputs, 21)
    sink, counts, _ = G.build_graph(inputs, 21)
    assert not _diff(sink.as_bytes(), want.files())


def test_multi_batch_merge_is_identical():
    text = _random_reads(77, 40_000, 30_000, 100, err=0.01)
    want, ost = O.build_graph([(text, O.FASTQ)], 31, min_count=2, threads=4)
    b = G.Builder(G.GRAPH, 31, min_count=2, max_batch_keys=1_500_000)     # forces several sort+reduce+merge rounds
    recs = text.split(b"\n@r")
    chunks = [b"\n@r".join(recs[i:i + 5000]) for i in range(0, len(recs), 5000)]
    for i, ch in enumerate(chunks):
        ch = (b"" if i == 0 else b"@r") + ch + (b"\n" if i + 1 < len(chunks) else b"")
        b.push(ch, G.FASTQ, last=True)
    counts = b.finish()
    st = b.stats()
    sink = G.MemorySink()
    b.emit("graph", sink)
    b.close()
    assert st.n_batches > 1
    assert not _diff(sink.as_bytes(), want.files())
    assert (counts.n_instances, counts.n_distinct, counts.n_kept) == (ost.n_instances, ost.n_distinct, ost.n_kept)


def test_reset_and_reuse_context():
    text = _random_reads(5, 20_000, 4000, 100)
    want, _ = O.build_graph([(text, O.FASTQ)], 25)
    b = G.Builder(G.GRAPH, 25)
    for _ in range(3):
        b.push(text, G.FASTQ)
        b.finish()
        sink = G.MemorySink()
        b.emit("graph", sink)
        assert not _diff(sink.as_bytes(), want.files())
        b.reset()
    b.close()


# ---- the C++ host executable ------------------------------------------------------------------------------
def test_goss_cli_writes_the_reference_file_set(tmp_path):
    """`goss build-graph / build-kmer-set` (gossamer_b200/host, C++): option surface of the reference commands, files on disk
    byte-identical to the oracle; small --block-mb so that the overlapped block pipeline runs through many blocks."""
    import subprocess
    goss = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gossamer_b200", "goss")
    text = _random_reads(5, 200_000, 40_000, 100, err=0.01)                      # ~8.7 MB of FASTQ -> 9 blocks of 1 MiB
    g = S.genome(50_000, seed=8)
    fasta = (">c1 something\n" + "\n".join(bytes(g[i:i + 70]).decode() for i in range(0, 50_000, 70)) + "\n").encode()
    fq, fa = tmp_path / "reads.fq", tmp_path / "ref.fa"
    fq.write_bytes(text)
    fa.write_bytes(fasta)

    def files(prefix):
        out = {}
        for p in tmp_path.iterdir():
            if p.name.startswith(prefix) and p.name not in ("reads.fq", "ref.fa", "reads.fq.gz", "ref.fa.bz2"):
                out[p.name] = p.read_bytes()
        return out

    want, _ = O.build_graph([(fasta, O.FASTA), (text, O.FASTQ)], 27, min_count=2, threads=4, base="g")   # FASTA inputs come first
    r = subprocess.run([goss, "build-graph", "-k", "27", "-m", "2", "-i", str(fq), "-I", str(fa), "-O", str(tmp_path / "g"), "--block-mb", "1", "-v"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not _diff(files("g"), want.files())
    want, _ = O.build_kmer_set([(fasta, O.FASTA)], 25, base="ks")
    r = subprocess.run([goss, "build-kmer-set", "-k", "25", "-I", str(fa), "-O", str(tmp_path / "ks")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not _diff(files("ks"), want.files())
    # compressed inputs (src/PhysicalFileFactory.cc:261-280): .gz through zlib, .bz2 through libbz2 -- same files
    import bz2
    import gzip
    (tmp_path / "reads.fq.gz").write_bytes(gzip.compress(text, 1))
    (tmp_path / "ref.fa.bz2").write_bytes(bz2.compress(fasta[:20_000], 5) + bz2.compress(fasta[20_000:], 9))
    want, _ = O.build_graph([(fasta, O.FASTA), (text, O.FASTQ)], 27, min_count=2, threads=4, base="z")
    r = subprocess.run([goss, "build-graph", "-k", "27", "-m", "2", "-i", str(tmp_path / "reads.fq.gz"), "-I", str(tmp_path / "ref.fa.bz2"), "-O", str(tmp_path / "z"),
                        "--block-mb", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not _diff(files("z"), want.files())
    # a parse error deep inside the file: the reference's message, exit code 1
    bad = tmp_path / "bad.fq"
    cut = text.index(b"\n", len(text) // 2) + 1
    bad.write_bytes(text[:cut] + b"oops\n" + text[cut:])
    r = subprocess.run([goss, "build-graph", "-k", "27", "-i", str(bad), "-O", str(tmp_path / "b"), "--block-mb", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "bad.fq" in r.stderr and "error performing build-graph" in r.stderr


def test_goss_cli_wrapped_fastq_across_blocks(tmp_path):
    """Multi-line (wrapped) FASTQ records, quality lines that start with '@' or '+', streamed as 1 MiB blocks by the C++ host:
    the block reader finds record boundaries with the reference's own state machine (src/FastqParser.hh:78-176) when the
    four-line heuristic finds none, and the device frames the irregular blocks; files identical to the oracle's."""
    import subprocess
    goss = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gossamer_b200", "goss")
    rng = np.random.default_rng(12)
    g = bytes(S.genome(40_000, 9))
    recs = []
    for i in range(16_000):
        a = int(rng.integers(0, len(g) - 150))
        seq = g[a:a + 150]
        qual = bytearray(b"I" * 150)
        if i % 7 == 0:
            qual[60] = ord("@")                               # a wrapped quality line that starts with '@'
        if i % 11 == 0:
            qual[120] = ord("+")
        recs.append(b"@w%d\n" % i + b"\n".join(seq[j:j + 60] for j in range(0, 150, 60)) + b"\n+\n"
                    + b"\n".join(bytes(qual[j:j + 60]) for j in range(0, 150, 60)) + b"\n")
    text = b"".join(recs)                                     # ~5.3 MB -> six blocks
    (tmp_path / "wrapped.fq").write_bytes(text)
    want, _ = O.build_graph([(text, O.FASTQ)], 25, min_count=1, threads=4, base="w")
    r = subprocess.run([goss, "build-graph", "-k", "25", "-i", str(tmp_path / "wrapped.fq"), "-O", str(tmp_path / "w"), "--block-mb", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {p.name: p.read_bytes() for p in tmp_path.iterdir() if p.name.startswith("w.") or p.name.startswith("w-")}
    assert not _diff(got, want.files())


def test_k_range_and_call_order_errors():
    with pytest.raises(G.GossamerError) as e:
        G.Builder(G.GRAPH, 63)
    assert e.value.status == -7 and "unable to build a graph with k=63" in e.value.message
    b = G.Builder(G.GRAPH, 25)
    with pytest.raises(G.GossamerError):
        b.emit("graph", G.MemorySink())
    b.close()


# ---- config-scale properties (full BASELINE sizes, size-independent checks) -------------------------------
def test_config1_full_size_properties():
    """BASELINE configs[0]: k=25, 200k x 100bp reads, 1 Mbp genome, e=0 -> 30,000,000 instances."""
    g = S.genome(1_000_000, 42)
    text = S.reads_fastq(g, 100, 200_000, err=0.0, seed=43)
    b = G.Builder(G.GRAPH, 25)
    b.push(text, G.FASTQ)
    counts = b.finish()
    lo, hi, cn = b.counts_arrays()
    sink = G.MemorySink()
    b.emit("graph", sink)
    b.close()
    assert counts.n_instances == 200_000 * 75 * 2 == 30_000_000
    assert int(cn.sum()) == counts.n_instances                       # checksum of counts
    assert np.all(lo[1:] > lo[:-1])                                   # strictly sorted, distinct
    # error-free reads: every edge or its reverse complement is a genome 26-mer
    assert counts.n_kept <= 2 * (1_000_000 - 25)
    k, rlo, rhi, rcn, total = O.read_graph(sink.as_bytes(), exercise_select=False)
    assert np.array_equal(rlo, lo) and np.array_equal(rcn.astype(np.uint64), cn) and total == counts.n_kept
    # and it is bit-exact with the oracle at this size too (the oracle takes a few seconds here)
    want, _ = O.build_graph([(text, O.FASTQ)], 25, threads=8)
    assert not _diff(sink.as_bytes(), want.files())


# ---- straight against the REAL reference (oracle/_ref/libgossref.so, prebuilt; travels with the repo) -------
import ref_py as R

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libgossref.so was not shipped")


@needs_ref
@pytest.mark.parametrize("k,threads", [(25, 1), (31, 4), (55, 2)])
def test_gpu_files_equal_the_reference_commands_own_output(k, threads):
    text = _random_reads(31 * k, 40_000, 10_000, 100, err=0.01)
    store, theirs = R.build_graph([(text, 1)], k, threads=threads, log_slots=22, base="graph")
    sink, counts, _ = G.build_graph([(text, G.FASTQ)], k)
    assert not _diff(sink.as_bytes(), theirs)
    # -m 2 against the reference's build-graph + trim-graph -C 1
    trimmed = R.trim_graph(store, "graph", "t", 1)
    sink2, counts2, _ = G.build_graph([(text, G.FASTQ)], k, min_count=2, prefix="t")
    assert not _diff(sink2.as_bytes(), trimmed)
    # and the reference's own reader opens the GPU's files
    kk, lo, hi, cn = R.read_graph(sink2.as_bytes(), base="t")
    assert kk == k and len(lo) == counts2.n_kept


@needs_ref
@pytest.mark.parametrize("k", [25, 40])
def test_gpu_kmer_set_equals_the_reference_commands_own_output(k):
    g = S.genome(50_000, seed=k)
    fasta = (">g\n" + "\n".join(bytes(g[i:i + 60]).decode() for i in range(0, 50_000, 60)) + "\n").encode()
    _, theirs = R.build_kmer_set([(fasta, 0)], k, threads=2, log_slots=20)
    sink, _, _ = G.build_kmer_set([(fasta, G.FASTA)], k)
    assert not _diff(sink.as_bytes(), theirs)
