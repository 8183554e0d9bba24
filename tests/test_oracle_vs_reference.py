"""Pins the restated oracle against the REAL reference code: the reference's own Graph::Builder,
KmerSet::Builder and SparseArray::Builder (compiled unmodified from /root/reference/src with a
Boost shim, oracle/ref/) must write exactly the bytes the oracle writes, and the reference's own
Graph::open / select / rank / multiplicity must read them back.  Skipped when oracle/_ref is absent."""
import numpy as np
import pytest

import oracle_py as O
import ref_py as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libgossref.so not built (needs /root/reference)")


def _split(vals):
    return (np.array([v & (2**64 - 1) for v in vals], np.uint64), np.array([v >> 64 for v in vals], np.uint64))


def _diff(a, b):
    out = []
    for n in sorted(set(a) | set(b)):
        if a.get(n) != b.get(n):
            out.append((n, None if n not in a else len(a[n]), None if n not in b else len(b[n])))
    return out


@pytest.mark.parametrize("bits,m", [(20, 100), (32, 3000), (52, 20000), (52, 8192), (52, 8193), (64, 9000), (72, 500),
                                    (100, 700), (112, 9000), (126, 300), (56, 2), (30, 0), (40, 1)])
def test_sparse_array_bytes_equal_reference_writer(bits, m):
    rng = np.random.default_rng(bits * 1000 + m)
    vals = sorted({int.from_bytes(rng.bytes(16), "little") & ((1 << bits) - 1) for _ in range(m)})
    lo, hi = _split(vals)
    ours = O.write_sparse_array(lo, hi, 1 << bits, len(vals), base="x").files()
    theirs = R.write_sparse_array(lo, hi, 1 << bits, len(vals), base="x")
    assert not _diff(ours, theirs)


@pytest.mark.parametrize("density", [0.5, 0.02, 0.002, 0.00005])
def test_select_directories_all_block_classes_equal_reference(density):
    rng = np.random.default_rng(int(1 / density))
    n = 60_000
    gaps = rng.geometric(density, n).astype(np.uint64)
    gaps[::977] += np.uint64(1 << 26)
    vals = np.cumsum(gaps)
    bits = int(vals[-1]).bit_length() + 1
    ours = O.write_sparse_array(vals, None, 1 << bits, n // 50, base="x").files()
    theirs = R.write_sparse_array(vals, None, 1 << bits, n // 50, base="x")
    assert not _diff(ours, theirs)


@pytest.mark.parametrize("k", [15, 27, 31, 40, 55, 62])
def test_graph_bytes_equal_reference_builder(k):
    rng = np.random.default_rng(k)
    n = 30_000
    bits = 2 * (k + 1)
    vals = sorted({int.from_bytes(rng.bytes(16), "little") & ((1 << bits) - 1) for _ in range(n)})
    lo, hi = _split(vals)
    counts = rng.integers(1, 200, len(vals)).astype(np.uint64)
    counts[rng.integers(0, len(vals), 600)] = rng.integers(256, 65536, 600)
    counts[rng.integers(0, len(vals), 40)] = rng.integers(65536, 2**32, 40)
    counts[5] = 2**32 + 9                     # truncated in the array, 64-bit in the hist (src/Graph.hh:101-106)
    ours = O.write_graph(lo, hi, counts, k).files()
    theirs = R.write_graph(lo, hi, counts, k)
    assert not _diff(ours, theirs)
    # and the reference's own reader opens the oracle's files
    kk, rlo, rhi, rcn = R.read_graph(ours)
    assert kk == k and np.array_equal(rlo, lo) and np.array_equal(rhi, hi)
    assert np.array_equal(rcn, (counts & np.uint64(0xFFFFFFFF)).astype(np.uint32))


def test_graph_with_size_estimate_different_from_count():
    # M_est != M (the spill/merge regime passes the sum of part sizes, src/AsyncMerge.tcc:288-291)
    rng = np.random.default_rng(3)
    vals = np.sort(rng.choice(1 << 50, 5000, replace=False)).astype(np.uint64)
    counts = rng.integers(1, 500, 5000).astype(np.uint64)
    for m_est in (5000, 7000, 20000, 100):
        assert not _diff(O.write_graph(vals, None, counts, 24, m_est=m_est).files(), R.write_graph(vals, None, counts, 24, m_est=m_est))


@pytest.mark.parametrize("k", [25, 32, 33, 63])
def test_kmer_set_bytes_equal_reference_builder(k):
    rng = np.random.default_rng(k)
    bits = 2 * k
    vals = sorted({int.from_bytes(rng.bytes(16), "little") & ((1 << bits) - 1) for _ in range(20_000)})
    lo, hi = _split(vals)
    assert not _diff(O.write_kmer_set(lo, hi, k).files(), R.write_kmer_set(lo, hi, k))


def test_whole_command_output_opens_with_reference_reader():
    # reference known answers (src/testGossCmdBuildGraph.cc:115-179) through the reference's own Graph::open
    fs, _ = O.build_graph([(b">\nAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n", O.FASTA)], k=27)
    k, lo, hi, cn = R.read_graph(fs.files())
    assert k == 27 and len(lo) == 2 and list(cn) == [1, 1]
    fs, _ = O.build_graph([(b">\nNACTTTTGATGCAATGTCAAATTCTCCNCGTCATTCGCAACTGAATACAAGNGAATTTGGAAGGAGAATNTGGTA\n", O.FASTA)], k=15)
    assert len(R.read_graph(fs.files())[1]) == 42


# ---- whole commands: the reference's own GossCmdBuildGraph / GossCmdBuildKmerSet / GossCmdTrimGraph ---------
import simreads_py as S

FA, FQ, LN = 0, 1, 2


def _noisy_reads(seed, glen, n, rlen, err):
    g = S.genome(glen, seed)
    arr = bytearray(bytes(S.reads_fastq(g, rlen, n, err=err, seed=seed + 1)))
    rng = np.random.default_rng(seed)
    for p in rng.integers(0, len(arr), 200):
        if arr[p] in b"ACGT":
            arr[p] = ord("N") if p % 3 == 0 else arr[p] | 0x20
    return bytes(arr)


SMALL_INPUTS = [
    ([(b">\nAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n", FA)], 27),
    ([(b">\nNACTTTTGATGCAATGTCAAATTCTCCNCGTCATTCGCAACTGAATACAAGNGAATTTGGAAGGAGAATNTGGTA\n", FA)], 15),
    ([(b">1\nTTTT\n>2\nTTTTATGTACTATTATCTTATTTCTAAATATTAACTATAGTATCCCCTGGCGTTAATACAGCTCTAGAAATC\n", FA)], 14),
    ([(b">r\r\nACGTACGTACGTAAACCCGGGTTT\r\nACGTACGTTTTTGGGGCCCCAAAA\r\n", FA)], 7),
    ([(b">a\nACGTAC\nGTACGT\n\nAAACCCGGGTTT\n>b\n>c\nacgtnACGTAGGATCCAGGATTACCA", FA)], 5),
    ([(b"@r1\nACGTAC\nGTAGGCT\n+\nIIIIII\nIIIIIII\n@r2\nGGCCAATTGGCCAA\n+\nJJJJJJJJJJJJJJ\n", FQ)], 5),
    ([(b"@r1\nACGTACGTAGGCT\n+\n@IIIIIIIIIII+\n@r2\nGGCCAATTGGCCAA\n+\n+JJJJJJJJJJJJJ", FQ)], 5),
    ([(b"@r\r\nACGTACGTAGGCT\r\n+\r\nIIIIIIIIIIIII\r\n", FQ)], 5),
    ([(b"ACGTACGTAGGCTAGGA\n\nGGNNACGTAGGCTAGACCA\nAC", LN)], 5),
    ([(b"@r1\nACGTACGTAGGCT\n+\nIIIIIIIIIIIII\n", FQ), (b">x\nGGGGACGTACGTAGGCTTTT\n", FA), (b"ACGTACGTAGGCTAGGA\n", LN)], 6),
]


@pytest.mark.parametrize("case", range(len(SMALL_INPUTS)))
def test_build_graph_small_inputs_equal_reference_command(case):
    inputs, k = SMALL_INPUTS[case]
    _, theirs = R.build_graph(inputs, k, threads=1, log_slots=12)
    ours, _ = O.build_graph(inputs, k)
    assert not _diff(ours.files(), theirs)


@pytest.mark.parametrize("k,threads", [(15, 1), (25, 1), (31, 4), (32, 1), (55, 2), (62, 1)])
def test_build_graph_random_reads_equal_reference_command(k, threads):
    text = _noisy_reads(10 + k, 30_000, 6000, 100, err=0.01)
    _, theirs = R.build_graph([(text, FQ)], k, threads=threads, log_slots=22)
    ours, st = O.build_graph([(text, O.FASTQ)], k)
    assert st.n_distinct > 100_000
    assert not _diff(ours.files(), theirs)


def test_min_count_equals_reference_build_then_trim():
    # -m 2  ==  build-graph, then trim-graph -C 1 (src/GossCmdTrimGraph.cc:97-124)
    text = _noisy_reads(77, 20_000, 8000, 100, err=0.01)
    store, _ = R.build_graph([(text, FQ)], 31, threads=1, log_slots=22, base="g")
    for m in (2, 3, 5):
        theirs = R.trim_graph(store, "g", f"t{m}", m - 1)
        ours, st = O.build_graph([(text, O.FASTQ)], 31, min_count=m, base=f"t{m}")
        assert 0 < st.n_kept < st.n_distinct
        assert not _diff(ours.files(), theirs)


@pytest.mark.parametrize("k", [25, 32, 40, 63])
def test_build_kmer_set_equals_reference_command(k):
    g = S.genome(40_000, seed=k)
    fasta = (">g1\n" + "\n".join(bytes(g[i:i + 60]).decode() for i in range(0, 20_000, 60)) + "\n>g2 desc\n" +
             "\n".join(bytes(g[i:i + 71]).decode() for i in range(20_000, 40_000, 71)) + "\n").encode()
    _, theirs = R.build_kmer_set([(fasta, FA)], k, threads=1, log_slots=20)
    ours, _ = O.build_kmer_set([(fasta, O.FASTA)], k)
    assert not _diff(ours.files(), theirs)


@pytest.mark.parametrize("text,fmt", [
    (b"r1\nACGT\n+\nIIII\n", FQ), (b"@r1\nACGT\n", FQ), (b"@r1\nACGT\n@r2\n", FQ), (b"@r1\nACGT\n+r2\nIIII\n", FQ),
    (b"@r1\nACGT\n+\nIII\n", FQ), (b"ACGT\n", FA)])
def test_parse_errors_match_reference_command(text, fmt):
    with pytest.raises(RuntimeError) as re_:
        R.build_graph([(text, fmt)], 3, threads=1, log_slots=10)
    with pytest.raises(O.OracleParseError) as oe:
        O.build_graph([(text, fmt)], 3)
    assert str(oe.value) in str(re_.value)


@pytest.mark.parametrize("k,n_bases,n_subst", [(15, 3000, 40), (25, 20000, 300), (31, 8000, 100), (32, 8000, 100), (40, 5000, 60), (9, 2000, 30)])
def test_xenome_index_steps_equal_reference_commands(k, n_bases, n_subst):
    """merge-and-annotate-kmer-sets and compute-near-kmers (src/XenoApp.cc:62-76): the restated oracle writes the bytes the
    reference's own commands write -- the union kmer set, the two membership bit vectors, and the bit vectors after the
    near-k-mer pass (including its variant-mask and discarded-normalize quirks)."""
    from xeno_cases import related_references
    graft, host = related_references(n_bases, n_subst, k)
    st1, f1 = R.build_kmer_set([(graft, 0)], k, base="ga")
    st2, f2 = R.build_kmer_set([(host, 0)], k, base="ho")
    st = R.Store()
    st.put_all(f1)
    st.put_all(f2)
    theirs = R.merge_and_annotate(st, "ga", "ho", "both")
    both_in = dict(f1)
    both_in.update(f2)
    ours, stats = O.merge_and_annotate(both_in, "ga", "ho", "both")
    assert not _diff(ours, theirs)
    assert stats[3] == stats[0] + stats[1] - stats[2] and 0 < stats[2] < min(stats[0], stats[1])
    theirs2 = R.compute_near_kmers(st, "both", threads=3)
    ours2, gray = O.compute_near_kmers(ours, "both")
    assert not _diff(ours2, theirs2)
    before = sum(bin(b).count("1") for b in ours["both.lhs-bits"])
    after = sum(bin(b).count("1") for b in ours2["both.lhs-bits"])
    assert gray > 0 and before - after <= gray


@pytest.mark.parametrize("case", ["identical", "disjoint", "k63"])
def test_xenome_index_steps_edge_cases_equal_reference(case):
    """Both sets the same (everything common, nothing can turn gray), unrelated sets (nothing common), and the widest k."""
    from xeno_cases import related_references
    import simreads_py as S
    k = 63 if case == "k63" else 21
    graft, host = related_references(4000, 50, 11)
    if case == "identical":
        host = graft
    elif case == "disjoint":
        host = b">other\n" + bytes(S.genome(5000, 999)) + b"\n"
    st1, f1 = R.build_kmer_set([(graft, 0)], k, base="ga")
    st2, f2 = R.build_kmer_set([(host, 0)], k, base="ho")
    st = R.Store()
    st.put_all(f1)
    st.put_all(f2)
    theirs = R.merge_and_annotate(st, "ga", "ho", "both")
    both_in = dict(f1)
    both_in.update(f2)
    ours, stats = O.merge_and_annotate(both_in, "ga", "ho", "both")
    assert not _diff(ours, theirs)
    if case == "identical":
        assert stats[0] == stats[1] == stats[2] == stats[3]
    if case == "disjoint":
        assert stats[2] == 0
    theirs2 = R.compute_near_kmers(st, "both", threads=2)
    ours2, gray = O.compute_near_kmers(ours, "both")
    assert not _diff(ours2, theirs2)
    if case == "identical":
        assert gray == 0
