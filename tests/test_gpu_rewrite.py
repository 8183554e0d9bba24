"""GPU parity of the commands that REWRITE existing file sets -- trim-graph, merge-graphs, merge-kmer-sets, dump-graph,
restore-graph -- against the reference's own commands (oracle/_ref, compiled unmodified).  The GPU side decodes the
succinct files in bulk (csrc/reader.cu), merges / filters the runs and writes the new file set with the ordinary emitters."""
import numpy as np
import pytest

import gossamer_b200 as G
import ref_py as R
import simreads_py as S

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="oracle/_ref was not built")]


def _reads(seed, glen, n, rlen, err=0.01):
    g = S.genome(glen, seed)
    return bytes(S.reads_fastq(g, rlen, n, err=err, seed=seed + 1))


def _ref_graph(text, k, base):
    store, files = R.build_graph([(text, 1)], k, threads=2, log_slots=22, base=base)
    return store, files


def _diff(a, b):
    return sorted(n for n in set(a) | set(b) if a.get(n) != b.get(n))


@pytest.mark.parametrize("k,cutoff", [(27, 1), (27, 3), (31, 2), (55, 1), (19, 40)])
def test_trim_graph_matches_reference(k, cutoff):
    text = _reads(500 + k, 30_000, 9_000, 100)
    store, files = _ref_graph(text, k, "g")
    want = R.trim_graph(store, "g", "t", cutoff)
    got = G.trim_graph(files, "g", "t", cutoff)
    assert not _diff(got, want)


def test_trim_graph_wide_counts():
    # counts above 255 and above 65535 exercise the ord1 / ord2 planes and their presence sets in the READER
    reads = _reads(9, 3_000, 30_000, 100, err=0.0)
    poly = b"".join(b"@p%d\n%s\n+\n%s\n" % (i, b"ACGT" * 25, b"I" * 100) for i in range(3000))
    store, files = _ref_graph(reads + poly, 21, "g")
    for cutoff in (1, 300, 70_000):
        want = R.trim_graph(store, "g", "t%d" % cutoff, cutoff)
        got = G.trim_graph(files, "g", "t%d" % cutoff, cutoff)
        assert not _diff(got, want)


@pytest.mark.parametrize("k,n_in,max_merge", [(27, 2, 8), (31, 3, 8), (55, 2, 8), (25, 5, 2), (25, 9, 8)])
def test_merge_graphs_matches_reference(k, n_in, max_merge):
    store = None
    files = {}
    names = []
    for i in range(n_in):
        text = _reads(40 + i, 20_000, 3_000 + 500 * i, 100)
        st, f = _ref_graph(text, k, f"in{i}")
        if store is None:
            store = st
        else:
            store.put_all(f)
        files.update(f)
        names.append(f"in{i}")
    want = R.merge_graphs(store, names, "out", max_merge=max_merge)
    got = G.merge_file_sets(files, names, "out", G.GRAPH, max_merge=max_merge)
    assert not _diff(got, want)


@pytest.mark.parametrize("k", [25, 40])
def test_merge_kmer_sets_matches_reference(k):
    store, files, names = None, {}, []
    for i in range(3):
        text = _reads(70 + i, 20_000, 3_000, 100)
        st, f = R.build_kmer_set([(text, 1)], k, threads=2, log_slots=22, base=f"s{i}")
        if store is None:
            store = st
        else:
            store.put_all(f)
        files.update(f)
        names.append(f"s{i}")
    want = R.merge_graphs(store, names, "u", kmer_sets=True)
    got = G.merge_file_sets(files, names, "u", G.KMERSET)
    assert not _diff(got, want)


@pytest.mark.parametrize("k", [21, 31, 47])
def test_dump_and_restore_match_reference(k):
    text = _reads(90 + k, 10_000, 2_000, 100)
    store, files = _ref_graph(text, k, "g")
    want_text = R.dump_graph(store, "g", "dump.txt")
    source = G.MemorySource(files)
    b = G.Builder(G.GRAPH, k)
    b.load("g", source)
    c = b.finish_loaded()
    sink = G.MemorySink()
    b.dump("dump.txt", sink)
    b.close()
    got_text = sink.as_bytes()["dump.txt"]
    assert got_text == want_text
    # restore-graph: the text back into a file set (host parses the lines, the device sorts and writes)
    want = R.restore_graph(store, want_text, "r")
    lines = want_text.split(b"\n")
    kk, n, flags = map(int, lines[1].split(b"\t"))
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    lo, hi, cn = [], [], []
    for ln in lines[2:]:
        if not ln:
            continue
        seq, cnt = ln.split(b"\t")
        v = 0
        for ch in seq:
            v = (v << 2) | code[ch]
        lo.append(v & (2**64 - 1)); hi.append(v >> 64); cn.append(int(cnt))
    b = G.Builder(G.GRAPH, kk)
    b.load_pairs(np.array(lo, np.uint64), np.array(hi, np.uint64), np.array(cn, np.uint64))
    b.finish_loaded(cutoff=0, m_est=n)
    sink = G.MemorySink()
    b.emit("r", sink)
    b.close()
    assert not _diff(sink.as_bytes(), want)
    assert c.n_kept == n


def test_goss_cli_rewrite_commands(tmp_path):
    """The C++ host: `goss trim-graph / merge-graphs / dump-graph / restore-graph` on real files, against the reference's
    own commands run on the same file sets."""
    import os
    import subprocess
    goss = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gossamer_b200", "goss")
    k = 27
    store = None
    for i in range(3):
        text = _reads(140 + i, 20_000, 4_000, 100)
        st, f = _ref_graph(text, k, f"in{i}")
        if store is None:
            store = st
        else:
            store.put_all(f)
        for n, v in f.items():
            (tmp_path / n).write_bytes(v)

    def files(prefix):
        return {p.name: p.read_bytes() for p in tmp_path.iterdir() if p.name.startswith(prefix + ".") or p.name.startswith(prefix + "-")}

    def run(*args):
        r = subprocess.run([goss, *args], capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        return r

    run("trim-graph", "-G", "in0", "-O", "t", "-C", "2", "-v")
    assert not _diff(files("t"), R.trim_graph(store, "in0", "t", 2))
    run("merge-graphs", "-G", "in0", "-G", "in1", "-G", "in2", "-O", "m")
    assert not _diff(files("m"), R.merge_graphs(store, ["in0", "in1", "in2"], "m"))
    run("merge-graphs", "-G", "in0", "-G", "in1", "-G", "in2", "--max-merge", "2", "-O", "m2")
    assert not _diff(files("m2"), R.merge_graphs(store, ["in0", "in1", "in2"], "m2", max_merge=2))
    run("dump-graph", "-G", "in1", "-o", "dump.txt")
    want_text = R.dump_graph(store, "in1", "dump.txt")
    assert (tmp_path / "dump.txt").read_bytes() == want_text
    r = subprocess.run([goss, "dump-graph", "-G", "in1"], capture_output=True, cwd=tmp_path)
    assert r.returncode == 0 and r.stdout == want_text
    run("restore-graph", "-f", "dump.txt", "-O", "r")
    assert not _diff(files("r"), R.restore_graph(store, want_text, "r"))
    # errors: unknown input, missing cutoff, k mismatch in a merge
    r = subprocess.run([goss, "trim-graph", "-G", "nope", "-O", "x", "-C", "1"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "error performing trim-graph" in r.stderr
    r = subprocess.run([goss, "trim-graph", "-G", "in0", "-O", "x"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "--cutoff" in r.stderr
    st2, f2 = _ref_graph(_reads(150, 20_000, 1_000, 100), 25, "other")
    for n, v in f2.items():
        (tmp_path / n).write_bytes(v)
    r = subprocess.run([goss, "merge-graphs", "-G", "in0", "-G", "other", "-O", "x"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "same kmer-size" in r.stderr
