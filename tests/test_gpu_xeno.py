"""xenome index, steps 3 and 4 on the GPU: merge-and-annotate-kmer-sets and compute-near-kmers through the C ABI
and through the C++ `goss`, byte for byte against the oracle and against the reference's own commands (oracle/_ref)."""
import os
import subprocess

import numpy as np
import pytest

import gossamer_b200 as G
import oracle_py as O
import ref_py as R
from xeno_cases import related_references

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _diff(a, b):
    return [(n, None if n not in a else len(a[n]), None if n not in b else len(b[n])) for n in sorted(set(a) | set(b)) if a.get(n) != b.get(n)]


def _popcount(b):
    return int(np.unpackbits(np.frombuffer(b, np.uint8)).sum())


@pytest.mark.parametrize("k,n_bases,n_subst", [(15, 3000, 40), (25, 60000, 900), (31, 20000, 300), (32, 20000, 300), (33, 20000, 300), (47, 9000, 120),
                                                (63, 9000, 120), (9, 2000, 30)])
def test_xenome_index_steps_bit_exact(k, n_bases, n_subst):
    graft, host = related_references(n_bases, n_subst, k)
    fg = O.build_kmer_set([(graft, O.FASTA)], k, base="ga")[0].files()
    fh = O.build_kmer_set([(host, O.FASTA)], k, base="ho")[0].files()
    both_in = dict(fg)
    both_in.update(fh)
    want, wstats = O.merge_and_annotate(both_in, "ga", "ho", "both")
    got, stats = G.merge_and_annotate_kmer_sets(both_in, "ga", "ho", "both")
    assert not _diff(got, want)
    assert stats == wstats and stats[2] > 0
    want2, wgray = O.compute_near_kmers(want, "both")
    got2, gray = G.compute_near_kmers(want, "both")
    assert not _diff(got2, want2)
    assert gray == wgray and gray > 0
    assert _popcount(want["both.lhs-bits"]) - _popcount(got2["both.lhs-bits"]) <= gray
    if R.available():
        st = R.Store()
        st.put_all(both_in)
        assert not _diff(got, R.merge_and_annotate(st, "ga", "ho", "both"))
        assert not _diff(got2, R.compute_near_kmers(st, "both", threads=2))


@pytest.mark.parametrize("case", ["identical", "disjoint"])
def test_xenome_index_steps_edge_cases(case):
    import simreads_py as S
    k = 21
    graft, host = related_references(4000, 50, 11)
    if case == "identical":
        host = graft
    else:
        host = b">other\n" + bytes(S.genome(5000, 999)) + b"\n"
    fg = O.build_kmer_set([(graft, O.FASTA)], k, base="ga")[0].files()
    fh = O.build_kmer_set([(host, O.FASTA)], k, base="ho")[0].files()
    both_in = dict(fg)
    both_in.update(fh)
    want, wstats = O.merge_and_annotate(both_in, "ga", "ho", "both")
    got, stats = G.merge_and_annotate_kmer_sets(both_in, "ga", "ho", "both")
    assert not _diff(got, want) and stats == wstats
    want2, wgray = O.compute_near_kmers(want, "both")
    got2, gray = G.compute_near_kmers(want, "both")
    assert not _diff(got2, want2) and gray == wgray
    if case == "identical":
        assert stats[2] == stats[3] and gray == 0
    else:
        assert stats[2] == 0


def test_xenome_errors():
    graft, host = related_references(2000, 20, 5)
    fg = O.build_kmer_set([(graft, O.FASTA)], 21, base="ga")[0].files()
    fh = O.build_kmer_set([(host, O.FASTA)], 23, base="ho")[0].files()
    both = dict(fg)
    both.update(fh)
    with pytest.raises(G.GossamerError) as e:                         # different k (the reference: throw "nonsense")
        G.merge_and_annotate_kmer_sets(both, "ga", "ho", "both")
    assert "k=" in e.value.message or "nonsense" in e.value.message
    empty = O.build_kmer_set([(b">e\nACGT\n", O.FASTA)], 21, base="em")[0].files()
    both = dict(fg)
    both.update(empty)
    with pytest.raises(G.GossamerError) as e:
        G.merge_and_annotate_kmer_sets(both, "ga", "em", "both")
    assert "nonsense" in e.value.message
    with pytest.raises(G.GossamerError):                              # no bit vectors next to the set
        G.compute_near_kmers(fg, "ga")


def test_goss_cli_xenome_index_pipeline(tmp_path):
    """The four steps of `xenome index` (src/XenoApp.cc:62-76) with the C++ host: two build-kmer-set, merge-and-annotate-kmer-sets,
    compute-near-kmers (which rewrites the bit vectors in place)."""
    goss = os.path.join(ROOT, "gossamer_b200", "goss")
    k = 25
    graft, host = related_references(80_000, 1200, 77)
    (tmp_path / "graft.fa").write_bytes(graft)
    (tmp_path / "host.fa").write_bytes(host)

    def run(*args):
        r = subprocess.run([goss] + list(args), capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        return r

    run("build-kmer-set", "-k", str(k), "-I", "graft.fa", "-O", "idx-graft")
    run("build-kmer-set", "-k", str(k), "-I", "host.fa", "-O", "idx-host")
    r = run("merge-and-annotate-kmer-sets", "-G", "idx-graft", "-G", "idx-host", "-O", "idx-both", "-v")
    fg = O.build_kmer_set([(graft, O.FASTA)], k, base="idx-graft")[0].files()
    fh = O.build_kmer_set([(host, O.FASTA)], k, base="idx-host")[0].files()
    both_in = dict(fg)
    both_in.update(fh)
    want, st = O.merge_and_annotate(both_in, "idx-graft", "idx-host", "idx-both")
    assert r.stdout == "%d\t%d\t%d\n" % st[:3]                                # the line the reference prints (:204)
    assert "writing out %d kmers." % st[3] in r.stderr

    def files():
        return {p.name: p.read_bytes() for p in tmp_path.iterdir() if p.name.startswith("idx-both.")}

    assert not _diff(files(), want)
    r = run("compute-near-kmers", "-G", "idx-both", "-T", "4", "-v")
    want2, gray = O.compute_near_kmers(want, "idx-both")
    after = dict(want)
    after.update(want2)
    assert not _diff(files(), after)
    assert "found %d gray bits (out of %d)." % (gray, st[3]) in r.stderr
    # usage errors
    r = subprocess.run([goss, "merge-and-annotate-kmer-sets", "-G", "idx-graft", "-O", "x"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "exactly twice" in r.stderr
