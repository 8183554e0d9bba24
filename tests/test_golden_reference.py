"""Golden vectors produced by the REAL reference (tests/golden/make_golden.py, run where
/root/reference exists) and committed as tests/golden/reference_golden.json: sha256 + size of every
output file of the reference's own build-graph (+ trim-graph) / build-kmer-set on fixed inputs.
The CPU oracle is checked against them everywhere; the CUDA path is checked on the GPU box."""
import hashlib
import json
import os

import pytest

import oracle_py as O
import simreads_py as S

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "reference_golden.json")))


def _inputs(entry):
    out = []
    for i in entry["inputs"]:
        if i["sim"]:
            seed, glen, n, rlen, err = i["sim"]
            g = S.genome(int(glen), int(seed))
            out.append((bytes(S.reads_fastq(g, int(rlen), int(n), err=err, seed=int(seed) + 1)), 1))
        else:
            out.append((bytes.fromhex(i["text_hex"]), i["format"]))
    return out


def _check(files, entry):
    want = entry["files"]
    assert sorted(files) == sorted(want)
    for n, meta in want.items():
        assert len(files[n]) == meta["size"], n
        assert hashlib.sha256(files[n]).hexdigest() == meta["sha256"], n
        if "hex" in meta:
            assert files[n] == bytes.fromhex(meta["hex"]), n


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_matches_reference_golden(name):
    e = GOLDEN[name]
    if e["command"] == "graph":
        fs, _ = O.build_graph(_inputs(e), e["k"], min_count=e["min_count"], base="g")
    else:
        fs, _ = O.build_kmer_set(_inputs(e), e["k"], base="g")
    _check(fs.files(), e)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_gpu_matches_reference_golden(name):
    import gossamer_b200 as G
    e = GOLDEN[name]
    if e["command"] == "graph":
        sink, _, _ = G.build_graph(_inputs(e), e["k"], min_count=e["min_count"], prefix="g")
    else:
        sink, _, _ = G.build_kmer_set(_inputs(e), e["k"], prefix="g")
    _check(sink.as_bytes(), e)


# ---- xenome index steps (tests/golden/make_golden_xeno.py) ---------------------------------------------------------------
XENO = json.load(open(os.path.join(HERE, "golden", "xeno_golden.json")))


def _xeno_inputs(e):
    from xeno_cases import related_references
    graft, host = related_references(e["n_bases"], e["n_subst"], e["seed"])
    fg = O.build_kmer_set([(graft, O.FASTA)], e["k"], base="ga")[0].files()
    fh = O.build_kmer_set([(host, O.FASTA)], e["k"], base="ho")[0].files()
    both = dict(fg)
    both.update(fh)
    return both


def _check_digest(files, want):
    assert sorted(files) == sorted(want)
    for n, meta in want.items():
        assert len(files[n]) == meta["size"] and hashlib.sha256(files[n]).hexdigest() == meta["sha256"], n


@pytest.mark.parametrize("name", sorted(XENO))
def test_oracle_xenome_steps_match_reference_golden(name):
    e = XENO[name]
    merged, _ = O.merge_and_annotate(_xeno_inputs(e), "ga", "ho", "both")
    _check_digest(merged, e["merged"])
    near, _ = O.compute_near_kmers(merged, "both")
    _check_digest(near, e["near"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(XENO))
def test_gpu_xenome_steps_match_reference_golden(name):
    import gossamer_b200 as G
    e = XENO[name]
    merged, _ = G.merge_and_annotate_kmer_sets(_xeno_inputs(e), "ga", "ho", "both")
    _check_digest(merged, e["merged"])
    near, _ = G.compute_near_kmers(merged, "both")
    _check_digest(near, e["near"])
