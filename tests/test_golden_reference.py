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
