"""Generates tests/golden/xeno_golden.json from the REAL reference (oracle/_ref): sha256 + size of every file the reference's
own merge-and-annotate-kmer-sets writes, and of the two bit vectors after its compute-near-kmers, on fixed inputs
(tests/xeno_cases.py).  Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden_xeno.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_py as R                                   # noqa: E402
from xeno_cases import related_references            # noqa: E402

CASES = {"k15": (15, 3000, 40, 15), "k25": (25, 20000, 300, 25), "k32": (32, 8000, 100, 32), "k47": (47, 9000, 120, 47)}   # k, bases, substitutions, seed


def digest(files):
    return {n: {"size": len(v), "sha256": hashlib.sha256(v).hexdigest()} for n, v in sorted(files.items())}


def main():
    assert R.available(), "build oracle/_ref first (make -C oracle/ref)"
    golden = {}
    for name, (k, n_bases, n_subst, seed) in CASES.items():
        graft, host = related_references(n_bases, n_subst, seed)
        _, f1 = R.build_kmer_set([(graft, 0)], k, base="ga")
        _, f2 = R.build_kmer_set([(host, 0)], k, base="ho")
        st = R.Store()
        st.put_all(f1)
        st.put_all(f2)
        merged = R.merge_and_annotate(st, "ga", "ho", "both")
        near = R.compute_near_kmers(st, "both", threads=2)
        golden[name] = {"k": k, "n_bases": n_bases, "n_subst": n_subst, "seed": seed, "merged": digest(merged), "near": digest(near)}
        print(name, len(merged), "files")
    with open(os.path.join(HERE, "xeno_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
