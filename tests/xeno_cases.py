"""Inputs for the xenome-index tests (merge-and-annotate-kmer-sets + compute-near-kmers): a "graft" and a "host" reference
that share most of their sequence, so that shared k-mers, one-sided k-mers and near (gray) k-mers all occur."""
import numpy as np

import simreads_py as S


def related_references(n_bases, n_subst, seed, line=70):
    g = bytes(S.genome(n_bases, seed))
    rng = np.random.default_rng(seed + 1)
    h = bytearray(g)
    for p in rng.integers(0, n_bases, n_subst):
        h[p] = b"ACGT"[(b"ACGT".index(h[p]) + 1 + int(rng.integers(0, 3))) % 4]
    # the host also has sequence of its own
    h += bytes(S.genome(n_bases // 3, seed + 2))

    def fasta(name, seq):
        return (">%s\n" % name).encode() + b"\n".join(seq[i:i + line] for i in range(0, len(seq), line)) + b"\n"

    return fasta("graft", g), fasta("host", bytes(h))
