"""CPU models (numpy) of the two algorithmic shortcuts the device takes on the counting path, checked against the oracle's
reference semantics (count every window AND its reverse complement, src/ReverseComplementAdapter.hh:34-55, then filter as
trim-graph does, src/GossCmdTrimGraph.cc:97-124).  They pin the ARGUMENT, independently of any CUDA code:

  * strand folding (csrc/fold.cu): count min(x, rc x) once per window, double self-complementary keys, filter, then add
    rc(y) for every surviving y != rc(y);
  * counting from a partial sort (csrc/sort.cu): keys bit-mixed with the splitmix64 finaliser, stably sorted on their
    low gb bits only; a group of equal low bits that holds one key is a run (length = count), groups with several keys
    are counted exactly on the side; the result does not depend on gb.
"""
import numpy as np
import pytest

import oracle_py as O
import simreads_py as S

M64 = (1 << 64) - 1


def _windows(k, n_reads=1500, glen=4000, rlen=60, err=0.02, seed=3):
    g = S.genome(glen, seed)
    text = bytes(S.reads_fastq(g, rlen, n_reads, err=err, seed=seed + 1))
    lo, hi, _ = O.extract([(text, O.FASTQ)], k + 1, O.MODE_GRAPH)          # x, rc(x), x', rc(x'), ...
    assert not hi.any()
    return lo


def _reference_counts(lo, k, m):
    rlo, rhi, rc = O.count(lo, None, 2 * (k + 1), min_count=m)
    return dict(zip(map(int, rlo), map(int, rc)))


def _fold_count_unfold(lo, k, m):
    folded = np.minimum(lo[0::2], lo[1::2])
    keys, cnt = np.unique(folded, return_counts=True)
    out = {}
    for y, c in zip(map(int, keys), map(int, cnt)):
        r = O.reverse_complement(y, k + 1)
        c = 2 * c if r == y else c
        if c >= m:
            out[y] = c
            out[r] = c
    return out


@pytest.mark.parametrize("k", [3, 4, 5, 9, 15, 31])
@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_strand_folding_reproduces_both_strand_counting(k, m):
    lo = _windows(k)
    assert _fold_count_unfold(lo, k, m) == _reference_counts(lo, k, m)


def _mix(z):
    z = z.copy()
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _unmix(z):
    z = z.copy()
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(31)) ^ (z >> np.uint64(62))
        z = z * np.uint64(0x319642B2D24D8EC3)
        z = z ^ (z >> np.uint64(27)) ^ (z >> np.uint64(54))
        z = z * np.uint64(0x96DE1B173F119089)
    return z ^ (z >> np.uint64(30)) ^ (z >> np.uint64(60))


def _count_from_partial_sort(folded, gb):
    """(key, occurrences) of every distinct folded key, from a stable sort on the low gb bits of the mixed keys only."""
    mixed = _mix(folded)
    grp = mixed & np.uint64((1 << gb) - 1) if gb < 64 else mixed
    order = np.argsort(grp, kind="stable")
    ms, gs = mixed[order], grp[order]
    n = ms.size
    head = np.ones(n, bool)
    head[1:] = gs[1:] != gs[:-1]
    differs = np.zeros(n, bool)
    differs[1:] = ~head[1:] & (ms[1:] != ms[:-1])                          # a different key inside a group
    starts = np.flatnonzero(head)
    ends = np.append(starts[1:], n)
    impure_group = np.add.reduceat(differs.astype(np.int64), starts) > 0
    out = {}
    n_impure = 0
    for s, e, bad in zip(starts, ends, impure_group):
        if not bad:                                                        # one key: the group is a run
            out[int(_unmix(ms[s:s + 1])[0])] = int(e - s)
        else:                                                              # several keys: counted exactly on the side
            n_impure += 1
            ks, cs = np.unique(_unmix(ms[s:e]), return_counts=True)
            for y, c in zip(map(int, ks), map(int, cs)):
                assert y not in out
                out[y] = c
    return out, n_impure


def test_key_mix_model_is_a_bijection():
    rng = np.random.default_rng(5)
    x = rng.integers(0, 1 << 63, 100_000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 100_000, dtype=np.uint64)
    assert np.array_equal(_unmix(_mix(x)), x)
    assert np.unique(_mix(x)).size == np.unique(x).size


@pytest.mark.parametrize("k", [5, 15, 31])
def test_counting_from_a_partial_sort_does_not_depend_on_the_group_bits(k):
    lo = _windows(k, n_reads=3000, glen=6000)
    folded = np.minimum(lo[0::2], lo[1::2])
    keys, cnt = np.unique(folded, return_counts=True)
    want = dict(zip(map(int, keys), map(int, cnt)))
    impure = {}
    for gb in (4, 8, 12, 16, 24, 32, 64):
        got, impure[gb] = _count_from_partial_sort(folded, gb)
        assert got == want, gb
    assert impure[64] == 0                                                 # all bits: a group IS a key
    if k >= 15:
        # log2(n) + 4 bits: only chance collisions are left (a key meets another one with probability D / 2^gb)
        n, d = folded.size, len(want)
        gb = int(np.ceil(np.log2(n))) + 4
        _, bad = _count_from_partial_sort(folded, gb)
        assert bad <= max(8, 4 * d * d / 2 ** (gb + 1))
