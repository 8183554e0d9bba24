"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): every rank counts its own reads, the
NCCL all-to-all range-partitions the reduced runs, rank 0 gathers and emits; the files must be
byte-identical to the single-process oracle over all reads."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import gossamer_b200 as G
import oracle_py as O
import simreads_py as S

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _reads(rank, n=20_000, mode=""):
    g = S.genome(60_000, 42)
    text = bytes(S.reads_fastq(g, 100, n + 3000 * rank, err=0.01, seed=43 + rank))
    if mode == "skew" and rank == 0:
        # one k-mer repeated 1.5 M times on one rank: its owner's receive window cannot hold its share of a uniform split,
        # the partition exchange must decline (on every rank alike) and the sampled exchange take over
        text += b"".join(b"@p%d\n%s\n+\n%s\n" % (i, b"A" * 100, b"I" * 100) for i in range(22_000))
    return text


def _worker(rank, world, nccl_id, k, min_count, mode, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import gossamer_b200 as G
    kind = G.KMERSET if mode == "kmerset" else G.GRAPH
    if mode == "sampled":                                                # the fallback of the pair-sort exchange: sampled splitters
        G.debug_set_tuning(G.SAMPLED_SURVIVORS)
    if mode == "pairgeo":                                                # tiny buckets: more passes after the exchange, overflowing buckets
        G.debug_set_pairsort(64, -1)
    if mode == "pairovf":                                                # nothing fits: the whole-slice radix-sort fallback
        G.debug_set_pairsort(2, 10)
    limited = mode == "batches" or (mode == "onespill" and rank == 0)       # onespill: only ONE rank ever flushes a batch early
    b = G.Builder(kind, k, min_count=min_count, device=rank, max_batch_keys=400_000 if limited else 0)
    b.attach(nccl_id, world, rank)
    text = _reads(rank, mode=mode)
    if mode in ("batches", "onespill"):                                  # several sort+reduce+merge rounds per rank before the exchange
        recs = text.split(b"\n@r")
        chunks = [b"\n@r".join(recs[i:i + 4000]) for i in range(0, len(recs), 4000)]
        for i, ch in enumerate(chunks):
            b.push((b"" if i == 0 else b"@r") + ch + (b"\n" if i + 1 < len(chunks) else b""), G.FASTQ, last=True)
    else:
        b.push(text, G.FASTQ)
    counts = b.finish()
    lo, hi, cn = b.counts_arrays()
    slice_info = (int(lo.size), (int(hi[0]) << 64 | int(lo[0])) if lo.size else None, (int(hi[-1]) << 64 | int(lo[-1])) if lo.size else None)
    if mode == "gather":
        b.gather_to_root()
    sink = G.MemorySink()
    b.emit("graph", sink)                                  # collective: every rank hands over its own pieces of every file
    files = {n: (bytes(v), sink.segments[n]) for n, v in sink.files.items()}
    out[rank] = (files, (counts.n_instances, counts.n_distinct, counts.n_kept), slice_info)
    b.close()


def _assemble(per_rank):
    """Put the pieces of all ranks together; pieces must not overlap and must cover what the oracle wrote."""
    names = set()
    for files in per_rank:
        names |= set(files)
    whole = {}
    for n in names:
        size = max(len(files[n][0]) for files in per_rank if n in files)
        buf, covered = bytearray(size), np.zeros(size, np.uint8)
        for files in per_rank:
            if n not in files:
                continue
            data, segs = files[n]
            for off, ln in segs:
                assert not covered[off:off + ln].any(), f"{n}: two ranks wrote bytes {off}..{off + ln}"
                covered[off:off + ln] = 1
                buf[off:off + ln] = data[off:off + ln]
        whole[n] = bytes(buf)
    return whole


@pytest.mark.parametrize("k,min_count,mode", [(25, 1, "dist"), (31, 2, "dist"), (55, 1, "dist"), (31, 3, "gather"), (27, 2, "batches"),
                                              (5, 2, "dist"), (25, 1, "kmerset"), (40, 1, "kmerset"), (31, 2, "skew"), (25, 1, "onespill"),
                                              (55, 2, "onespill"), (31, 2, "sampled"), (55, 1, "sampled"), (31, 1, "pairgeo"), (55, 2, "pairgeo"),
                                              (25, 1, "pairovf"), (7, 1, "dist")])
def test_multi_gpu_build_bit_exact(k, min_count, mode):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    nccl_id = G.make_nccl_id()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, nccl_id, k, min_count, mode, out), nprocs=world, join=True)
    inputs = [(_reads(r, mode=mode), O.FASTQ) for r in range(world)]
    if mode == "kmerset":
        want, ost = O.build_kmer_set(inputs, k, threads=4, base="graph")
    else:
        want, ost = O.build_graph(inputs, k, min_count=min_count, threads=4)
    files = _assemble([out[r][0] for r in range(world)])
    ref = want.files()
    assert set(files) == set(ref)
    assert not [n for n in ref if files[n] != ref[n]]
    counts = out[0][1]
    if mode == "kmerset":
        assert (counts[0], counts[2]) == (ost.n_instances, ost.n_kept)
    else:
        assert counts == (ost.n_instances, ost.n_distinct, ost.n_kept)
    if mode != "gather":
        assert sum(1 for r in range(world) if out[r][0]) == world      # every rank wrote pieces
    # slices are contiguous, ordered ranges of the global order
    last = -1
    for r in range(world):
        n, first, final = out[r][2]
        if n:
            assert first > last
            last = final


def test_goss_cli_builds_on_several_gpus(tmp_path):
    """The C++ host drives the multi-GPU build: `goss build-graph --devices 0,1[,2,3]` forks one worker process per GPU (NCCL id
    over pipes), every worker pwrites its own byte ranges into the shared output files; the result is compared with the
    single-process oracle, byte for byte."""
    import subprocess
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    goss = os.path.join(os.path.dirname(HERE), "gossamer_b200", "goss")
    g = S.genome(200_000, 77)
    text = bytes(S.reads_fastq(g, 100, 60_000, err=0.01, seed=78))               # ~13 MB -> 13 blocks of 1 MiB over the workers
    fasta = (">c1\n" + "\n".join(bytes(g[i:i + 70]).decode() for i in range(0, 40_000, 70)) + "\n").encode()
    (tmp_path / "reads.fq").write_bytes(text)
    (tmp_path / "ref.fa").write_bytes(fasta)
    devices = ",".join(str(d) for d in range(world))

    def files(prefix):
        return {p.name: p.read_bytes() for p in tmp_path.iterdir() if p.name.startswith(prefix + ".") or p.name.startswith(prefix + "-")}

    for k, m in ((27, 2), (55, 1)):
        want, _ = O.build_graph([(fasta, O.FASTA), (text, O.FASTQ)], k, min_count=m, threads=4, base="g%d" % k)
        # a stale, larger file under the prefix must not leak into the result
        (tmp_path / ("g%d-edges.high-bits" % k)).write_bytes(b"\xff" * 10_000_000)
        r = subprocess.run([goss, "build-graph", "-k", str(k), "-m", str(m), "-i", "reads.fq", "-I", "ref.fa", "-O", "g%d" % k, "--block-mb", "1",
                            "--devices", devices, "-v"], capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        got, ref = files("g%d" % k), want.files()
        assert set(got) == set(ref)
        assert not [n for n in ref if got[n] != ref[n]]
    want, _ = O.build_kmer_set([(fasta, O.FASTA), (text, O.FASTQ)], 25, threads=4, base="ks")
    r = subprocess.run([goss, "build-kmer-set", "-k", "25", "-i", "reads.fq", "-I", "ref.fa", "-O", "ks", "--block-mb", "1", "--devices", "0-%d" % (world - 1)],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    got, ref = files("ks"), want.files()
    assert set(got) == set(ref) and not [n for n in ref if got[n] != ref[n]]
    # a parse error in one worker's block: every worker stops, exit code 1
    cut = text.index(b"\n", len(text) // 2) + 1
    (tmp_path / "bad.fq").write_bytes(text[:cut] + b"oops\n" + text[cut:])
    r = subprocess.run([goss, "build-graph", "-k", "27", "-i", "bad.fq", "-O", "b", "--block-mb", "1", "--devices", devices], capture_output=True, text=True,
                       cwd=tmp_path, timeout=120)
    assert r.returncode == 1 and "bad.fq" in r.stderr
