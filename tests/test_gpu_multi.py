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


def _reads(rank, n=20_000):
    g = S.genome(60_000, 42)
    return bytes(S.reads_fastq(g, 100, n + 3000 * rank, err=0.01, seed=43 + rank))


def _worker(rank, world, nccl_id, k, min_count, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import gossamer_b200 as G
    b = G.Builder(G.GRAPH, k, min_count=min_count, device=rank)
    b.attach(nccl_id, world, rank)
    b.push(_reads(rank), G.FASTQ)
    counts = b.finish()
    lo, hi, cn = b.counts_arrays()
    slice_info = (int(lo.size), (int(hi[0]) << 64 | int(lo[0])) if lo.size else None, (int(hi[-1]) << 64 | int(lo[-1])) if lo.size else None)
    b.gather_to_root()
    files = None
    if rank == 0:
        sink = G.MemorySink()
        b.emit("graph", sink)
        files = sink.as_bytes()
    out[rank] = (files, (counts.n_instances, counts.n_distinct, counts.n_kept), slice_info)
    b.close()


@pytest.mark.parametrize("k,min_count", [(25, 1), (31, 2), (55, 1)])
def test_multi_gpu_build_graph_bit_exact(k, min_count):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    nccl_id = G.make_nccl_id()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, nccl_id, k, min_count, out), nprocs=world, join=True)
    want, ost = O.build_graph([(_reads(r), O.FASTQ) for r in range(world)], k, min_count=min_count, threads=4)
    files, counts, _ = out[0]
    ref = want.files()
    assert set(files) == set(ref)
    assert not [n for n in ref if files[n] != ref[n]]
    assert counts == (ost.n_instances, ost.n_distinct, ost.n_kept)
    # slices are contiguous, ordered ranges of the global order
    last = -1
    for r in range(world):
        n, first, final = out[r][2]
        if n:
            assert first > last
            last = final
