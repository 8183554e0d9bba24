"""world_size-2 (and 3) CPU test of the multi-GPU plan: sample -> splitters -> range partition ->
all-to-all -> merge, over the gloo backend.  The splitter planning is the product's own host code
(gsb_plan_splitters, no device needed); the per-rank counting is done by the CPU oracle here because
this container has no GPU -- on the GPU box tests/test_gpu_multi.py runs the same scenario through
NCCL.  The property checked is the one the design rests on: after the exchange rank r holds the r-th
contiguous slice of the global sorted (key,count) list, so the slices concatenate into exactly the
single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reads(seed, n):
    sys.path.insert(0, HERE)
    import simreads_py as S
    g = S.genome(30_000, 42)
    return bytes(S.reads_fastq(g, 100, n, err=0.01, seed=seed))


def _local_run(text, k):
    import oracle_py as O
    lo, hi, _ = O.extract([(text, O.FASTQ)], k + 1, O.MODE_GRAPH)
    return O.count(lo, hi, 2 * (k + 1))


def _worker(rank, world, port, k, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import gossamer_b200 as G
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    klo, khi, cnt = _local_run(_reads(100 + rank, 1500 + 500 * rank), k)
    m = klo.size
    # the sampling rule of exchange.cu::sample_kernel
    S = G.samples_per_rank()
    idx = np.minimum(((np.arange(S) + 0.5) * m / S).astype(np.int64), m - 1)
    mine = torch.from_numpy(np.stack([klo[idx], khi[idx]]).astype(np.int64))
    pooled = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(pooled, mine)
    plo = np.concatenate([p[0].numpy().astype(np.uint64) for p in pooled])
    phi = np.concatenate([p[1].numpy().astype(np.uint64) for p in pooled])
    slo, shi = G.plan_splitters(plo, phi, world)
    # partition bounds = lower_bound of each splitter (exchange.cu::bounds_kernel)
    keys = [(int(h) << 64) | int(l) for l, h in zip(klo, khi)]
    splitters = [(int(h) << 64) | int(l) for l, h in zip(slo, shi)]
    import bisect
    bounds = [0] + [bisect.bisect_left(keys, s) for s in splitters] + [m]
    send = [torch.from_numpy(np.stack([klo[bounds[r]:bounds[r + 1]], khi[bounds[r]:bounds[r + 1]], cnt[bounds[r]:bounds[r + 1]]]).astype(np.int64))
            for r in range(world)]
    sizes = torch.tensor([s.shape[1] for s in send])
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    recv = [torch.zeros((3, int(all_sizes[src][rank])), dtype=torch.int64) for src in range(world)]
    # all-to-all as pairwise send/recv (gloo has no all_to_all on CPU tensors of ragged shape)
    reqs = []
    for peer in range(world):
        if peer == rank:
            recv[peer].copy_(send[peer])
            continue
        reqs.append(dist.isend(send[peer].contiguous(), peer))
        reqs.append(dist.irecv(recv[peer], peer))
    for r in reqs:
        r.wait()
    got = torch.cat(recv, dim=1).numpy().astype(np.uint64)
    merged = {}
    for l, h, c in zip(got[0], got[1], got[2]):
        key = (int(h) << 64) | int(l)
        merged[key] = merged.get(key, 0) + int(c)
    mine_sorted = sorted(merged.items())
    out[rank] = (mine_sorted, splitters)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_range_partition_exchange_concatenates_to_global_order(world):
    k = 20
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, k, out), nprocs=world, join=True)
    # single-process truth over all ranks' reads
    truth = {}
    for rank in range(world):
        klo, khi, cnt = _local_run(_reads(100 + rank, 1500 + 500 * rank), k)
        for l, h, c in zip(klo, khi, cnt):
            key = (int(h) << 64) | int(l)
            truth[key] = truth.get(key, 0) + int(c)
    concat = []
    for rank in range(world):
        part, splitters = out[rank]
        lo_b = splitters[rank - 1] if rank > 0 else 0
        hi_b = splitters[rank] if rank < world - 1 else 1 << 128
        assert all(lo_b <= key < hi_b for key, _ in part)          # rank r owns [splitter[r-1], splitter[r])
        concat.extend(part)
    assert concat == sorted(truth.items())
    sizes = [len(out[r][0]) for r in range(world)]
    assert max(sizes) < 1.5 * (sum(sizes) / world)                 # sampled splitters balance the partitions
