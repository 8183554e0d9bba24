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


# ---- round 2's exchange design, modelled end to end on CPU ------------------------------------------------------------------
# (exchange.cu: exchange_partition_pull + exchange_pairs_msd).  The arithmetic that decides WHO OWNS WHAT is restated here in
# numpy -- equal shares of the bit-mixed key space for the instances; contiguous, balanced ranges of the 1024 children of the
# real key's top bits for the survivors (pair_owner_plan_kernel) -- and run over gloo with one process per rank; the GPU box runs
# the kernels themselves against the same oracle (tests/test_gpu_multi.py).

_M64 = (1 << 64) - 1


def _key_mix(z):
    """splitmix64 finaliser (gossamer_b200/csrc/keys.h: key_mix) on uint64 arrays."""
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _rc(keys, w):
    """reverse complement of w-symbol keys (w <= 32) held in uint64."""
    x = (~keys.astype(np.uint64)) & np.uint64((1 << (2 * w)) - 1 if w < 32 else _M64)
    out = np.zeros_like(x)
    for _ in range(w):
        out = (out << np.uint64(2)) | (x & np.uint64(3))
        x = x >> np.uint64(2)
    return out


def _exchange(send, world, rank):
    """ragged all-to-all of int64 matrices [rows, n_r] over gloo"""
    sizes = torch.tensor([s.shape[1] for s in send])
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    recv = [torch.zeros((send[0].shape[0], int(all_sizes[src][rank])), dtype=torch.int64) for src in range(world)]
    reqs = []
    for peer in range(world):
        if peer == rank:
            recv[peer].copy_(send[peer])
            continue
        reqs.append(dist.isend(send[peer].contiguous(), peer))
        reqs.append(dist.irecv(recv[peer], peer))
    for r in reqs:
        r.wait()
    return torch.cat(recv, dim=1).numpy()


def _worker_r2(rank, world, port, k, min_count, out):
    sys.path.insert(0, HERE)
    import oracle_py as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w = k + 1
    lo, hi, _ = O.extract([(_reads(200 + rank, 1200 + 400 * rank), O.FASTQ)], w, O.MODE_GRAPH)   # both strands, stream order
    x = lo[0::2].astype(np.uint64)                                  # one window per pair (x, rc x)
    folded = np.minimum(x, _rc(x, w))                               # strand folding: min(x, rc x)
    # 1. instance exchange: child c of the top bits0 mixed bits belongs to rank (c * N) >> bits0
    bits0 = 8
    mixed = _key_mix(folded)
    owner = ((mixed >> np.uint64(64 - bits0)).astype(np.int64) * world) >> bits0
    got = _exchange([torch.from_numpy(folded[owner == r].astype(np.int64)[None, :]) for r in range(world)], world, rank)[0].astype(np.uint64)
    # 2. local count, self-complement doubling, min-count filter (bucket_count_kernel)
    keys, counts = np.unique(got, return_counts=True)
    counts = counts.astype(np.int64)
    self_rc = _rc(keys, w) == keys
    counts[self_rc] *= 2
    keep = counts >= min_count
    keys, counts = keys[keep], counts[keep]
    # 3. survivors: both strands, split by the top 10 bits of the REAL key; children dealt out as balanced contiguous ranges
    rc = _rc(keys, w)
    both_k = np.concatenate([keys, rc[rc != keys]])
    both_c = np.concatenate([counts, counts[rc != keys]])
    key_bits, b10 = 2 * w, 10
    child = (both_k >> np.uint64(key_bits - b10)).astype(np.int64)
    hist = torch.from_numpy(np.bincount(child, minlength=1 << b10).astype(np.int64))
    hists = [torch.zeros_like(hist) for _ in range(world)]
    dist.all_gather(hists, hist)
    tot = sum(h.numpy() for h in hists)
    P = np.concatenate([[0], np.cumsum(tot)[:-1]])
    own = np.minimum(world - 1, (P * world) // max(1, int(tot.sum())))       # pair_owner_plan_kernel: floor(P[c] * N / T)
    assert (np.diff(own) >= 0).all()                                          # contiguous ranges, in key order
    dest = own[child]
    got2 = _exchange([torch.from_numpy(np.stack([both_k[dest == r].astype(np.int64), both_c[dest == r]])) for r in range(world)], world, rank)
    order = np.argsort(got2[0].astype(np.uint64), kind="stable")
    out[rank] = (got2[0].astype(np.uint64)[order], got2[1][order], own)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,min_count", [(2, 1), (2, 2), (3, 2)])
def test_round2_exchange_design_concatenates_to_the_oracles_edges(world, min_count):
    import oracle_py as O
    k = 20
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_r2, args=(world, _free_port(), k, min_count, out), nprocs=world, join=True)
    # single-process truth: the oracle's build-graph counting over every rank's reads
    lo = np.concatenate([O.extract([(_reads(200 + r, 1200 + 400 * r), O.FASTQ)], k + 1, O.MODE_GRAPH)[0] for r in range(world)])
    tk, tc = np.unique(lo.astype(np.uint64), return_counts=True)
    keep = tc >= min_count
    tk, tc = tk[keep], tc[keep]
    ck = np.concatenate([out[r][0] for r in range(world)])
    cc = np.concatenate([out[r][1] for r in range(world)])
    assert (np.diff(ck.astype(np.uint64)) > 0).all()                 # the slices concatenate into one strictly increasing run
    assert np.array_equal(ck, tk) and np.array_equal(cc, tc.astype(np.int64))
    sizes = [out[r][0].size for r in range(world)]
    assert max(sizes) < 1.3 * sum(sizes) / world                     # balanced to within a child's size
