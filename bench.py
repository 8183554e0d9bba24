#!/usr/bin/env python
"""bench.py -- build-graph k-mer edges/sec on B200 (BASELINE.json metric), one JSON line.

A "step" is one complete pass of the hot path over the workload: raw FASTQ text ->
line scan / framing -> 2-bit pack -> (k+1)-mer + reverse-complement extraction -> radix sort ->
run-length reduce -> min-count filter -> succinct Graph file set.

  value  : instances/s with the FASTQ text already resident in HBM and the files built in HBM
           (gsb_push_device_block ... gsb_emit(sink=NULL)), device-timed with CUDA events on the
           library's stream, max over ranks.
  e2e    : the same metric through the public C ABI with HOST buffers: pinned host text in
           (H2D inside the timed region), every output file copied back (D2H) and handed to the sink.
  roofline: the radix sweep kernel (dominant), algorithmic bytes = 2 * keys-per-launch * key_bytes (one folded key per
           window: keys-per-launch = n_inst / 2 for graphs).
  cpu_baseline: the reference's own build-graph (+ trim-graph) from oracle/_ref on a bounded sample of the reads.

Workload = BASELINE.json configs[1]: build-graph -k 31 -m 2, 5 Mbp random genome, 50x coverage of
150-bp reads, 1% substitution errors (synthetic, seeds 42/43).
`--impl reference` times the reference's own GossCmdBuildGraph (+ trim-graph for min-count), built
unmodified from /root/reference with a Boost shim into oracle/_ref (shipped to the GPU box as a
prebuilt .so), with -T = all host cores, on a bounded sample per step; without that library it
falls back to the CPU oracle port and says so (`kind`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (k, min_count, genome, read_len, n_reads, err)
    "c2": dict(k=31, min_count=2, genome=5_000_000, read_len=150, n_reads=1_666_667, err=0.01,
               desc="build-graph -k 31 -m 2, 5 Mbp random genome, 50x 150-bp reads, 1% errors"),
    "c1": dict(k=25, min_count=1, genome=1_000_000, read_len=100, n_reads=200_000, err=0.0,
               desc="build-graph -k 25, 200k 100-bp reads from a 1 Mbp random genome"),
    "c4s": dict(k=55, min_count=1, genome=5_000_000, read_len=150, n_reads=1_000_000, err=0.01,
                desc="build-graph -k 55 (128-bit keys), 1M 150-bp reads from a 5 Mbp genome (scaled-down configs[3])"),
}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_reads(wl, rank, world, out=None):
    """Weak scaling: per-GPU work is fixed, so the genome grows with the number of GPUs at constant
    coverage (N x 5 Mbp at 50x for c2, like BASELINE configs[4]'s large genome spread over 8 GPUs);
    every rank draws its reads from the whole genome."""
    import simreads_py as S
    g = S.genome(wl["genome"] * world, 42)
    return S.reads_fastq(g, wl["read_len"], wl["n_reads"], err=wl["err"], seed=43 + rank, out=out)


def _reference_lib():
    """The REAL reference (built unmodified from /root/reference with the Boost shim, release flags),
    if oracle/_ref/libgossref_release.so was shipped with the repo; else None."""
    try:
        import ref_py as R
        rel = R.PATH.replace("libgossref.so", "libgossref_release.so")
        if os.path.exists(rel):
            R.PATH = rel
            R.lib()
            return R
    except Exception:
        pass
    return None


def cpu_sample(wl, frac, threads):
    """Time the CPU implementation on the first `frac` of rank 0's reads.
    Preferred: the reference's own GossCmdBuildGraph (+ GossCmdTrimGraph -C m-1 for min-count m) with -T threads,
    kind "reference".  Fallback: the CPU oracle port, kind "port"."""
    import oracle_py as O
    import simreads_py as S
    n = max(1, int(wl["n_reads"] * frac))
    g = S.genome(wl["genome"], 42)
    text = S.reads_fastq(g, wl["read_len"], n, err=wl["err"], seed=43)
    n_inst = n * (wl["read_len"] - wl["k"]) * 2
    R = _reference_lib()
    if R is not None:
        log_slots = max(16, min(30, (4 * n_inst).bit_length()))       # table never spills: single-pass regime
        data = bytes(text)
        t0 = time.perf_counter()
        store, _ = R.build_graph([(data, 1)], wl["k"], threads=threads, log_slots=log_slots, base="g")
        if wl["min_count"] > 1:
            R.trim_graph(store, "g", "t", wl["min_count"] - 1)
        dt = time.perf_counter() - t0
        return n_inst, dt, n, "reference"
    t0 = time.perf_counter()
    fs, st = O.build_graph([(text, O.FASTQ)], wl["k"], min_count=wl["min_count"], threads=threads)
    dt = time.perf_counter() - t0
    return st.n_instances, dt, n, "port"


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frac = args.cpu_frac / 4
    for _ in range(args.warmup):
        cpu_sample(wl, frac / 8, threads)
    t_total, inst_total, n_sample, kind = 0.0, 0, 0, "port"
    for _ in range(args.steps):
        inst, dt, n_sample, kind = cpu_sample(wl, frac, threads)
        t_total += dt
        inst_total += inst
    value = inst_total / t_total
    what = ("data61/gossamer GossCmdBuildGraph + GossCmdTrimGraph, unmodified sources, -O3 -DNDEBUG, Boost shim, in-memory files"
            if kind == "reference" else "CPU oracle port (restatement, not the reference binary)")
    line = {
        "impl": "reference", "metric": "build-graph k-mer edges/sec", "value": value, "unit": "edge instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u128" if kind == "reference" else "u64",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "k": wl["k"], "min_count": wl["min_count"]},
        "cpu_baseline": {"value": value, "unit": "edge instances/s", "cores": threads, "kind": kind,
                         "sample": f"first {n_sample} of {wl['n_reads']} reads per step; {what}"},
        "e2e": {"value": value, "unit": "edge instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-frac", type=float, default=0.125, help="fraction of the reads the CPU baseline runs on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import gossamer_b200 as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gossamer_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- inputs: pinned host text (e2e) and a device-resident copy (value) -------------------------
    import simreads_py as S
    nbytes = S.fastq_bytes(wl["n_reads"], wl["read_len"])
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    text = make_reads(wl, rank, world, out=host.numpy())
    assert text.size == nbytes
    dev = host.to(f"cuda:{local_rank}", non_blocking=False)
    torch.cuda.synchronize()

    b = G.Builder(G.GRAPH, wl["k"], min_count=wl["min_count"], device=local_rank)
    if world > 1:
        idt = torch.zeros(G.NCCL_ID_BYTES, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(G.make_nccl_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        b.attach(bytes(idt.cpu().numpy().tobytes()), world, rank)

    def step_device():
        b.reset()
        b.push_device(dev.data_ptr(), nbytes, G.FASTQ)
        c = b.finish()
        b.emit("graph", None)          # multi-GPU: collective, every rank builds its own byte ranges of every file
        return c

    # end to end: the host program streams the file as blocks cut at record boundaries; with GSB_BLOCK_ASYNC the copy
    # of block i+1 overlaps scan / pack / extraction of block i
    n_blocks = 8
    cuts = [0]
    for i in range(1, n_blocks):
        t = nbytes * i // n_blocks
        window = bytes(text[t:t + 4096])
        cuts.append(t + window.index(b"\n@r") + 1)
    cuts.append(nbytes)

    def step_e2e(sink):
        b.reset()
        for i in range(n_blocks):
            b.push_pointer(host.data_ptr() + cuts[i], cuts[i + 1] - cuts[i], G.FASTQ, last=True, overlap=True)
        c = b.finish()
        b.emit("graph", sink)
        return c

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ---------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        counts = step_device()
    launches0 = b.stats().kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    b.timer_begin()
    sweeps_ms, sweeps_n, phase = 0.0, 0, {}
    a2a_ms, a2a_bytes = 0.0, 0
    for _ in range(args.steps):
        counts = step_device()
        st = b.stats()
        sweeps_ms += st.ms_sort_sweeps
        sweeps_n += st.sort_passes
        a2a_ms += st.ms_all_to_all
        a2a_bytes += st.exchange_bytes_sent
        for key in ("ms_scan", "ms_extract", "ms_sort", "ms_reduce", "ms_unfold", "ms_merge", "ms_exchange", "ms_emit"):
            phase[key] = phase.get(key, 0.0) + getattr(st, key)
        if os.environ.get("GSB_BENCH_TRACE"):
            print(f"[bench rank {rank}] step phases:", {k: round(getattr(st, k), 2) for k in ("ms_scan", "ms_extract", "ms_sort", "ms_reduce", "ms_unfold", "ms_exchange", "ms_emit")}, file=sys.stderr, flush=True)
    ms_dev = b.timer_end()
    barrier()
    clocks = sampler.stop()
    ms_dev = max_over_ranks(ms_dev)
    launches = b.stats().kernel_launches - launches0
    st = b.stats()
    n_inst_total = counts.n_instances if world > 1 else counts.n_instances   # finish() sums over ranks when attached
    n_inst_rank = n_inst_total // world
    value = n_inst_total * args.steps / (ms_dev * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    sink = G.MemorySink(keep_data=False)
    for _ in range(2):
        step_e2e(sink)
    barrier()
    b.timer_begin()
    t0 = time.perf_counter()
    e2e_phase = {}
    for _ in range(args.steps):
        step_e2e(sink)
        st2 = b.stats()
        for key in ("ms_h2d", "ms_scan", "ms_extract", "ms_sort", "ms_reduce", "ms_unfold", "ms_exchange", "ms_emit"):
            e2e_phase[key] = e2e_phase.get(key, 0.0) + getattr(st2, key)
    ms_e2e = b.timer_end()
    wall_e2e = time.perf_counter() - t0
    barrier()
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e * 1e3))
    d2h = b.stats().bytes_out
    e2e_value = n_inst_total * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        b.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (radix sweep), measured live ---------------------------------
    peak, peak_src = measured_peak_gbs()
    key_bytes = int(st.sort_key_bytes)
    sweep_ms = sweeps_ms / max(1, sweeps_n)
    # graph mode sorts ONE folded key per window (min of the window and its reverse complement): the sweep
    # kernel really moves n_sorted_keys keys per launch; B_sort below stays the reference-defined figure
    n_sorted = int(st.n_sorted_keys) or n_inst_rank
    sweep_bytes = 2.0 * n_sorted * key_bytes
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    passes_model = int(st.sort_passes_model)
    m_kept, m_distinct = counts.n_kept, counts.n_distinct
    b_sort = n_inst_rank * key_bytes * (2 + 2 * passes_model) + (m_distinct // world) * (key_bytes + 8)
    t_sort = (phase["ms_sort"] + phase["ms_reduce"] + phase["ms_unfold"]) / args.steps * 1e-3
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except Exception:
        pass

    line = {
        "metric": "build-graph k-mer edges/sec", "value": value, "unit": "edge instances/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64" if key_bytes == 8 else "u128",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "k": wl["k"], "min_count": wl["min_count"], "n_reads_per_gpu": wl["n_reads"],
                   "read_len": wl["read_len"], "genome": wl["genome"] * world, "err": wl["err"], "seeds": [42, 43],
                   "fastq_bytes_per_gpu": int(nbytes), "n_instances": int(n_inst_total), "n_distinct": int(m_distinct),
                   "n_edges_kept": int(m_kept), "parallelism": f"range-partition x{world}" if world > 1 else "single GPU",
                   "l2_note": "inputs (FASTQ text and key buffers) are larger than the 126 MB L2; no flush needed"},
        "gb_per_s": n_inst_total * key_bytes * args.steps / (ms_dev * 1e-3) / 1e9,
        "e2e": {"value": e2e_value, "unit": "edge instances/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(d2h),
                "input": f"{n_blocks} blocks cut at record boundaries, pinned host memory, GSB_BLOCK_ASYNC (copy of block i+1 overlaps the device work of block i)",
                "ms_per_step": ms_e2e / args.steps, "phases_ms_per_step": {k: v / args.steps for k, v in e2e_phase.items()}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "onesweep_kernel (one 8-bit radix sweep: read + write every key once)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "traffic": traffic, "launch_ms": sweep_ms, "bytes_per_launch": sweep_bytes, "keys_per_launch": n_sorted,
                     "sweeps_run_per_step": sweeps_n / args.steps, "sweeps_model_per_step": passes_model,
                     "sort_phase": {"what": "SURVEY 8d: B_sort = n_inst*keyB*(2+2P) + M*(keyB+8) over sort + reduce + unfold time. n_inst and P are "
                                            "the reference-defined figures (both strands, every digit); the device moves far fewer bytes -- one "
                                            "folded key per window, and with a min-count filter only enough low digits to group equal keys -- so "
                                            "this fraction can exceed 1; `achieved`/`frac` above are the per-launch figures of the sweep kernel",
                                    "b_sort_bytes": b_sort, "t_sort_ms": t_sort * 1e3,
                                    "achieved_gbs": b_sort / t_sort / 1e9 if t_sort > 0 else 0.0,
                                    "frac": (b_sort / t_sort / 1e9 / peak) if t_sort > 0 else 0.0}},
        "phases_ms_per_step": {k: v / args.steps for k, v in phase.items()},
        "exchange": None if world == 1 else {
            "what": "one all-to-all of raw instance keys routed to the rank owning their key range (rank 0's view)",
            "path": "fused partition+transfer kernel storing into peer windows over NVLink (CUDA IPC)" if b.stats().exchange_peer_memory
                    else "staged partition + grouped ncclSend/ncclRecv",
            "bytes_sent_per_gpu_per_step": a2a_bytes / args.steps, "all_to_all_ms": a2a_ms / args.steps,
            "gb_per_s_per_gpu": (a2a_bytes / max(a2a_ms, 1e-9)) / 1e6,
            "frac_of_nvlink_770": (a2a_bytes / max(a2a_ms, 1e-9)) / 1e6 / 770.0},
        "clocks": clocks,
        "hbm_peak_bytes": int(st.hbm_peak_bytes),
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        inst, dt, n_sample, kind = cpu_sample(wl, args.cpu_frac / 4, threads)
        what = ("the reference's own GossCmdBuildGraph + GossCmdTrimGraph (unmodified sources, -O3 -DNDEBUG, Boost shim)"
                if kind == "reference" else "CPU oracle port (restatement, not the reference binary)")
        line["cpu_baseline"] = {"value": inst / dt, "unit": "edge instances/s", "cores": threads, "kind": kind,
                                "sample": f"first {n_sample} of {wl['n_reads']} reads, {inst} instances in {dt:.2f} s; {what}"}
    print(json.dumps(line), flush=True)
    b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
