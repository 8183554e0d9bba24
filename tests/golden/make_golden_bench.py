"""Generates tests/golden/bench_golden.json: the REAL reference's per-file sha256 for the exact inputs bench.py
times on one GPU (oracle/_ref/libgossref_release.so = GossCmdBuildGraph [+ GossCmdTrimGraph -C m-1], unmodified
sources, single-pass regime: the hash table is sized so that it never spills, SURVEY.md section 8c "R1 / R1t").

    python tests/golden/make_golden_bench.py [c2 c1 c4s]

Run in the build container only (needs oracle/_ref, i.e. /root/reference); the JSON is committed and bench.py
compares the sha256 of every file its e2e sink received with it (`"parity"` in the bench line).  c2 at full size
takes ~12 GB of RAM and a few minutes of CPU."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import ref_py as R          # noqa: E402
import simreads_py as S     # noqa: E402


def main():
    import bench                       # the workload table and the read generator are bench.py's own
    names = sys.argv[1:] or ["c2", "c1", "c4s"]
    rel = R.PATH.replace("libgossref.so", "libgossref_release.so")
    assert os.path.exists(rel), "build oracle/_ref first (make -C oracle/ref)"
    R.PATH = rel
    path = os.path.join(HERE, "bench_golden.json")
    golden = json.load(open(path)) if os.path.exists(path) else {}
    for name in names:
        wl = bench.WORKLOADS[name]
        text = bytes(bench.make_reads(wl, 0, 1))
        n_inst = wl["n_reads"] * (wl["read_len"] - wl["k"]) * 2
        log_slots = max(16, min(31, n_inst.bit_length()))                  # >= one slot per INSTANCE: never spills
        t0 = time.time()
        store, files = R.build_graph([(text, 1)], wl["k"], threads=os.cpu_count() or 1, log_slots=log_slots, base="graph")
        if wl["min_count"] > 1:
            files = R.trim_graph(store, "graph", "t", wl["min_count"] - 1)
            files = {("graph" + n[1:]) if n.startswith("t") else n: v for n, v in files.items()}
        golden[name] = {
            "workload": wl["desc"], "k": wl["k"], "min_count": wl["min_count"], "genome": wl["genome"], "read_len": wl["read_len"],
            "n_reads": wl["n_reads"], "err": wl["err"], "seeds": [42, 43], "input_sha256": hashlib.sha256(text).hexdigest(),
            "input_bytes": len(text), "reference": "GossCmdBuildGraph" + (" + GossCmdTrimGraph -C %d" % (wl["min_count"] - 1) if wl["min_count"] > 1 else "")
                         + f", libgossref_release.so, log_slots={log_slots}",
            "files": {n: {"size": len(v), "sha256": hashlib.sha256(v).hexdigest()} for n, v in sorted(files.items())}}
        print(name, len(files), "files", sum(len(v) for v in files.values()), "bytes", f"{time.time() - t0:.0f} s", flush=True)
        del store, files
        with open(path, "w") as f:
            json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
