"""Generates tests/golden/reference_golden.json from the REAL reference (oracle/_ref/libgossref.so,
built by oracle/ref/Makefile from the unmodified sources under /root/reference/src).  Run in the
build container only (it needs /root/reference); the JSON it writes is committed so that the tests
on a box without the reference can still compare against the reference's own bytes.

    python tests/golden/make_golden.py

Each case stores the sha256 and size of every output file; cases marked `full` also store the file
bytes (hex) so that a mismatch can be diffed."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_py as R          # noqa: E402
import simreads_py as S     # noqa: E402

FA, FQ, LN = 0, 1, 2


def reads(seed, glen, n, rlen, err):
    g = S.genome(glen, seed)
    return bytes(S.reads_fastq(g, rlen, n, err=err, seed=seed + 1))


CASES = [
    # name, command, inputs (text, format), k, min_count, full
    ("polyA_tiny", "graph", [(b">\nAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n", FA)], 27, 1, True),
    ("read_with_ns", "graph", [(b">\nNACTTTTGATGCAATGTCAAATTCTCCNCGTCATTCGCAACTGAATACAAGNGAATTTGGAAGGAGAATNTGGTA\n", FA)], 15, 1, True),
    ("fastq_wrapped_at_quality", "graph", [(b"@r1\nACGTAC\nGTAGGCT\n+\n@IIIII\n+IIIIII\n@r2\nGGCCAATTGGCCAA\n+r2\nJJJJJJJJJJJJJJ\n", FQ)], 5, 1, True),
    ("mixed_formats", "graph", [(b"@r1\nACGTACGTAGGCT\n+\nIIIIIIIIIIIII\n", FQ), (b">x\r\nGGGGACGTacgtAGGCTTTT\r\n", FA), (b"ACGTACGTAGGCTAGGA\n\nAC", LN)], 6, 1, True),
    ("sim_k25", "graph", [("sim", (11, 30_000, 6_000, 100, 0.01))], 25, 1, False),
    ("sim_k31_m2", "graph", [("sim", (12, 30_000, 8_000, 100, 0.01))], 31, 2, False),
    ("sim_k55", "graph", [("sim", (13, 30_000, 5_000, 150, 0.01))], 55, 1, False),
    ("sim_k62_m3", "graph", [("sim", (14, 20_000, 6_000, 100, 0.02))], 62, 3, False),
    ("kmerset_k25", "kmerset", [("sim", (15, 30_000, 5_000, 100, 0.01))], 25, 1, False),
    ("kmerset_k63", "kmerset", [("sim", (16, 30_000, 3_000, 100, 0.0))], 63, 1, False),
]


def materialise(inputs):
    out = []
    for data, fmt in inputs:
        if data == "sim":
            out.append((reads(*fmt), FQ))
        else:
            out.append((data, fmt))
    return out


def main():
    assert R.available(), "build oracle/_ref first (make -C oracle/ref)"
    golden = {}
    for name, cmd, inputs, k, m, full in CASES:
        ins = materialise(inputs)
        if cmd == "graph":
            store, files = R.build_graph(ins, k, threads=1, log_slots=22, base="g")
            if m > 1:
                files = R.trim_graph(store, "g", "t", m - 1)
                files = {n.replace("t", "g", 1) if n.startswith("t") else n: v for n, v in files.items()}
        else:
            _, files = R.build_kmer_set(ins, k, threads=1, log_slots=22, base="g")
        entry = {"command": cmd, "k": k, "min_count": m,
                 "inputs": [{"format": f, "sim": list(inputs[i][1]) if inputs[i][0] == "sim" else None,
                             "text_hex": None if inputs[i][0] == "sim" else inputs[i][0].hex()} for i, (d, f) in enumerate(ins)],
                 "files": {n: {"size": len(v), "sha256": hashlib.sha256(v).hexdigest(), **({"hex": v.hex()} if full else {})}
                           for n, v in sorted(files.items())}}
        golden[name] = entry
        print(name, len(files), "files", sum(len(v) for v in files.values()), "bytes")
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
