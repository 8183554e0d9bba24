#!/usr/bin/env python
"""Per-kernel share of ONE build step from an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv ...).

A step starts at a `count_newlines_kernel` launch (first kernel of gsb_push_*_block); `--step i` picks the i-th
step of the list (default: the last complete device-resident one, i.e. one block per step).
usage: tools/launch_shares.py gpurun_out/launches.csv [--step I] > profiles/rNN_launch_shares_c2.txt
"""
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"gsb::", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("unsigned long long", "u64").replace("unsigned int", "u32").replace("unsigned char", "u8")
    return name


def main():
    path = sys.argv[1]
    step = None
    if "--step" in sys.argv:
        step = int(sys.argv[sys.argv.index("--step") + 1])
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")) * 1e-6))
    starts = [i for i, r in enumerate(rows) if r[1].startswith("count_newlines_kernel") or "::count_newlines_kernel" in r[1]]
    # device-resident steps push ONE block: a step is the launches between two starts whose next start is also a step start
    spans = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
    if not spans:
        raise SystemExit("no complete step in the list")
    if step is None:
        # the last span that contains an emit kernel (a whole build, not one block of a multi-block push)
        whole = [s for s in spans if any("high_bits_kernel" in rows[j][1] for j in range(s[0], s[1]))]
        span = whole[-1] if whole else spans[-1]
    else:
        span = spans[step]
    sel = rows[span[0]:span[1]]
    total = sum(r[2] for r in sel)
    agg = {}
    for _, n, ms in sel:
        a = agg.setdefault(short(n), [0.0, 0])
        a[0] += ms
        a[1] += 1
    print(f"# per-kernel share of one build step (launch ids {sel[0][0]}..{sel[-1][0]} of {path}); ncu --metrics gpu__time_duration.sum "
          f"--clock-control none: cold-cache, serialised launches -- compare SHARES, not absolutes")
    print(f"# {len(sel)} launches, {total:.3f} ms in total")
    for n, (ms, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:8.3f} ms {100 * ms / total:5.1f}%  x{cnt:<3d} {n}")


if __name__ == "__main__":
    main()
