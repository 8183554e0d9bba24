#!/usr/bin/env python
"""Instructions executed and stall samples per SOURCE LINE of one kernel in an ncu report (needs -lineinfo and
--import-source on).  usage: tools/ncu_lines.py report.ncu-rep <kernel name substring> [top]"""
import csv
import io
import subprocess
import sys

rep, match = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, take, hdr, path, seen = {}, False, None, "", set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        take = match in r[1] and (path, r[1]) not in seen      # every (file, launch) section once: the first captured launch
        if take:
            seen.add((path, r[1]))
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if take and hdr and len(r) == len(hdr) and r[0].isdigit():
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        a = agg.setdefault((path, int(r[0])), [r[1].strip(), 0, 0])
        a[1] += int(r[i_inst] or 0)
        a[2] += int(r[i_samp] or 0)
ti = sum(a[1] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print(f"# {match}: {ti} warp instructions, {ts} stall samples; per source line (line: instr%, samples%, source)")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f}:{ln:<5d} {100 * a[1] / ti:5.1f}% {100 * a[2] / ts:5.1f}%  {a[0][:110]}")
