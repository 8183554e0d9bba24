"""Summarise an ncu report: headline metrics of every captured launch + stall samples per SASS instruction of the first
launch whose kernel name contains `match` (default: the first launch).  Usage: ncu_stalls.py report.ncu-rep [top] [match]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
match = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__grid_size"]
for r in rows[2:]:
    print("---")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w}: {r[i][:120]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
take = match is None
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break
        take = match is None or (len(r) > 1 and match in r[1])
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if take and h and len(r) == len(h):
        data.append(r)
i_s, i_src = h.index("# Samples"), h.index("Source")
cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[i_s]) for r in data)
print("=== instructions", len(data), "samples", tot)
agg = {h[i]: sum(int(r[i]) for r in data) for i in cols}
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"{k}: {v} ({100 * v / tot:.1f}%)")
for j in sorted(sorted(range(len(data)), key=lambda j: -int(data[j][i_s]))[:top_n]):
    r = data[j]
    st = sorted(((h[i], int(r[i])) for i in cols if int(r[i]) > 0), key=lambda x: -x[1])[:2]
    print(j, r[i_s], f"{100 * int(r[i_s]) / tot:.1f}%", r[i_src].strip()[:64], st)
