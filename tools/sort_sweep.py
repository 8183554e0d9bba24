"""Times the radix sweep variants on random keys (device-generated).  Usage: python tools/sort_sweep.py [n] [bits]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gossamer_b200 as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 396_666_746
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tunings = [int(x) for x in os.environ["GSB_TUNINGS"].split(",")] if "GSB_TUNINGS" in os.environ else (range(8) if bits <= 64 else [0])
peak = 6551.7
kb = 8 if bits <= 64 else 16
for t in tunings:
    sw, tot, nsw = G.debug_sort_bench(n, bits, iters=3, tuning=t)
    gbs = 2 * n * kb / (sw * 1e-3) / 1e9
    print(json.dumps({"tuning": t, "n": n, "bits": bits, "sweep_ms": round(sw, 4), "sweeps": nsw, "sort_ms": round(tot, 3),
                      "sweep_GBs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}), flush=True)
