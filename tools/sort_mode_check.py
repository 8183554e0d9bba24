"""Correctness + per-sweep time of sweep-kernel variants (gsb_debug_set_tuning ids: tile shapes / look-back widths, see
kShapes64 in csrc/sort.cu), e.g. `python tools/sort_mode_check.py 7 6 5`.  Sorts random and skewed key sets through the C ABI
with the FIRST variant and compares with numpy, then times 8 sweeps over 198.3 M 64-bit keys (the config-2 instance count)
for the default and every variant given."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gossamer_b200 as G

tunings = [int(a) for a in sys.argv[1:]] or [0]
tuning = tunings[0]
L = G.lib()
rng = np.random.default_rng(1)
ok = True
for n, bits in ((1, 64), (1000, 64), (4097, 64), (300_001, 52), (1_000_003, 64)):
    lo = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    if bits < 64:
        lo &= np.uint64((1 << bits) - 1)
    L.gsb_debug_set_tuning(tuning)
    slo, _, passes = G.debug_sort_keys(lo, None, bits)
    good = np.array_equal(slo, np.sort(lo))
    ok &= good
    print(f"tuning {tuning} n={n} bits={bits}: {'sorted' if good else 'WRONG'} ({passes} sweeps)", flush=True)
skew = (rng.integers(0, 4, 500_000, dtype=np.uint64) << np.uint64(40)) | rng.integers(0, 3, 500_000, dtype=np.uint64)
L.gsb_debug_set_tuning(tuning)
slo, _, passes = G.debug_sort_keys(skew, None, 64)
good = np.array_equal(slo, np.sort(skew))
ok &= good
print(f"tuning {tuning} skewed digits: {'sorted' if good else 'WRONG'} ({passes} sweeps)", flush=True)
L.gsb_debug_set_tuning(0)
for t in [0] + tunings:
    t0 = time.time()
    sweep_ms, sort_ms, sweeps = G.debug_sort_bench(198_333_373, 64, iters=3, tuning=t)
    print(f"tuning {t}: {sweep_ms:.3f} ms per sweep, {sort_ms:.2f} ms per sort ({sweeps} sweeps) [{time.time() - t0:.1f} s]", flush=True)
print("OK" if ok else "FAILED")
