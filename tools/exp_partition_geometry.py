import sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gossamer_b200 as G
import bench
for ms, tb in ((0, 0), (4096, 17), (4096, 16), (2048, 18)):
    G.debug_set_partition(ms, tb)
    sys.argv = ["bench.py", "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--no-other-configs"]
    import io, contextlib
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            bench.main()
    except SystemExit:
        pass
    l = json.loads(buf.getvalue().strip().splitlines()[-1])
    print(ms, tb, round(l["ms_per_step"], 3), {k: round(v, 3) for k, v in l["phases_ms_per_step"].items() if v}, l["parity"]["match"], flush=True)
