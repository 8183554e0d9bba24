import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gossamer_b200 as G, simreads_py as S
g = S.genome(5_000_000, 42)
text = S.reads_fastq(g, 150, 1_666_667, err=0.01, seed=43)
b = G.Builder(G.GRAPH, 31, min_count=1)
for it in range(3):
    b.reset()
    b.push(text, G.FASTQ)
    c = b.finish()
    b.emit("g", None)
    st = b.stats()
    print(it, c.n_kept, "emit ms", round(st.ms_emit, 2), "bytes_out", st.bytes_out, flush=True)
