import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gossamer_b200 as G
import oracle_py as O
import simreads_py as S

g = S.genome(20000, 42)
text = bytes(S.reads_fastq(g, 100, 3000, err=0.01, seed=43))
for k, m in ((25, 1), (31, 2), (55, 1)):
    want, ost = O.build_graph([(text, O.FASTQ)], k, min_count=m)
    sink, counts, stats = G.build_graph([(text, G.FASTQ)], k, min_count=m)
    assert sink.as_bytes() == want.files(), k
fa = (">g\n" + "\n".join(bytes(g[i:i + 60]).decode() for i in range(0, 20000, 60)) + "\n").encode()
want, _ = O.build_kmer_set([(fa, O.FASTA)], 33)
sink, _, _ = G.build_kmer_set([(fa, G.FASTA)], 33)
assert sink.as_bytes() == want.files()
b = G.Builder(G.GRAPH, 31, min_count=2, max_batch_keys=200000)
for i in range(0, 3000, 500):
    b.push(bytes(S.reads_fastq(g, 100, 500, err=0.01, seed=50 + i, first_idx=i)), G.FASTQ)
b.finish()
b.emit("g", G.MemorySink())
b.close()
print("sanitizer run complete")
