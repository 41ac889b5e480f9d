"""Does a concurrent host->device upload on a second context slow sample() down?"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
def gen(g):
    return synth.gene_W("timeseries", g)
with Pool(32) as p:
    genes = p.map(gen, range(G), chunksize=16)
packed = pack_genes(genes, pinned=True)
a, b = BatchSampler(0), BatchSampler(0)
for s in (a, b):
    s.ctx.set_trace_budget(int(56e9))
    s.run(packed, seed=1)
for name, s, other in (("ctx a alone", a, None), ("ctx b alone", b, None), ("ctx a + upload on b", a, b), ("ctx b + upload on a", b, a),
                       ("ctx a + 3 uploads on b", a, (b, 3))):
    n_up = 1
    if isinstance(other, tuple):
        other, n_up = other
    th = None
    if other is not None:
        def up():
            for _ in range(n_up):
                t = time.perf_counter(); other.upload(packed); print("    upload took %.1f ms" % ((time.perf_counter() - t) * 1e3))
        th = threading.Thread(target=up); th.start()
    t = time.perf_counter()
    s.sample(seed=2)
    dt = (time.perf_counter() - t) * 1e3
    if th: th.join()
    print("%-24s sample() %.1f ms  wall stages %s" % (name, dt, {k: round(v, 1) for k, v in s.stats["wall_ms"].items()}))
