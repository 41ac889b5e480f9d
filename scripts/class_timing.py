"""Per-isoform-count timing of the main run (development aid): `n` genes of the time-series workload with a forced
isoform count, full pipeline, pass times from DICEGP_VERBOSE.
    class_timing.py C[,C..] n [env ...]      env = "KEY=VAL;KEY=VAL" applied for one more timed run each
`{C}` in an env string is replaced by the isoform count (e.g. "DICEGP_PLAN_K{C}=128,4")."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprocessing import Pool
from diceseq_b200 import synth
from diceseq_b200.batch import BatchSampler
from diceseq_b200.packer import pack_genes

Cs = [int(c) for c in sys.argv[1].split(",")]
n = int(sys.argv[2])
envs = sys.argv[3:]
def gen(a):
    return synth.gene_W("timeseries", a[1], C=a[0])
s = BatchSampler(0)
for C in Cs:
    with Pool(min(32, os.cpu_count())) as p:
        genes = p.map(gen, [(C, g) for g in range(n)], chunksize=16)
    packed = pack_genes(genes)
    s.upload(packed)
    s.sample(seed=1, summarize=False)                       # warm-up
    for env in [""] + envs:
        kv = [e.split("=", 1) for e in env.replace("{C}", str(C)).split(";") if e]
        for k, v in kv:
            os.environ[k] = v
        try:
            s.sample(seed=1, summarize=False)
            st = s.stats
            print("C=%d n=%d [%s]: prerun %.1f ms main %.1f ms steps %d rounds %d W-bytes %.1f GB (%.3e evals/s main)" % (
                C, n, env, st["prerun_kernel_ms"], st["main_kernel_ms"], st["main_steps"], st.get("main_rounds", 0),
                st.get("main_w_bytes", 0) / 1e9, st["main_read_evals"] / st["main_kernel_ms"] * 1e3), flush=True)
        except Exception as exc:
            print("C=%d [%s]: FAILED %r" % (C, env, exc), flush=True)
        for k, _ in kv:
            os.environ.pop(k, None)
