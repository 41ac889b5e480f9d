#!/usr/bin/env python
"""bench.py -- read-likelihood evals/s and genes/s of the Psi_GP_MH hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload timeseries|static|bias|human] [--genes G] [--stream device|host]

One *step* = one pass of the whole hot path over the synthetic gene set: for every gene T pre-runs (100 MH
steps each), the device-side jump-variance reduction, the joint chain to Geweke convergence (<= 20000
steps), the posterior summaries and -- with several GPUs -- the gather of the summaries on rank 0: what
`get_psi` does around `Psi_GP_MH` (diceseq/utils/run_utils.py:106-131) and what the reference's worker pool
does over genes (diceseq/diceseq.py:226-235).

Workload (BASELINE.json configs[2], the one the north-star target is quoted on): ONE synthetic time
series of 10k genes x 2-8 isoforms x 8 time points x 1k reads/gene/time for the WHOLE job.  With N GPUs the
gene set is split by read-count-balanced greedy (LPT) partitioning on C x reads (`partition.lpt_partition`);
there is no data-path collective, only the final gather: STRONG scaling (`"scaling": "strong"`).  The
weak-scaling figure (10k genes per GPU) is reported as the extra key `weak` when N > 1.

Printed JSON (rank 0): `value` = read-likelihood evaluations per second with the packed matrices already
resident in HBM (device-timed, max over ranks); `e2e` = the same metric through the public API
(`PipelinedSampler.run_stream`, a stream of batches) from pinned host buffers, every step's host->device
copy of W and device->host read of its summaries inside the timed region; `roofline` = the main-run launch
set of the chain kernels against the ceiling it sits closest to, with all three ceilings in
`roofline.ceilings` (HBM copy peak, L2 read bandwidth, fp64 FMA rate), from what the kernels counted
themselves in the timed run (`dicegp_work_counters`); `cpu_baseline` = the oracle port of the reference
timed on this box's host cores on a bounded gene sample; `other_configs` (N = 1) = the same figures for
BASELINE configs 2, 4 and 5 (static, bias-weighted, human-scale sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "timeseries": dict(config="timeseries", T=8, genes=10000,
                       label="synthetic time series 10k genes x 2-8 isoforms x 8 time points x 1k reads/gene/time"),
    "static": dict(config="static", T=1, genes=10000,
                   label="single-time-point static mode, synthetic 10k genes x 3 isoforms x 1k reads"),
    "bias": dict(config="bias", T=8, genes=20000,
                 label="bias-corrected run: per-read bias weights folded into R, 20k genes x 2-8 isoforms x 8 time points x 1k reads"),
    "human": dict(config="human", T=4, genes=3000,
                  label="human-scale: 3000-gene sample of the 60k-gene set, up to 20 isoforms, heavy-tailed 10-1M reads/gene/time, 4 time points"),
}
# measured ceilings of this pool's B200 (scripts/ubench/*.cu, numbers in profiles/ubench_r02.json)
UBENCH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ubench_r02.json")
MCMC = dict(M=20000, initial=1000, gap=100)          # diceseq --mcmc default (diceseq.py:80-83)


# ----------------------------------------------------------------------------------------
# synthetic genes (parallel generation; untimed set-up)
# ----------------------------------------------------------------------------------------
def _gen_gene(args):
    from diceseq_b200 import synth
    from diceseq_b200.packer import pack_gene
    cfg, g = args
    return pack_gene(*synth.gene_W(cfg, g))               # generated AND packed in the worker


def make_batch(cfg, gene_ids, pinned):
    from multiprocessing import get_context
    from diceseq_b200.packer import pack_packed
    nproc = max(2, min(32, os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    with get_context("fork").Pool(nproc) as pool:
        items = pool.map(_gen_gene, [(cfg, int(g)) for g in gene_ids], chunksize=32)
    return pack_packed(items, pinned=pinned)


# ----------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every 100 ms DURING the timed region.  NVML in-process
    (nvidia_ml_py): a polling `nvidia-smi -lms` child was seen to stall this process's CUDA calls
    for tens of ms at a time; nvidia-smi stays as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml = None
        self._stop = threading.Event()

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index()))
            self.proc = "nvml"
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def mark(self):
        """The timed region starts here: earlier samples are dropped."""
        self.rows = []

    def _poll_nvml(self):
        # The constant (max clock) is asked once; a poll is two cheap queries every 200 ms.  (NVML queries take driver
        # locks: polled every 100 ms with three queries each, single host-side CUDA calls of the timed steps were
        # seen to stall for 50-80 ms once or twice per run.)
        nv, h = self.nvml
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._stop.wait(float(os.environ.get("BENCH_CLOCK_PERIOD", "0.2")))

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.nvml is not None:
            self._stop.set()
        else:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------------------
# the reference arm / cpu baseline: oracle port of the reference on host cores
# ----------------------------------------------------------------------------------------
def _cpu_gene(args):
    """One gene through the reference's sampler logic (pre-runs + joint run), oracle port
    with the reference's per-step cost (inverse/determinant/svd redone every step)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from diceseq_b200 import synth
    from oracle import psi_gp_mh_oracle as orc
    cfg, g, T, mcmc = args
    R, L, P = synth.gene_matrices(cfg, g)
    X = np.arange(T)
    theta2 = float(synth.default_theta2(X))
    rng = np.random.RandomState(12345 + g)
    n_reads = sum(r.shape[0] for r in R)
    with np.errstate(all="ignore"):
        res, _var, _ = orc.prerun_and_main(R, L, P, X, 3.0, theta2, mcmc["M"], mcmc["initial"], mcmc["gap"],
                                           rng=rng, faithful_cost=True)
    m = res[6]
    steps_main = (m + 1) if len(res[3]) == m else mcmc["M"]
    # evals: main steps x all reads + T pre-runs x 100 steps x that time point's reads (+ initial evaluations)
    evals = (steps_main + 1) * n_reads + sum((100 + 1) * r.shape[0] for r in R)
    return evals, steps_main


class CpuPool:
    """ONE fork pool for a whole arm, fed through imap_unordered with chunksize 1: the cores stay busy until
    the last few genes of a sample (a fresh pool per step with two genes per core left half of them idle
    behind the slowest gene)."""

    def __init__(self, cores):
        from multiprocessing import get_context
        self.cores = cores
        self.pool = get_context("fork").Pool(cores)

    def sample(self, cfg, T, gene_ids, mcmc=MCMC):
        t0 = time.perf_counter()
        out = list(self.pool.imap_unordered(_cpu_gene, [(cfg, int(g), T, mcmc) for g in gene_ids], chunksize=1))
        dt = time.perf_counter() - t0
        return sum(o[0] for o in out), len(out), dt

    def close(self):
        self.pool.close()
        self.pool.join()


CPU_FIRST_GENE = 2000000            # both arms time the same genes of the workload on the host


def cpu_mcmc(workload):
    """The human-scale genes reach 1M reads per time point: 20000 reference steps of one such gene take
    minutes on a core, so its CPU sample is cut at 1000 joint steps (evals/s does not depend on the length
    of the chain)."""
    return dict(M=1000, initial=1000, gap=100) if workload == "human" else MCMC


def result_config(wl):
    """`config` of the JSON line: the same keys in both arms."""
    return {"workload": wl["label"], "mcmc": MCMC}


def run_reference(args, wl):
    """--impl reference: the reference's CPU path (oracle port; the Python reference itself does not exist on
    the GPU box) on all host cores.  A step = a bounded sample of the workload's genes (3 per core); all
    steps go through one pool back to back, so the cores are idle only behind the very last genes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 3 * cores
    mcmc = cpu_mcmc(args.workload)
    pool = CpuPool(cores)
    if args.warmup > 0:
        pool.sample(wl["config"], wl["T"], range(CPU_FIRST_GENE - cores, CPU_FIRST_GENE), mcmc)
    evals, genes, dt = pool.sample(wl["config"], wl["T"], range(CPU_FIRST_GENE, CPU_FIRST_GENE + per_step * args.steps), mcmc)
    pool.close()
    value = evals / dt
    line = {
        "impl": "reference", "metric": "read_likelihood_evals_per_s", "value": value, "unit": "evals/s",
        "genes_per_s": genes / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": result_config(wl),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d genes of the workload per step (%d in all, one multiprocessing.Pool(%d), "
                                   "imap_unordered), oracle port with the reference's per-step inv/det/svd cost%s"
                                   % (per_step, per_step * args.steps, cores,
                                      "" if mcmc is MCMC else ", joint chains cut at %d steps" % mcmc["M"])},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json_line(line)


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def algorithmic_bytes(packed, res, T):
    """per MH step per gene 8*C*N (W) + 8*((C-1)T+1) (stream) + 8*(2CT+(C-1)+2)
    (trace), summed over the main-run steps of every gene."""
    C = packed.n_iso.astype(np.int64)
    N = packed.reads_per_gene().astype(np.int64)
    steps = np.where(res.state == 1, res.m + 1, res.n_rows).astype(np.int64)
    per_step = 8 * C * N + 8 * ((C - 1) * T + 1) + 8 * (2 * C * T + (C - 1) + 2)
    return int((steps * per_step).sum()), int((steps * N).sum())


def _shape(args):
    from diceseq_b200 import synth
    C, n_t = synth.gene_shape(*args)
    return C, int(np.sum(n_t))


def partition_genes(cfg, G, world):
    """The north-star split: LPT-greedy on C x reads over the ranks (every rank computes the same bins)."""
    from multiprocessing import get_context
    from diceseq_b200 import partition
    if world == 1:
        return [np.arange(G, dtype=np.int64)]
    with get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        shapes = pool.map(_shape, [(cfg, g) for g in range(G)], chunksize=256)
    costs = partition.gene_cost([s_[0] for s_ in shapes], [s_[1] for s_ in shapes])
    return partition.lpt_partition(costs, world)


def load_ceilings(clk_mhz):
    """Roofline denominators: HBM copy peak from MEASURED_PEAKS.json (driver-written), L2 read bandwidth and fp64
    FMA issue rate from this repo's microbenchmarks (profiles/ubench_r02.json), the latter at the SM clock
    seen during the timed region."""
    peaks, ub = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    try:
        ub = json.load(open(UBENCH))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    l2 = float(ub.get("l2_read_gbs", 0.0)) or None
    per_clk = float(ub.get("dfma_lane_ops_per_clk_per_sm", 0.0)) or None
    sms = int(ub.get("sm_count", 148))
    fp64 = per_clk * sms * (clk_mhz or float(peaks.get("sm_max_mhz", 1965.0))) * 1e6 if per_clk else None
    src = {"hbm": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
           "l2": "profiles/ubench_r02.json l2_read_gbs (scripts/ubench/l2_bw.cu, 64 MB working set)",
           "fp64": "profiles/ubench_r02.json dfma_lane_ops_per_clk_per_sm (scripts/ubench/dp_latency.cu) x SMs x SM clock in the timed region"}
    return hbm, l2, fp64, src


def roofline_block(stats_sum, steps, clk, traffic_profile=None):
    """Main-run launch set of the chain kernels: what they streamed and multiplied (their own counters, this
    run) over their device time, against the three ceilings; the top-level figures are those of the ceiling
    the kernels sit closest to."""
    ms = stats_sum["main_kernel_ms"]
    if ms <= 0:
        return None
    sec = ms * 1e-3
    hbm, l2, fp64, src = load_ceilings(clk.get("sm_mhz"))
    streamed = stats_sum["main_w_bytes"] / sec / 1e9               # GB/s into shared memory (L2 hits + HBM)
    fma = stats_sum["main_fma"] / sec                              # fp64 FMA lane-ops per second
    ceil = {
        "l2": {"achieved": streamed, "peak": l2, "unit": "GB/s", "frac": streamed / l2 if l2 else None,
               "what": "bytes of W the kernels brought into shared memory (TMA / cp.async; L2 hits included)"},
        "fp64": {"achieved": fma / 1e12, "peak": fp64 / 1e12 if fp64 else None, "unit": "TFMA/s",
                 "frac": fma / fp64 if fp64 else None,
                 "what": "fused multiply-adds of every evaluated candidate (speculated ones included)"},
        "hbm": {"achieved": streamed, "peak": hbm, "unit": "GB/s", "frac": streamed / hbm,
                "what": "the same streamed bytes against the HBM copy peak: an upper bound of the DRAM traffic (part is served by L2)"},
    }
    if traffic_profile:
        t = traffic_profile["main_run_dram_bytes_per_step"]
        ceil["hbm"]["profile_traffic"] = {"dram_bytes_per_step": t, "gbs": t / (sec / steps) / 1e9,
                                          "frac": t / (sec / steps) / 1e9 / hbm, "source": traffic_profile.get("source")}
    # top level: the streamed bytes against the measured HBM copy peak (north_star's "fraction of the memory
    # roofline"); `closest` names the ceiling with the largest fraction
    fr = {k: v["frac"] for k, v in ceil.items() if v["frac"] is not None}
    top = ceil["hbm"]
    return {"bound": "hbm", "closest": max(fr, key=fr.get),
            "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
            "traffic": traffic_profile["main_run_dram_bytes_per_step"] if traffic_profile else None,
            "ceilings": ceil, "peak_source": src,
            "main_run_ms_per_step": ms / steps, "rounds_per_step": stats_sum["main_rounds"] / steps,
            "committed_steps_per_round": stats_sum["main_steps"] / max(stats_sum["main_rounds"], 1),
            "kernel": "dicegp_chain_kernel + dicegp_wide_kernel (main-run launch set: one launch per residency class, concurrent)",
            "note": "achieved = counters the kernels incremented in the timed run (dicegp_work_counters) / their device "
                    "time; `traffic` = DRAM bytes of the same launch set from the committed ncu launch list (profile)"}


def sample_workload(sampler, packed, kw, steps, warmup, on_step=None):
    """`steps` timed passes of BatchSampler.sample over resident genes; returns (device ms, summed stats, last result)."""
    for w in range(warmup):
        sampler.sample(seed=1000 + w, **kw)
    tot = {}
    sampler.ctx.timer_start()
    res = None
    for k in range(steps):
        res = sampler.sample(seed=k, **kw)
        for key, v in sampler.stats.items():
            if isinstance(v, (int, float)):
                tot[key] = tot.get(key, 0) + v
        if on_step:
            on_step(k, res)
    return sampler.ctx.timer_stop(), tot, res


def other_config_line(name, sampler, cpu_pool, steps=2, warmup=1):
    """evals/s, genes/s, converged fraction and roofline of one more BASELINE configuration (N = 1), with its
    own CPU sample beside it."""
    wl = WORKLOADS[name]
    t0 = time.time()
    ids = np.arange(wl["genes"], dtype=np.int64)
    packed = make_batch(wl["config"], ids, pinned=False)
    gen_s = time.time() - t0
    sampler.upload(packed)
    kw = dict(stream="device", gene_ids=ids, **MCMC)
    ms, tot, res = sample_workload(sampler, packed, kw, steps, warmup)
    clk = {"sm_mhz": None}
    out = {"workload": wl["label"], "genes": wl["genes"], "w_gb": packed.w_bytes / 1e9, "steps": steps,
           "value": tot["read_evals"] / (ms * 1e-3), "unit": "evals/s", "genes_per_s": wl["genes"] * steps / (ms * 1e-3),
           "ms_per_step": ms / steps, "converged_fraction": float((res.state == 1).mean()), "median_m": float(np.median(res.m)),
           "main_run_ms_per_step": tot["main_kernel_ms"] / steps, "prerun_ms_per_step": tot["prerun_kernel_ms"] / steps,
           "roofline": roofline_block(tot, steps, clk), "generate_s": gen_s}
    if cpu_pool is not None:
        mcmc = cpu_mcmc(name)
        n = 2 * cpu_pool.cores
        e, g, dt = cpu_pool.sample(wl["config"], wl["T"], range(CPU_FIRST_GENE, CPU_FIRST_GENE + n), mcmc)
        out["cpu_baseline"] = {"value": e / dt, "unit": "evals/s", "genes_per_s": g / dt, "cores": cpu_pool.cores, "kind": "port",
                               "sample": "%d genes, one Pool(%d), %.1f s%s" % (n, cpu_pool.cores, dt, "" if mcmc is MCMC else ", joint chains cut at %d steps" % mcmc["M"])}
    return out


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from diceseq_b200 import partition
    from diceseq_b200.batch import BatchSampler, PipelinedSampler

    G = args.genes or wl["genes"]                       # genes of the WHOLE job
    T = wl["T"]
    cpu_line = None
    cpu_pool = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before any CUDA work in this process: the worker pool is forked
        cores = os.cpu_count() or 1
        cpu_pool = CpuPool(cores)
        n = 6 * cores                                   # ~15-20 s of CPU work on all cores
        mcmc = cpu_mcmc(args.workload)
        e, g, dt = cpu_pool.sample(wl["config"], T, range(CPU_FIRST_GENE, CPU_FIRST_GENE + n), mcmc)
        cpu_line = {"value": e / dt, "unit": "evals/s", "genes_per_s": g / dt, "cores": cores, "kind": "port",
                    "sample": "%d genes of the same workload, one multiprocessing.Pool(%d) (imap_unordered), oracle port "
                              "with the reference's per-step inv/det/svd cost, %.1f s" % (n, cores, dt)}
    t0 = time.time()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bins = partition_genes(wl["config"], G, world)      # strong scaling: ONE gene set, LPT over the ranks
    gene_ids = bins[rank]
    packed = make_batch(wl["config"], gene_ids, pinned=True)
    if rank == 0:
        print("[bench] rank0: %d of %d genes generated+packed in %.1fs, W = %.2f GB per rank"
              % (len(gene_ids), G, time.time() - t0, packed.w_bytes / 1e9), file=sys.stderr)
    # the packed batch and the generator's leftovers are millions of long-lived Python objects: move
    # them out of the collector's reach, or a full collection (~50 ms) lands inside a timed step
    import gc
    gc.collect()
    gc.freeze()
    sampler = BatchSampler(local)
    if args.trace_gb > 0:
        sampler.ctx.set_trace_budget(int(args.trace_gb * 1e9))
    kw = dict(stream=args.stream, gene_ids=gene_ids, **MCMC)

    # warm-up: full pipeline, different seeds
    # (nvidia-smi's start-up -- NVML attaching to the GPUs -- stalls CUDA calls of this process for
    # ~70 ms, so the clock sampler is started before the last warm-up step and only the rows it
    # reports inside the timed region are kept)
    clocks = ClockSampler(local)
    for w in range(args.warmup):
        if w == args.warmup - 1:
            clocks.start()
        res = sampler.run(packed, seed=1000 + w, **kw)
        partition.gather_batch(res, gene_ids, device=local)

    # ---- (1) resident: W already in HBM, device-timed, gather of the summaries included ----
    sampler.upload(packed)
    launches0 = sampler.ctx.counters()["kernel_launches"]
    barrier()
    if clocks.proc is None:
        clocks.start()
    clocks.mark()
    gathered = [None]
    step_ms = []                                        # this rank's sample() wall per resident step (diagnostic)

    def on_step(k, res):
        step_ms.append(round(sampler.stats.get("sample_ms", 0.0), 1))
        gathered[0] = partition.gather_batch(res, gene_ids, device=local)    # the final gather (rank 0 gets all genes)
        if rank == 0:
            st_ = sampler.stats
            print("[bench] resident step %d: prerun %.1f ms (%d launches), main %.1f ms (%d launches), summary %.1f ms, "
                  "total sample() %.1f ms" % (k, st_["prerun_kernel_ms"], st_["prerun_launches"], st_["main_kernel_ms"],
                                               st_["main_launches"], st_.get("summary_ms", 0.0), st_.get("sample_ms", 0.0)),
                  file=sys.stderr)
            print("[bench]   host wall per stage (ms): " + ", ".join("%s %.1f" % kv for kv in st_.get("wall_ms", {}).items()),
                  file=sys.stderr)

    resident_ms, tot, res = sample_workload(sampler, packed, kw, args.steps, 0, on_step)
    barrier()
    clk = clocks.stop()
    launches = sampler.ctx.counters()["kernel_launches"] - launches0
    resident_ms = max_over_ranks(resident_ms)
    total_evals = sum_over_ranks(tot["read_evals"])
    total_genes = G * args.steps
    if rank == 0 and world > 1:
        assert gathered[0] is not None and len(gathered[0]["gene_ids"]) == G, "the gather did not return every gene"
    rank_ms = [step_ms]
    if world > 1:                                       # outside the timed region: every rank's own step times
        rank_ms = [None] * world
        dist.all_gather_object(rank_ms, step_ms)
    conv = sum_over_ranks(float((res.state == 1).sum())) / G
    m_med = float(np.median(gathered[0]["m"])) if rank == 0 else 0.0
    main_ms = max_over_ranks(tot["main_kernel_ms"])
    tot_all = {k: sum_over_ranks(float(tot.get(k, 0))) for k in ("main_w_bytes", "main_fma", "main_rounds", "main_steps")}
    tot_all["main_kernel_ms"] = main_ms
    if world > 1:                                       # per-GPU rates against per-GPU ceilings
        for k in ("main_w_bytes", "main_fma"):
            tot_all[k] /= world

    # ---- (2) end to end: pinned host W -> device -> summaries on host (+ gather) ----------------
    # A stream of batches through PipelinedSampler: every step copies its own W from pinned host memory inside
    # the timed region; from the second step on the copy runs on the idle context while the previous batch is
    # being sampled (the first copy is fully exposed).
    sampler.close()
    del sampler
    pipe = PipelinedSampler(local, int(min(args.trace_gb, 56.0) * 1e9) if args.trace_gb > 0 else 0)
    for _res, _st in pipe.run_stream((packed, dict(seed=2000 + w, **kw)) for w in range(2)):
        pass                                                             # both contexts allocate their buffers
    e2e_evals = 0
    d2h = 0
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for res, st_ in pipe.run_stream((packed, dict(seed=k, **kw)) for k in range(args.steps)):
        e2e_evals += st_["read_evals"]
        d2h = res.d2h_bytes()
        partition.gather_batch(res, gene_ids, device=local)
        if rank == 0:
            print("[bench] e2e step: sample() %.1f ms, elapsed %.1f ms" % (st_["sample_ms"], (time.perf_counter() - t0) * 1e3),
                  file=sys.stderr)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3                            # wall clock: host work is part of e2e
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_total = sum_over_ranks(e2e_evals)
    h2d = sum_over_ranks(float(packed.w_bytes + packed.len_iso.nbytes + packed.n_iso.nbytes + packed.n_reads.nbytes))
    pipe.close()

    # ---- (3) N > 1: the weak-scaling figure (10k genes per GPU) as an extra key ------------------
    weak = None
    if world > 1 and not args.no_weak:
        ids_w = np.arange(rank * G, (rank + 1) * G, dtype=np.int64)
        packed_w = make_batch(wl["config"], ids_w, pinned=False)
        s2 = BatchSampler(local)
        if args.trace_gb > 0:
            s2.ctx.set_trace_budget(int(args.trace_gb * 1e9))
        s2.upload(packed_w)
        barrier()
        ms_w, tot_w, _ = sample_workload(s2, packed_w, dict(stream=args.stream, gene_ids=ids_w, **MCMC), min(args.steps, 3), 1)
        barrier()
        ms_w = max_over_ranks(ms_w)
        ev_w = sum_over_ranks(tot_w["read_evals"])
        weak = {"value": ev_w / (ms_w * 1e-3), "unit": "evals/s", "genes_per_gpu": G, "genes_per_s": G * world * min(args.steps, 3) / (ms_w * 1e-3),
                "ms_per_step": ms_w / min(args.steps, 3), "steps": min(args.steps, 3)}
        s2.close()

    traffic = None
    try:
        if args.workload == "timeseries" and G == WORKLOADS["timeseries"]["genes"] and world == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
    except (OSError, KeyError, ValueError):
        traffic = None

    line = {
        "metric": "read_likelihood_evals_per_s", "value": total_evals / (resident_ms * 1e-3), "unit": "evals/s",
        "genes_per_s": total_genes / (resident_ms * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": resident_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": result_config(wl),
        "details": {"genes": G, "genes_on_rank0": int(len(gene_ids)), "partition": "LPT greedy on C x reads, final gather of the summaries on rank 0" if world > 1 else "single GPU",
                    "stream": args.stream, "l2": "inputs larger than L2 (%.2f GB of W on rank 0), no flush" % (packed.w_bytes / 1e9),
                    "converged_fraction": conv, "median_m": m_med,
                    "rank_sample_ms": rank_ms},
        "clocks": clk,
        "e2e": {"value": e2e_total / (e2e_ms * 1e-3), "unit": "evals/s", "genes_per_s": total_genes / (e2e_ms * 1e-3),
                "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(sum_over_ranks(float(d2h)))},
        "gpu_launches": int(sum_over_ranks(float(launches))),
        "roofline": roofline_block(tot_all, args.steps, clk, traffic),
    }
    if weak is not None:
        line["weak"] = weak
    if cpu_line is not None:
        line["cpu_baseline"] = cpu_line
    if rank == 0 and world == 1 and not args.no_other_configs:
        s3 = BatchSampler(local)
        if args.trace_gb > 0:
            s3.ctx.set_trace_budget(int(args.trace_gb * 1e9))
        line["other_configs"] = []
        for name in ("static", "bias", "human"):
            if name == args.workload:
                continue
            try:
                oc = other_config_line(name, s3, cpu_pool)
                line["other_configs"].append(oc)
                print("[bench] %s: %.3e evals/s, %.0f genes/s, %.1f ms/step, converged %.3f" % (
                    name, oc["value"], oc["genes_per_s"], oc["ms_per_step"], oc["converged_fraction"]), file=sys.stderr)
            except Exception as exc:                     # a side line must not lose the headline
                line["other_configs"].append({"workload": WORKLOADS[name]["label"], "error": repr(exc)})
        s3.close()
    if cpu_pool is not None:
        cpu_pool.close()
    if rank == 0:
        emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints its
    version banner on the first communicator), so everything else that targets fd 1 is sent to
    stderr and the result line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_json_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="timeseries", choices=sorted(WORKLOADS))
    ap.add_argument("--genes", type=int, default=0, help="genes of the whole job (default: the workload's)")
    ap.add_argument("--stream", default="device", choices=["device", "host"])
    ap.add_argument("--trace-gb", type=float, default=64.0, help="trace pool budget per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the static / bias / human side lines (N = 1)")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling extra key (N > 1)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
