#!/usr/bin/env python
"""bench.py -- read-likelihood evals/s and genes/s of the Psi_GP_MH hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload timeseries|static] [--genes G] [--stream device|host]

One *step* = one pass of the whole hot path over one batch of synthetic genes: for every
gene T pre-runs (100 MH steps each), the device-side jump-variance reduction, the joint
chain to Geweke convergence (<= 20000 steps) and the posterior summaries -- i.e. what
`get_psi` does around `Psi_GP_MH` (diceseq/utils/run_utils.py:106-131).

Workload (BASELINE.json configs[2], the one the north-star target is quoted on):
synthetic time series, 10k genes x 2-8 isoforms x 8 time points x 1k reads/gene/time PER
GPU (weak scaling: every rank samples its own 10k genes; genes are independent, there is
no data-path collective, only the final gather of a few counters).

Printed JSON (rank 0): `value` = read-likelihood evaluations per second with the packed
matrices already resident in HBM (device-timed, max over ranks); `e2e` = the same metric
through the public API (`PipelinedSampler.run_stream`, a stream of batches) from pinned host
buffers, every step's host->device copy of W and device->host read of its summaries inside the
timed region; `roofline` for the
dominant kernel (the main-run launch of dicegp_chain_kernel); `cpu_baseline` = the oracle
port of the reference timed on this box's host cores on a bounded gene sample.
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
}
MCMC = dict(M=20000, initial=1000, gap=100)          # diceseq --mcmc default (diceseq.py:80-83)


# ----------------------------------------------------------------------------------------
# synthetic genes (parallel generation; untimed set-up)
# ----------------------------------------------------------------------------------------
def _gen_gene(args):
    from diceseq_b200 import synth
    cfg, g = args
    return synth.gene_W(cfg, g)


def make_batch(cfg, gene_ids, pinned):
    from multiprocessing import get_context
    from diceseq_b200.packer import pack_genes
    nproc = min(32, os.cpu_count() or 1)
    with get_context("fork").Pool(nproc) as pool:
        genes = pool.map(_gen_gene, [(cfg, int(g)) for g in gene_ids], chunksize=32)
    return pack_genes(genes, pinned=pinned)


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
        nv, h = self.nvml
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._stop.wait(0.1)

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
    cfg, g, T = args
    R, L, P = synth.gene_matrices(cfg, g)
    X = np.arange(T)
    theta2 = float(synth.default_theta2(X))
    rng = np.random.RandomState(12345 + g)
    n_reads = sum(r.shape[0] for r in R)
    with np.errstate(all="ignore"):
        res, _var, _ = orc.prerun_and_main(R, L, P, X, 3.0, theta2, MCMC["M"], MCMC["initial"], MCMC["gap"],
                                           rng=rng, faithful_cost=True)
    m = res[6]
    steps_main = (m + 1) if len(res[3]) == m else MCMC["M"]
    # evals: main steps x all reads + T pre-runs x 100 steps x that time point's reads (+ initial evaluations)
    evals = (steps_main + 1) * n_reads + sum((100 + 1) * r.shape[0] for r in R)
    return evals, steps_main


def cpu_sample(cfg, T, n_genes, first_gene, cores):
    from multiprocessing import get_context
    t0 = time.perf_counter()
    with get_context("fork").Pool(cores) as pool:
        out = pool.map(_cpu_gene, [(cfg, first_gene + i, T) for i in range(n_genes)], chunksize=1)
    dt = time.perf_counter() - t0
    evals = sum(o[0] for o in out)
    return evals, n_genes, dt


def run_reference(args, wl):
    """--impl reference: the reference's CPU path (oracle port; the Python reference itself
    does not exist on the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_genes = max(2 * cores, 16)
    for w in range(args.warmup):
        cpu_sample(wl["config"], wl["T"], min(n_genes, cores), 900000 + w * n_genes, cores)
    evals = genes = 0
    dt = 0.0
    for k in range(args.steps):
        e, g, t = cpu_sample(wl["config"], wl["T"], n_genes, 1000000 + k * n_genes, cores)
        evals += e; genes += g; dt += t
    value = evals / dt
    line = {
        "impl": "reference", "metric": "read_likelihood_evals_per_s", "value": value, "unit": "evals/s",
        "genes_per_s": genes / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "mcmc": MCMC},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": "%d genes of the workload per step, multiprocessing.Pool(%d), oracle port with the "
                                   "reference's per-step inv/det/svd cost" % (n_genes, cores)},
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

    from diceseq_b200.batch import BatchSampler, PipelinedSampler

    G = args.genes or wl["genes"]
    T = wl["T"]
    cpu_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before any CUDA work in this process: the worker pool is forked
        cores = os.cpu_count() or 1
        n = max(4 * cores, 32)                      # ~15-20 s of CPU work on all cores
        e, g, dt = cpu_sample(wl["config"], T, n, 2000000, cores)
        cpu_line = {"value": e / dt, "unit": "evals/s", "genes_per_s": g / dt, "cores": cores, "kind": "port",
                    "sample": "%d genes of the same workload, multiprocessing.Pool(%d), oracle port with the "
                              "reference's per-step inv/det/svd cost, %.1f s" % (n, cores, dt)}
    gene_ids = np.arange(rank * G, (rank + 1) * G, dtype=np.int64)     # weak scaling: G genes per rank
    t0 = time.time()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    packed = make_batch(wl["config"], gene_ids, pinned=True)
    if rank == 0:
        print("[bench] rank0: %d genes generated+packed in %.1fs, W = %.2f GB (larger than L2: no flush needed)"
              % (G, time.time() - t0, packed.w_bytes / 1e9), file=sys.stderr)
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
        sampler.run(packed, seed=1000 + w, **kw)

    # ---- (1) resident: W already in HBM, device-timed -------------------------------------
    sampler.upload(packed)
    launches0 = sampler.ctx.counters()["kernel_launches"]
    evals = 0
    genes_done = 0
    main_ms = 0.0
    alg_bytes = 0
    barrier()
    if clocks.proc is None:
        clocks.start()
    clocks.mark()
    sampler.ctx.timer_start()
    for k in range(args.steps):
        res = sampler.sample(seed=k, **kw)
        if rank == 0:
            st_ = sampler.stats
            print("[bench] resident step %d: prerun %.1f ms (%d launches), main %.1f ms (%d launches), summary %.1f ms, "
                  "total sample() %.1f ms" % (k, st_["prerun_kernel_ms"], st_["prerun_launches"], st_["main_kernel_ms"],
                                               st_["main_launches"], st_.get("summary_ms", 0.0), st_.get("sample_ms", 0.0)),
                  file=sys.stderr)
            print("[bench]   host wall per stage (ms): " + ", ".join("%s %.1f" % kv for kv in st_.get("wall_ms", {}).items()),
                  file=sys.stderr)
        evals += sampler.stats["read_evals"]
        genes_done += G
        main_ms += sampler.stats["main_kernel_ms"]
        alg_bytes += algorithmic_bytes(packed, res, T)[0]
    resident_ms = sampler.ctx.timer_stop()
    barrier()
    clk = clocks.stop()
    launches = sampler.ctx.counters()["kernel_launches"] - launches0
    resident_ms = max_over_ranks(resident_ms)
    total_evals = sum_over_ranks(evals)
    total_genes = sum_over_ranks(genes_done)
    conv = float((res.state == 1).mean())
    m_med = float(np.median(res.m))

    # ---- (2) end to end: pinned host W -> device -> summaries on host ----------------------
    # A stream of batches through PipelinedSampler: every step copies its own 3 GB of W from pinned
    # host memory inside the timed region; from the second step on the copy runs on the idle
    # context while the previous batch is being sampled (the first copy is fully exposed).
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
        if rank == 0:
            print("[bench] e2e step: sample() %.1f ms, elapsed %.1f ms" % (st_["sample_ms"], (time.perf_counter() - t0) * 1e3),
                  file=sys.stderr)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3                            # wall clock: host work is part of e2e
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_total = sum_over_ranks(e2e_evals)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
    # DRAM bytes of the main-run launch set of ONE step, from the committed ncu launch list of this
    # command (dram__bytes_read.sum + dram__bytes_write.sum); only valid for the profiled configuration
    traffic = None
    try:
        if args.workload == "timeseries" and G == WORKLOADS["timeseries"]["genes"]:
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", "traffic_r01b.json")))["main_run_dram_bytes_per_step"])
    except (OSError, KeyError, ValueError):
        traffic = None

    line = {
        "metric": "read_likelihood_evals_per_s", "value": total_evals / (resident_ms * 1e-3), "unit": "evals/s",
        "genes_per_s": total_genes / (resident_ms * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": resident_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "genes_per_gpu": G, "mcmc": MCMC, "stream": args.stream,
                   "l2": "inputs larger than L2 (%.2f GB of W per GPU), no flush" % (packed.w_bytes / 1e9),
                   "converged_fraction": conv, "median_m": m_med},
        "clocks": clk,
        "e2e": {"value": e2e_total / (e2e_ms * 1e-3), "unit": "evals/s", "genes_per_s": total_genes / (e2e_ms * 1e-3),
                "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(packed.w_bytes + packed.len_iso.nbytes + packed.n_iso.nbytes + packed.n_reads.nbytes),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "traffic_unit": "bytes per step (main-run launch set, profiles/traffic_r01b.json)",
                     "traffic_gbs": (traffic / (main_ms / args.steps * 1e-3) / 1e9) if traffic and main_ms > 0 else None,
                     "traffic_frac": (traffic / (main_ms / args.steps * 1e-3) / 1e9 / peak) if traffic and main_ms > 0 and peak else None,
                     "main_run_ms_per_step": main_ms / args.steps,
                     "algorithmic_bytes_per_step": alg_bytes / args.steps,
                     "kernel": "dicegp_chain_kernel (main-run launch set: one launch per residency class, concurrent)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "note": "algorithmic bytes = sum over main-run MH steps of 8*C*N + stream + trace row; W is "
                             "re-read from shared memory / L2 every step, so achieved may exceed the HBM peak; "
                             "traffic_gbs = profiled DRAM bytes of the same launch set / live main-run time "
                             "(what the HBM actually delivers), traffic_frac = that / peak"},
    }
    if cpu_line is not None:
        line["cpu_baseline"] = cpu_line
    if rank == 0:
        emit_json_line(line)
    pipe.close()
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
    ap.add_argument("--genes", type=int, default=0, help="genes per GPU (default: the workload's 10000)")
    ap.add_argument("--stream", default="device", choices=["device", "host"])
    ap.add_argument("--trace-gb", type=float, default=64.0, help="trace pool budget per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
