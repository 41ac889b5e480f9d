"""`diceseq` command line, drop-in for `diceseq/diceseq.py`.

Same flags, same `.dice` / `.sample.gz` formats, same transcript order (the reference sorts
its output back into annotation order, `run_utils.py:39-46`).  The work is restructured into
three stages: (1) host: read->isoform matrices of every gene (optionally over a process
pool, `-p`), (2) GPU: all multi-isoform genes sampled in batches by `BatchSampler`
(pre-runs, jump variance, joint chains, summaries on the device), (3) host: formatting.

Differences, all deliberate: single-isoform genes get psi = 1 rows instead of crashing the
worker; `.sample.gz` is written in text mode (Appendix B #11);
`--thetas x learn` runs host-driven chains, one gene at a time, with the likelihood on the GPU
(theta2_mode.py) and -- like the reference's sampler -- only handles two-isoform genes;
genes beyond the library's shape limits (isoforms, isoforms x time points) are named and skipped;
extra flags `--device`, `--gpus` (shard the genes over the GPUs of the box), `--seed`, `--batch`.

    python -m diceseq_b200.cli -a anno.gtf -s t1.bam---t2.bam---t3.bam -t 1.5,2.5,5 -o out/joint --add_premRNA
"""
from __future__ import annotations

import gzip
import os
import sys
import time
from optparse import OptionGroup, OptionParser

import numpy as np

from .annotation import loadgene
from .bam import load_samfile
from .run_utils import format_dice_lines, format_sample_lines, gene_inputs, single_isoform_result


def build_parser():
    parser = OptionParser()
    parser.add_option("--anno_file", "-a", dest="anno_file", default=None,
                      help="Annotation file for genes and transcripts in GTF or GFF3")
    parser.add_option("--sam_list", "-s", dest="sam_list", default=None,
                      help=("Sorted and indexed bam/sam files, use ',' for replicates "
                            "and '---' for time points, e.g., T1_rep1.bam,T1_rep2.bam---T2.bam"))
    parser.add_option("--time_seq", "-t", dest="time_seq", default=None,
                      help="The time for the input samples [Default: 0,1,2,...]")
    parser.add_option("--out_file", "-o", dest="out_file", default="output",
                      help="Prefix of the output files with full path")
    group = OptionGroup(parser, "Optional arguments")
    group.add_option("--nproc", "-p", type="int", dest="nproc", default="4",
                     help="Number of subprocesses for building the read matrices [default: %default]")
    group.add_option("--add_premRNA", action="store_true", dest="add_premRNA", default=False,
                     help="Add the pre-mRNA as a transcript")
    group.add_option("--fLen", type="float", nargs=2, dest="frag_leng", default=[None, None],
                     help="Two arguments for fragment length: mean and standard diveation, default: auto-detected")
    group.add_option("--bias", nargs=3, dest="bias_args", default=["unif", "None", "None"],
                     help="Three argments for bias correction: BIAS_MODE,REF_FILE,BIAS_FILE(s) [default: unif None None]")
    group.add_option("--thetas", nargs=2, dest="thetas", default=[3, "None"],
                     help="Two arguments for hyperparameters in GP model: theta1,theta2. default: [3 None], "
                          "where theta2 covers 1/3 duration.")
    group.add_option("--mcmc", type="int", nargs=4, dest="mcmc_run", default=[0, 20000, 1000, 100],
                     help="Four arguments for in MCMC iterations: save_sample,max_run,min_run,gap_run. "
                          "[default: 0 20000 1000 100]")
    group.add_option("--device", type="int", dest="device", default=0, help="GPU index (first GPU with --gpus) [default: %default]")
    group.add_option("--gpus", dest="gpus", default="1",
                     help="GPUs of this box to shard the genes over (read-count-balanced greedy partition): a count N "
                          "(devices --device .. --device+N-1) or a comma-separated list of device indices [default: %default]")
    group.add_option("--seed", type="int", dest="seed", default=0, help="seed of the device random stream")
    group.add_option("--batch", type="int", dest="batch", default=20000, help="genes per GPU batch")
    parser.add_option_group(group)
    return parser


def _inputs_worker(args):
    """Matrices of one gene; multi-isoform genes come back cleaned (model_GP.py:250-264) and packed in the device
    layout (packer.pack_gene), so the process that feeds the GPU only concatenates."""
    gene, sam_list, FLmean, FLstd, bias_mode, ref_file, bias_file, keep_raw = args
    R_all, len_all, prob_all, count = gene_inputs(gene, sam_list, FLmean, FLstd, bias_mode, ref_file, bias_file)
    if gene.tranNum < 2:
        return None, count
    if keep_raw:                                                     # theta2-sampling mode works on the raw lists
        return (R_all, len_all, prob_all), count
    from .packer import clean_inputs, pack_gene
    clean_inputs(R_all, len_all, prob_all)
    return pack_gene([R * P for R, P in zip(R_all, prob_all)], len_all), count


def sample_genes(idx, gene_inputs_list, device, batch, sample_num, run_kw, progress=None):
    """Stage 2 on ONE GPU: the multi-isoform genes `idx` (annotation indices, used as the global gene ids of
    the random streams; `gene_inputs_list[k]` = the packed matrices of idx[k], packer.pack_gene) in batches of
    `batch` through BatchSampler (pre-runs, jump variance, joint chains,
    summaries on the device).  Returns ({idx: GeneResult}, {idx: Y samples (n, C, T)}).  Runs in the CLI
    process, or in one worker process per GPU with --gpus."""
    from .batch import BatchSampler, PipelinedSampler
    from .packer import pack_packed
    T = len(run_kw["X"])
    where = {i: k for k, i in enumerate(idx)}
    chunks = [idx[b0:b0 + batch] for b0 in range(0, len(idx), batch)]

    def batches():
        for ch in chunks:
            yield pack_packed([gene_inputs_list[where[i]] for i in ch], pinned=len(chunks) > 1), dict(run_kw, gene_ids=np.array(ch))

    # several batches: the copy of batch k+1 overlaps the sampling of batch k (two contexts)
    pipe = PipelinedSampler(device) if len(chunks) > 1 else None
    single = BatchSampler(device) if len(chunks) == 1 else None

    def run_all():
        if pipe is not None:
            for res, _st in pipe.run_stream(batches()):
                yield res, pipe.current
        elif single is not None:
            for packed, kw in batches():
                yield single.run(packed, **kw), single

    results, samples, done = {}, {}, 0
    for (res, sampler), ch in zip(run_all(), chunks):
        for k, i in enumerate(ch):
            results[i] = res[k]
            if sample_num > 0:
                m = res[k].m
                n_keep = min(int(0.5 * m), sample_num)
                if n_keep > 0:
                    _psi, Y, _pro, _lik = sampler.ctx.download_trace(k, n_keep, first_row=int(m / 2))
                    samples[i] = Y
                else:
                    samples[i] = np.zeros((0, int(res.n_iso[k]), T))
        done += len(ch)
        if progress:
            progress(done)
    for s_ in (pipe, single):
        if s_ is not None:
            s_.close()
    return results, samples


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    options, _args = parser.parse_args(argv)
    if len(argv) == 0:
        print("Welcome to diceseq!\n")
        print("use -h or --help for help on argument.")
        sys.exit(1)
    if options.anno_file is None:
        print("[DICEseq] Error: need --anno_file for annotation.")
        sys.exit(1)
    sys.stdout.write("\r[DICEseq] loading annotation file...")
    sys.stdout.flush()
    genes = loadgene(options.anno_file)
    sys.stdout.write("\r[DICEseq] loading annotation file... Done.\n")
    if options.sam_list is None:
        print("[DICEseq] Error: need --sam_list for aliged & indexed reads.")
        sys.exit(1)
    sam_list = [tp.split(",") for tp in options.sam_list.split("---")]
    TOTAL_READ = []
    for files in sam_list:
        cnt = 0.0
        for ss in files:
            if not os.path.isfile(ss):
                print("Error: No such file\n    -- %s" % ss)
                sys.exit(1)
            cnt += load_samfile(ss).mapped_total()                 # pysam.idxstats column 3 (diceseq.py:141-147)
        TOTAL_READ.append(cnt)

    FLmean, FLstd = options.frag_leng
    sample_num, Mmax, Mmin, Mgap = options.mcmc_run
    X = np.arange(len(sam_list)) if options.time_seq is None else np.array(options.time_seq.split(","), "float")
    theta1, theta2 = options.thetas
    theta1 = float(theta1)
    if theta2 is None or theta2 in ("None", "Auto", "auto"):
        theta2 = ((max(X) - min(X) + 0.1) / 3.0) ** 2                # diceseq.py:169
    elif theta2 == "learn":
        theta2 = None                                                # diceseq.py:170-171
    else:
        theta2 = max(0.00001, float(theta2))
    bias_mode, ref_file, bias_file = options.bias_args               # diceseq.py:175-190
    if bias_mode == "unif":
        ref_file = bias_file = None
    elif ref_file == "None":
        print("[DICEseq] No reference sequence, change to uniform mode.")
        ref_file = bias_file = None
        bias_mode = "unif"
    elif bias_file == "None":
        print("[DICEseq] No bias parameter file, change to uniform mode.")
        ref_file = bias_file = None
        bias_mode = "unif"
    else:
        if bias_mode not in ("end5", "end3", "both"):
            print("[DICEseq] Error: --bias mode must be unif, end5, end3 or both.")
            sys.exit(1)
        bias_file = bias_file.split("---")
        for f in [ref_file] + bias_file:
            if not os.path.isfile(f):
                print("Error: No such file\n    -- %s" % f)
                sys.exit(1)

    if options.add_premRNA:
        for g in genes:
            g.add_premRNA()
    T = len(X)
    start_time = time.time()
    print("[DICEseq] running diceseq for %d genes on GPU %d (matrices on %d cores)..." % (
        len(genes), options.device, options.nproc))

    # ---- stage 1: host matrices -------------------------------------------------------------
    jobs = [(g, sam_list, FLmean, FLstd, bias_mode, ref_file, bias_file, theta2 is None) for g in genes]
    if options.nproc <= 1 or len(genes) < 4:
        inputs = [_inputs_worker(j) for j in jobs]
    else:
        import multiprocessing
        with multiprocessing.get_context("fork").Pool(options.nproc) as pool:
            inputs = pool.map(_inputs_worker, jobs, chunksize=max(1, len(jobs) // (8 * options.nproc)))

    # ---- stage 2: GPU sampling of the multi-isoform genes -----------------------------------
    multi = [i for i, g in enumerate(genes) if g.tranNum > 1]
    results = {}
    samples = {}
    sample_theta2 = {}
    if theta2 is None:
        # theta2 sampled per gene (model_GP.py:317-326): the proposal stream depends on the chain
        # state, so these chains are host-driven, one gene after the other, with the likelihood on
        # the GPU.  The reference's sampler only survives this mode for two-isoform genes; its
        # worker pool drops the others silently (diceseq.py:230-233), here they are named.
        from .run_utils import sample_gene_theta2
        if options.seed is not None:
            np.random.seed(options.seed)
        dropped = []
        for n_done, i in enumerate(multi):
            R_all, len_all, prob_all = inputs[i][0]
            try:
                r = sample_gene_theta2(R_all, len_all, prob_all, X, theta1, Mmax, Mmin, Mgap, sample_num,
                                       device=options.device)
            except np.linalg.LinAlgError:
                dropped.append(genes[i].geneID)
                continue
            results[i], samples[i], sample_theta2[i] = r, r.Y_tail, r.theta2_mean
            sys.stdout.write("\r[DICEseq] %d / %d genes sampled in %.1f sec." % (n_done + 1, len(multi), time.time() - start_time))
            sys.stdout.flush()
        if dropped:
            print("\n[DICEseq] Warning: --thetas x learn only works for two-isoform genes (as in the reference); "
                  "no output for %d genes: %s" % (len(dropped), ",".join(dropped[:20]) + ("..." if len(dropped) > 20 else "")))
        multi = []
    # genes beyond the compiled shape limits of the library are named and skipped (the reference has no
    # such limit; the rest of the run must not be lost to them)
    from . import _lib
    lim = _lib.shape_limits()
    too_big = [i for i in multi if genes[i].tranNum > lim["max_isoforms"] or T > lim["max_times"]
               or genes[i].tranNum * T > lim["max_latent"]]
    if too_big:
        print("[DICEseq] Warning: %d genes exceed the sampler's shape limits (isoforms <= %d, isoforms x time points <= %d); "
              "no output for: %s" % (len(too_big), lim["max_isoforms"], lim["max_latent"],
                                     ",".join(genes[i].geneID for i in too_big[:20]) + ("..." if len(too_big) > 20 else "")))
        drop = set(too_big)
        multi = [i for i in multi if i not in drop]

    run_kw = dict(X=X, theta1=theta1, theta2=theta2, M=Mmax, initial=Mmin, gap=Mgap, seed=options.seed)
    devices = ([int(d) for d in str(options.gpus).split(",")] if "," in str(options.gpus)
               else [options.device + r for r in range(max(1, int(options.gpus)))])
    n_gpus = len(devices)
    if n_gpus == 1 or len(multi) < 2 * n_gpus:
        def progress(done):
            sys.stdout.write("\r[DICEseq] %d / %d genes sampled in %.1f sec." % (done, len(multi), time.time() - start_time))
            sys.stdout.flush()
        res_d, samp_d = sample_genes(multi, [inputs[i][0] for i in multi], devices[0], options.batch, sample_num, run_kw, progress)
        results.update(res_d)
        samples.update(samp_d)
    elif multi:
        # one process per GPU (the reference fans genes over a process pool, diceseq.py:226-235): LPT partition on
        # isoforms x reads, no communication but the final gather of the per-gene results
        import multiprocessing
        from .partition import gene_cost, lpt_partition
        cost = gene_cost([genes[i].tranNum for i in multi], [int(inputs[i][0][1].sum()) for i in multi])
        bins = lpt_partition(cost, n_gpus)
        ctx = multiprocessing.get_context("spawn")
        with ctx.Pool(n_gpus) as pool:
            jobs2 = [pool.apply_async(sample_genes, ([multi[k] for k in b], [inputs[multi[k]][0] for k in b],
                                                     devices[r], options.batch, sample_num, run_kw))
                     for r, b in enumerate(bins)]
            for r, j in enumerate(jobs2):
                res_d, samp_d = j.get()
                results.update(res_d)
                samples.update(samp_d)
                sys.stdout.write("\r[DICEseq] GPU %d done: %d / %d genes sampled in %.1f sec." % (
                    devices[r], len(results), len(multi), time.time() - start_time))
                sys.stdout.flush()

    # ---- stage 3: output, in annotation order -----------------------------------------------
    out_dir = os.path.dirname(options.out_file)
    if out_dir and not os.path.isdir(out_dir):
        os.makedirs(out_dir, exist_ok=True)
    with open(options.out_file + ".dice", "w") as fid1:
        headline = "tran_id\tgene_id\tlogLik\ttransLen"
        for i in range(T):
            _t = str(X[i])
            headline += "\tFPKM_T%s\tratio_T%s\tratio_lo_T%s\tratio_hi_T%s" % (_t, _t, _t, _t)
        fid1.write(headline + "\n")
        fid2 = None
        if sample_num > 0:
            fid2 = gzip.open(options.out_file + ".sample.gz", "wt")
            fid2.write("# MCMC samples for latent Y\n# @gene|transcripts|theta2\n# y_c1t1,y_c2t1;y_c1t2,y_c2t2;...\n")
        for i, g in enumerate(genes):
            count = inputs[i][1]
            if g.tranNum > 1:
                if i not in results:
                    continue                                         # dropped (theta2-sampling mode, shape limits)
                r = results[i]
                psi_mean, psi_95, lik_marg = r.psi_mean, r.psi_95, r.lik_marg
            else:
                psi_mean, psi_95, lik_marg = single_isoform_result(T)
            fid1.write(format_dice_lines(g, X, count, TOTAL_READ, psi_mean, psi_95, lik_marg))
            if fid2 is not None and g.tranNum > 1:
                fid2.write(format_sample_lines(g, sample_theta2.get(i, theta2), samples[i], g.tranNum))
        if fid2 is not None:
            fid2.close()
    print("\n[DICEseq] done: %s.dice (%d genes, %.1f sec)." % (options.out_file, len(genes), time.time() - start_time))


if __name__ == "__main__":
    main()
