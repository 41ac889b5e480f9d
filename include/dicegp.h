/*
 * dicegp.h -- C ABI of the B200-native DICEseq GP/MH isoform sampler (libdicegp.so).
 *
 * This library replaces ONE function of the reference and nothing else:
 *     Psi_GP_MH(R_mat, len_isos, prob_isos, X, Ymean, var, theta1, theta2, M, initial,
 *               gap, randomS, ...)             diceseq/models/model_GP.py:188-395
 * together with the way its only caller drives it (T single-time-point pre-runs that set
 * the jump variance, then the joint run: diceseq/utils/run_utils.py:106-122).
 * The reference has no FFI for this path (pure Python); the entry points below are what
 * a ctypes binding of that function would bind -- see INTEGRATION.md for the stub.
 *
 * Vocabulary (the reference's): a *gene* has C isoforms and T time points; at time t it
 * has N_t reads and a read-by-isoform probability matrix W_t = R_t * prob_t
 * (model_GP.py:338).  A *chain* is one Metropolis-Hastings run over a contiguous range
 * of a gene's time points (the pre-runs are T chains of one time point, the main run
 * is one chain of all T).  A *step* is one MH iteration (model_GP.py:311-391).
 *
 * Conventions: every call returns 0 on success and a negative dicegp_status otherwise
 * (never throws); dicegp_last_error() gives the text.  The caller owns every host
 * buffer (pinned memory recommended); the library owns the device memory behind the
 * handle.  One handle per GPU; a handle is not thread-safe.  Per-chain numerical
 * trouble (log-likelihood of -inf / NaN) is reported in the chain status, not as a call
 * failure -- the reference does not raise there either.
 * There is no CPU fallback: without a CUDA device dicegp_create fails.
 */
#ifndef DICEGP_H
#define DICEGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DICEGP_ABI_VERSION 2

typedef struct dicegp_ctx dicegp_ctx;

enum dicegp_status {
    DICEGP_OK = 0,
    DICEGP_ERR_ARG = -1,     /* bad argument (text in dicegp_last_error)            */
    DICEGP_ERR_CUDA = -2,    /* a CUDA runtime call failed                           */
    DICEGP_ERR_STATE = -3,   /* call sequence violated (e.g. run before upload)      */
    DICEGP_ERR_NOMEM = -4,   /* device allocation failed / trace pool exhausted      */
    DICEGP_ERR_LIMIT = -5    /* shape beyond a compiled limit (C, T)                 */
};

/* Compiled shape limits (dicegp_limits fills them in). */
typedef struct dicegp_limits {
    int32_t max_isoforms;    /* C  */
    int32_t max_times;       /* T  */
    int32_t abi_version;
    int32_t sm_count;        /* of the device behind the handle */
    int32_t max_latent;      /* C * T of one gene: every latent value has its own thread in the state
                                warps of a chain's block; what dicegp_upload_genes accepts runs */
    int32_t reserved;
} dicegp_limits;

/* Chain state codes reported by dicegp_chain_status. */
enum dicegp_chain_state {
    DICEGP_CHAIN_RUNNING = 0,    /* more steps to do                                  */
    DICEGP_CHAIN_CONVERGED = 1,  /* Geweke rule met (model_GP.py:369-391)             */
    DICEGP_CHAIN_EXHAUSTED = 2,  /* reached M steps without converging                */
    DICEGP_CHAIN_NOMEM = 3       /* stopped: the trace pool is full (raise the budget or
                                    use fewer chains per batch); rows so far are kept   */
};

/* ------------------------------------------------------------------ lifetime ------- */
int dicegp_abi_version(void);
/* Replaces: process start of a multiprocessing worker (diceseq/diceseq.py:226). */
int dicegp_create(int device, dicegp_ctx **out);
int dicegp_destroy(dicegp_ctx *ctx);
const char *dicegp_last_error(const dicegp_ctx *ctx);   /* ctx may be NULL: global text */
int dicegp_get_limits(const dicegp_ctx *ctx, dicegp_limits *out);

/* ------------------------------------------------------------------ genes ---------- */
/*
 * Upload a batch of G genes that share T time points.  Replaces the R_mat / prob_isos /
 * len_isos arguments of Psi_GP_MH (model_GP.py:188) AFTER the host has done the input
 * cleaning of model_GP.py:250-264 and formed W = R * prob (model_GP.py:338).
 *
 *  n_iso[g]            C_g (>= 1)
 *  n_reads[g*T + t]    N_gt, rows of W_gt after cleaning (may be 0)
 *  W                   packed fp64: for g, for t: a (C_g x pad2(N_gt)) block stored
 *                      ISOFORM-MAJOR (row k holds isoform k's column of W_gt, padded with
 *                      zeros to an even count pad2(N) = N + (N & 1)).  Blocks are
 *                      concatenated without gaps; w_count = total doubles.
 *  len_iso             for g, for t: C_g effective lengths (model_GP.py:282,336)
 * A second upload replaces the first and discards all chains.
 */
int dicegp_upload_genes(dicegp_ctx *ctx, int32_t G, int32_t T, const int32_t *n_iso,
                        const int32_t *n_reads, const double *W, int64_t w_count,
                        const double *len_iso, int64_t len_count);

/* ------------------------------------------------------------------ chains --------- */
/*
 * One chain = one Psi_GP_MH call.  Host-computed constants (numpy on the reference
 * side) are passed in so that the device reproduces the reference's arithmetic:
 *  gene, t0, Tc      the chain covers time points [t0, t0+Tc) of `gene`
 *  M, initial, gap   model_GP.py:189 (max steps, first convergence check, check period)
 *  kinv              Tc x Tc row-major inverse of the GP kernel K (model_GP.py:83,287)
 *  prior_const       -(Tc*log(2 pi) + log det K)/2                    (model_GP.py:89)
 *  jump_scale[c]     var[c]*5/(Tc*C*theta1), cov_jmp = K*jump_scale   (model_GP.py:328)
 *  jump_const[c]     -(Tc*log(2 pi) + log det cov_jmp_c)/2            (model_GP.py:330)
 *  y0                C x Tc start value (the `Ymean` argument, model_GP.py:275)
 *  prop_factor       Tc x Tc row-major matrix A_K with  z @ A_K ~ N(0, K) (numpy's svd
 *                    recipe); only used by the device-generated stream, where the
 *                    proposal increment is sqrt(jump_scale[c]) * (z @ A_K).
 * Arrays of n_chains descriptors; the double payload is passed as one pool with
 * per-chain offsets (in doubles) to keep the ABI free of nested pointers.
 */
typedef struct dicegp_chain_desc {
    int32_t gene, t0, Tc;
    int32_t M, initial, gap;
    int64_t kinv_off;         /* Tc*Tc doubles    */
    int64_t prop_factor_off;  /* Tc*Tc doubles, or -1 if only host streams are used */
    int64_t jump_scale_off;   /* C-1 doubles      */
    int64_t jump_const_off;   /* C-1 doubles      */
    int64_t y0_off;           /* C*Tc doubles, or -1 for all zeros */
    double  prior_const;
} dicegp_chain_desc;

/*
 * Trace memory.  The reference allocates M rows per call up front (model_GP.py:295-310);
 * 10k genes x M=20000 would need >100 GB that way while most chains converge after
 * 1-2k steps.  Traces therefore live in a paged pool: every chain gets its first page
 * (<= 1024 rows) when it is created and grows page by page on the device.  The budget
 * (bytes; 0 = 60% of the free device memory) caps the pool; a chain that cannot get a
 * page stops with DICEGP_CHAIN_NOMEM.  Call before dicegp_set_chains.
 */
int dicegp_set_trace_budget(dicegp_ctx *ctx, int64_t bytes);

/* Discards existing chains, creates n_chains new ones (ids 0..n-1), allocates traces. */
int dicegp_set_chains(dicegp_ctx *ctx, int32_t n_chains, const dicegp_chain_desc *desc,
                      const double *pool, int64_t pool_count);

/*
 * The pre-run chain set of get_psi (run_utils.py:107-114) for every uploaded gene, built
 * inside the library: chain g*T+j = gene g, time point j alone (K = [[theta1]]), start 0,
 * jump variance var0 (0.1 in the reference), M/initial/gap = 100/100/50 in the reference.
 */
int dicegp_set_prerun_chains(dicegp_ctx *ctx, double theta1, double var0, int32_t M,
                             int32_t initial, int32_t gap);

/*
 * Build the main-run chains from finished pre-run chains ON DEVICE: chain g*T+j must be
 * the pre-run of gene g, time j (M=100).  Computes
 *     var[c] = (sum_j Var(Y_j[int(m/4):, c, 0]) + 1e-6) / T       run_utils.py:115-116
 * and replaces the chain set by G main chains (gene g, all T time points, y0 = 0).
 * kinv/prop_factor are T x T, prior_const / logdet_K as above; jump constants are derived
 * on device: jump_scale = var*5/(T*C*theta1), jump_const = prior_const - T*log(jump_scale)/2.
 * var_out (may be NULL) receives G x max_isoforms doubles (row g: C_g-1 valid entries).
 */
int dicegp_main_chains_from_prerun(dicegp_ctx *ctx, int32_t M, int32_t initial, int32_t gap,
                                   double theta1, const double *kinv,
                                   const double *prop_factor, double prior_const,
                                   double *var_out);

/* ------------------------------------------------------------------ running -------- */
/*
 * Advance the listed chains by up to n_steps MH steps using HOST-generated streams
 * (the reference consumes numpy's MT19937: model_GP.py:329,351).
 *  chain_ids[i]   chains to advance (each at most once per call)
 *  delta          for i, for step s < n_steps, for c < C-1, for t < Tc: proposal increment
 *                 Y_try[c,t] - Y_now[c,t] (= z @ A of numpy's multivariate_normal)
 *  delta_off[i]   offset (doubles) of chain i's block in `delta`
 *  u              for i, for step s: the uniform of model_GP.py:351; block i at i*n_steps
 * A chain stops early inside the call when it converges or reaches M; remaining stream
 * entries are ignored.  The call is synchronous (returns after the kernel finished).
 */
int dicegp_run_host_stream(dicegp_ctx *ctx, int32_t n, const int32_t *chain_ids,
                           int32_t n_steps, const double *delta, const int64_t *delta_off,
                           int64_t delta_count, const double *u);

/*
 * Run every unfinished chain to convergence / M with a DEVICE-generated stream
 * (Philox4x32-10 keyed by (seed, stream_id[chain], step); counter based, so results do not
 * depend on how chains are spread over GPUs).  stream_id may be NULL (= chain id).
 * get_psi never seeds (run_utils.py:112-122), so production runs need not follow MT19937.
 * Chains over several time points run one thread block (or cluster) each, in passes: every
 * chain to step 4096 (1024 / 2048 for batches of <= 3000 / <= 7000 chains) in throughput shapes,
 * then the stragglers with one SM each (16-candidate rounds), and -- when fewer chains than SMs are
 * left -- the survivors of every 4096 further steps on clusters of 2 / 4 / 8 SMs per chain (large genes
 * always run on clusters of 2 ... 16 SMs);
 * single-time-point chains -- the pre-runs of get_psi, static mode -- of a large enough batch run
 * one warp each to their end.
 * Which kernel advances a chain does not change its accept decisions (same stream, same
 * expressions); DICEGP_PRERUN_KERNEL=0/1 pins the choice.
 */
int dicegp_run_device_stream(dicegp_ctx *ctx, uint64_t seed, const int64_t *stream_id);

/*
 * The device-generated stream of one installed chain, written out instead of consumed: for steps
 * [step0, step0 + n_steps) the proposal increments delta[s][c][t] = Y_try - Y_now (c < C-1, the
 * role of np.random.multivariate_normal's output minus its mean, model_GP.py:329) and the
 * uniforms u[s] (np.random.rand, model_GP.py:351), bit for bit what dicegp_run_device_stream
 * feeds the chain with for (seed, stream_id).  A test instrument: replaying the dump through a
 * CPU restatement of Psi_GP_MH step-checks the device-stream mode.  Synchronous.
 */
int dicegp_dump_stream(dicegp_ctx *ctx, int32_t chain, uint64_t seed, int64_t stream_id,
                       int32_t step0, int32_t n_steps, double *delta, double *u);

/* ------------------------------------------------------------------ results -------- */
/*
 * Per-chain status: m_ret is the `m` the reference returns (model_GP.py:395: number of
 * kept rows when converged, M-1 otherwise), n_rows the number of trace rows kept, cnt the
 * number of acceptances, state a dicegp_chain_state, flags bit0 = a non-finite
 * log-probability was seen.  Any output pointer may be NULL.
 */
int dicegp_chain_status(dicegp_ctx *ctx, int32_t n, const int32_t *chain_ids, int32_t *m_ret,
                        int32_t *n_rows, int32_t *cnt, int32_t *state, int32_t *flags);

/*
 * Trace download for one chain.  Row layout matches the reference's return arrays:
 * psi / y are (n_rows, C, Tc) row-major (Psi_all, Y_all), pro / lik are (n_rows,)
 * (Pro_all, Lik_all), model_GP.py:362-366,395.  first_row/n_rows select a row range.
 */
int dicegp_download_trace(dicegp_ctx *ctx, int32_t chain, int32_t first_row, int32_t n_rows,
                          double *psi, double *y, double *pro, double *lik);

/*
 * Posterior summary on device (run_utils.py:125-131) for the listed chains:
 *  psi_mean  (C x Tc)           mean of Psi_all[int(m/4):]
 *  psi_ci    (C x Tc x 2)       [.., 0] = sorted[-idx], [.., 1] = sorted[idx], idx = int(n*0.025)
 *  lik_marg  scalar             harmonic-mean log-likelihood of Lik_all[int(m/4):]
 * Outputs are packed per chain at out_off[i] (doubles, for psi_mean; psi_ci at 2x that).
 */
int dicegp_posterior_summary(dicegp_ctx *ctx, int32_t n, const int32_t *chain_ids,
                             double *psi_mean, double *psi_ci, const int64_t *out_off,
                             double *lik_marg);

/*
 * Log-likelihood of ONE given latent state, per time point (model_GP.py:335-345):
 *   psi[:,t] = softmax(y[:,t]); fsi = len_t*psi/sum(len_t*psi);
 *   lik[t]   = sum_n log( sum_k W_t[n,k] * fsi[k,t] )        for t in [t0, t0+Tc) of `gene`.
 * y: C x Tc row-major (host), lik: Tc doubles (host); synchronous.  The caller adds the
 * time points in ascending order like model_GP.py:338-345.  This is the device half of
 * the theta2-sampling mode (theta2=None, model_GP.py:317-326): there the number of
 * normals the reference draws per step depends on the chain state (rejection loop of the
 * truncated random walk) and the proposal covariance changes with every theta2_try, so
 * the stream cannot be produced ahead of the chain; the host proposes step by step and
 * the device evaluates the reads.  Needs dicegp_upload_genes only (no chains).
 */
int dicegp_eval_loglik(dicegp_ctx *ctx, int32_t gene, int32_t t0, int32_t Tc, const double *y,
                       double *lik);

/* Device time (ms, CUDA events on the launch stream) and kernel launches of the most
 * recent run_* call, and the MH steps x reads it evaluated.  For bench.py. */
int dicegp_last_run_stats(dicegp_ctx *ctx, double *kernel_ms, int64_t *launches,
                          int64_t *read_evals, int64_t *steps);

/* Device timer around any sequence of calls: CUDA events on the handle's launch stream
 * (every kernel and copy of the library is ordered on, or joined to, that stream). */
int dicegp_timer_start(dicegp_ctx *ctx);
int dicegp_timer_stop(dicegp_ctx *ctx, double *elapsed_ms);

/* Totals since dicegp_create: kernels launched, bytes uploaded with the gene batches. */
int dicegp_counters(dicegp_ctx *ctx, int64_t *kernel_launches, int64_t *h2d_bytes,
                    int64_t *d2h_bytes);

/* Work the chain kernels themselves counted since the last reset (a few atomic adds per chain and
 * launch), out[0 .. n_out), n_out <= 8:
 *   [0] bytes of W brought from L2 / HBM into shared memory (a resident chain counts its W once per launch,
 *       a streamed chain once per round), [1] rounds = passes over W, [2] candidate evaluations = read
 *       likelihoods evaluated including speculated candidates that were not committed (rounds x candidates
 *       x reads), [3] MH steps committed, [4] fp64 fused multiply-adds of [2] (x isoforms).
 * For the roofline figures of bench.py. */
int dicegp_work_counters(dicegp_ctx *ctx, int64_t *out, int32_t n_out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* DICEGP_H */
