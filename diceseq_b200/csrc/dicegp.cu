// dicegp.cu -- kernels' global entry points and the C ABI declared in include/dicegp.h.
//
// Host side: device tables for genes and chains, launch classes (which chains run with W
// resident in shared memory at which occupancy, which stream W from L2/HBM), host- and
// device-generated stream drivers, status / trace download.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dicegp.h"

namespace dg {
constexpr int DICEGP_CHAIN_CONVERGED_ = 1;
constexpr int DICEGP_CHAIN_EXHAUSTED_ = 2;
constexpr int DICEGP_CHAIN_NOMEM_ = 3;
}  // namespace dg
#include "dicegp_chain.cuh"
#include "dicegp_wide.cuh"
#include "dicegp_prerun.cuh"

// =======================================================================================
// kernels
// =======================================================================================
// Static assignment (queue == nullptr): block b advances chain ids[b].
// Persistent (queue != nullptr): blocks pull slots from an atomic counter until empty;
// ids are sorted by decreasing cost so the heavy chains start first.
#ifndef DG_LB_THREADS
#define DG_LB_THREADS 512                // 512 threads/SM -> up to 128 registers per thread
#endif
// Candidate nodes evaluated per round (= per pass over W).  Resident tiers are bound by the
// serial part of a round: 3 nodes.  Streamed tiers also pay for the bytes of W they pull from
// L2 / HBM every round: with wide genes a 4th node commits more steps per byte than it costs.
#ifndef DG_NODES_RESIDENT
#define DG_NODES_RESIDENT 3
#endif
// (measured per isoform count on the bench workload: 4 nodes win from C = 4 up, 5-6 nodes lose)
__host__ __device__ constexpr int dg_nodes_streamed(int KC) {
#ifdef DG_NODES_STREAMED
    return DG_NODES_STREAMED;
#else
    return (KC >= 4 && KC <= 8) ? 4 : 3;
#endif
}
// Registers: the streamed variants with a compiled isoform count run at most DG_LB_THREADS_STAGED threads per SM
// (2-3 chains of 4-6 warps; measured shapes), which leaves them 160 registers per thread; under the 128 of a
// 512-thread bound ptxas spilled ~300 bytes per thread into the round loop (5 % of the load instructions, long-scoreboard
// stalls on loop counters).  The resident and run-time-C variants keep 512 threads x 128 registers.
#ifndef DG_LB_THREADS_STAGED
#define DG_LB_THREADS_STAGED 384
#endif
template <int KC, int WMODE, bool CLUSTER = false, int NJ = DG_NODES_RESIDENT>
__global__ void __maxnreg__((KC > 0 && WMODE == dg::W_STAGED) ? (65536 / DG_LB_THREADS_STAGED) / 16 * 16 : (65536 / DG_LB_THREADS) / 16 * 16)   // registers are allocated in units of 16 per thread
dicegp_chain_kernel(GeneTab gt, ChainTab ct, StreamArgs sa, const int32_t *__restrict__ ids, int n_ids,
                    int n_steps, int use_tma, int *queue) {
    extern __shared__ __align__(128) double dg_smem[];
    __shared__ int next_slot;
    // cluster launch (large genes): the blocks of a cluster work on the same chain
    const int cs = CLUSTER ? (int)dg::cluster_size() : 1, crank = CLUSTER ? (int)dg::cluster_rank() : 0;
    int slot = blockIdx.x / cs;
    while (true) {
        if (queue != nullptr) {
            if (cs == 1) {
                if (threadIdx.x == 0) next_slot = atomicAdd(queue, 1);
                __syncthreads();
                slot = next_slot;
                __syncthreads();
            } else {
                if (crank == 0 && threadIdx.x == 0) {
                    const int sl = atomicAdd(queue, 1);
                    for (int r = 0; r < cs; ++r) dg::st_cluster(&next_slot, r, sl);
                }
                dg::cluster_barrier();
                slot = next_slot;
                dg::cluster_barrier();
            }
        }
        if (slot >= n_ids) break;
        // every block of the cluster reads the chain state BEFORE the lead block may change it
        const bool running = ct.state[ids[slot]] == 0;
        if (cs > 1) dg::cluster_barrier();
        if (running) dg::run_chain<KC, WMODE, NJ, CLUSTER>(gt, ct, sa, ids[slot], slot, n_steps, dg_smem, use_tma, cs, crank);
        if (cs == 1) __syncthreads(); else dg::cluster_barrier();
        if (queue == nullptr) break;
    }
}

// numpy pairwise sum (n <= 128) over a strided global series, used where the reference
// reduces a 1-D view (np.var of a (rows,1) slice collapses to 1-D and goes pairwise).
__device__ double dg_pw_block(const double *v, int n, size_t stride) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += v[i * stride];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = v[j * stride];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += v[(i + j) * stride];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += v[i * stride];
    return res;
}
// numpy splits n > 128 recursively (n2 = n/2 rounded down to a multiple of 8); walked below with an
// explicit stack because device recursion needs a run-time stack size.
// The same pairwise sum over column `col` of trace rows [first, first + n) (any n), optionally of the
// squared deviations from `mean` (np.var's second pass).
__device__ double dg_pw_trace_block(const TraceView &tv, int first, int col, int n, bool sq, double mean) {
    auto v = [&](int i) {
        const double x = __ldcg(tv.row(first + i) + col);
        if (!sq) return x;
        const double e = x - mean;
        return __dmul_rn(e, e);
    };
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += v(i);
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = v(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += v(i + j);
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += v(i);
    return res;
}
__device__ double dg_pw_trace(const TraceView &tv, int first, int col, int n, bool sq, double mean) {
    if (n <= 128) return dg_pw_trace_block(tv, first, col, n, sq, mean);
    int f_off[32], f_n[32], f_state[32], fs = 0, rs = 0;
    double res[34];
    f_off[0] = 0; f_n[0] = n; f_state[0] = 0; fs = 1;
    while (fs > 0) {
        const int k = fs - 1;
        if (f_n[k] <= 128) {
            res[rs++] = dg_pw_trace_block(tv, first + f_off[k], col, f_n[k], sq, mean);
            --fs;
            continue;
        }
        int n2 = f_n[k] / 2;
        n2 -= n2 % 8;
        if (f_state[k] == 0) {
            f_state[k] = 1;
            f_off[fs] = f_off[k]; f_n[fs] = n2; f_state[fs] = 0; ++fs;
        } else if (f_state[k] == 1) {
            f_state[k] = 2;
            f_off[fs] = f_off[k] + n2; f_n[fs] = f_n[k] - n2; f_state[fs] = 0; ++fs;
        } else {
            const double r = res[rs - 1], l = res[rs - 2];
            rs -= 2;
            res[rs++] = l + r;
            --fs;
        }
    }
    return res[0];
}

// Jump variance from the pre-runs (run_utils.py:111-116).  One thread per (gene, isoform):
//   var[c] = (sum_j np.var(Y_j[int(m_j/4):, c, 0]) + 1e-6) / T
// np.var(axis=0) of the (rows, C-1) slice adds rows sequentially, except that a (rows, 1)
// slice (C == 2) collapses to 1-D and is summed pairwise (probed on numpy 2.3.5).
__global__ void dicegp_prerun_var_kernel(GeneTab gt, ChainTab ct, int G, double *__restrict__ var_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = idx / DG_MAXC, c = idx % DG_MAXC;
    if (g >= G) return;
    const int C = gt.C[g], T = gt.T;
    if (c >= C - 1) { var_out[(size_t)g * DG_MAXC + c] = 0.0; return; }
    double acc = 0.0;
    for (int j = 0; j < T; ++j) {
        const int chain = g * T + j;
        const int state = ct.state[chain];
        const int m_ret = ct.m_ret[chain];
        const int n_rows = (state == dg::DICEGP_CHAIN_CONVERGED_) ? m_ret : ct.M[chain];
        const int b = (int)((double)m_ret / 4.0);
        const int n = n_rows - b;
        const int rowlen = 2 * C + 2;                       // Tc == 1
        TraceView tv;                                       // rows through the page table (M may exceed a page)
        tv.pool = ct.trace; tv.tab = ct.ptab + ct.ptab_off[chain]; tv.shift = ct.pshift[chain]; tv.rowlen = rowlen;
        auto yv = [&](int i) { return __ldcg(tv.row(b + i) + C + c); };
        double mean, ss;
        if (C == 2 && n <= 128) {
            // a (rows, 1) slice collapses to 1-D: pairwise sums (one block of <= 128 values)
            double d[128];
            for (int i = 0; i < n; ++i) d[i] = yv(i);
            mean = dg_pw_block(d, n, 1) / (double)n;
            for (int i = 0; i < n; ++i) { const double e = d[i] - mean; d[i] = __dmul_rn(e, e); }   // numpy squares, then adds: no fma
            ss = dg_pw_block(d, n, 1);
        } else if (C == 2) {
            mean = dg_pw_trace(tv, b, C + c, n, false, 0.0) / (double)n;
            ss = dg_pw_trace(tv, b, C + c, n, true, mean);
        } else {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += yv(i);
            mean = s / (double)n;
            ss = 0.0;
            for (int i = 0; i < n; ++i) { const double e = yv(i) - mean; ss += __dmul_rn(e, e); }
        }
        acc += ss / (double)n;
    }
    var_out[(size_t)g * DG_MAXC + c] = (acc + 0.000001) / (double)T;
}

// Copies rows [first, first+n) of one chain's paged trace into a dense (n x rowlen) buffer.
__global__ void dicegp_gather_rows_kernel(ChainTab ct, int chain, int rowlen, int first, int n,
                                          double *__restrict__ out) {
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ct.ptab + ct.ptab_off[chain]; tv.shift = ct.pshift[chain]; tv.rowlen = rowlen;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const double *src = tv.row(first + r);
        for (int i = threadIdx.x; i < rowlen; i += blockDim.x) out[(size_t)r * rowlen + i] = __ldcg(src + i);
    }
}

// ---------------------------------------------------------------------------------------
// The device random stream of one chain, written out instead of consumed: the proposal
// increments delta[s][c][t] and the uniforms u[s] of steps [step0, step0+n) exactly as the chain
// kernels form them (dg::produce_stream / dicegp_prerun_kernel: same Philox counters, the same
// Box-Muller pairing, z @ A_K accumulated j ascending with fma, times sqrt(jump_scale[c])).
// Test instrument: the dumped stream is replayed through the oracle to step-check the
// device-stream mode (tests/test_gpu_device_stream.py).  One block; z staged in shared memory.
// ---------------------------------------------------------------------------------------
__global__ void dicegp_stream_dump_kernel(GeneTab gt, ChainTab ct, int chain, uint64_t seed, int64_t sid, int step0,
                                          int n, double *__restrict__ delta_out, double *__restrict__ u_out) {
    __shared__ double z[DG_MAXCT];
    const int C = gt.C[ct.gene[chain]], Tc = ct.Tc[chain], CTm = (C - 1) * Tc;
    const double *pfac = ct.pool + ct.pf_off[chain], *js = ct.pool + ct.js_off[chain];
    const int npair = (CTm + 1) >> 1;
    for (int s = 0; s < n; ++s) {
        const uint32_t step = (uint32_t)(step0 + s);
        for (int i = threadIdx.x; i < npair; i += blockDim.x) {
            double z0, z1;
            dg_normal_pair(seed, sid, step, (uint32_t)i, z0, z1);
            z[2 * i] = z0;
            if (2 * i + 1 < CTm) z[2 * i + 1] = z1;
        }
        if (threadIdx.x == 0) u_out[s] = dg_uniform(seed, sid, step);
        __syncthreads();
        for (int i = threadIdx.x; i < CTm; i += blockDim.x) {
            const int c = i / Tc, t = i - c * Tc;
            double a = 0.0;
            for (int j = 0; j < Tc; ++j) a = fma(z[c * Tc + j], pfac[j * Tc + t], a);
            delta_out[(size_t)s * CTm + i] = sqrt(js[c]) * a;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Log-likelihood of ONE given latent state (model_GP.py:335-345), for chains whose proposal
// stream depends on the chain state and are therefore driven step by step from the host
// (theta2 sampling, model_GP.py:317-326).  grid = (blocks per time point, Tc): a block forms
//   e = exp(y[:, t]), f = len*e / sum(len*e)       (= fsi of model_GP.py:336-337)
// for its time point, sums log(sum_k W_t[n,k] f[k]) over its slice of the reads with real
// logs (so that -inf / NaN propagate like np.log(...).sum()), and the last block to finish
// a time point adds the per-block partial sums in block order (deterministic).
// ---------------------------------------------------------------------------------------
#define DG_EVAL_THREADS 256
__global__ void __launch_bounds__(DG_EVAL_THREADS)
dicegp_loglik_kernel(GeneTab gt, int g, int t0, const double *__restrict__ y, int Tc,
                     double *__restrict__ partial, unsigned *__restrict__ done, double *__restrict__ lik_out) {
    __shared__ double f[DG_MAXC];
    __shared__ double red[DG_EVAL_THREADS / 32];
    __shared__ bool last;
    const int tt = blockIdx.y, t = t0 + tt, C = gt.C[g], T = gt.T;
    const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t;
    const int n = gt.ncnt[(size_t)g * T + t], padn = roff[1] - roff[0];
    const double *Wt = gt.W + gt.woff[g] + (size_t)C * roff[0];
    if (threadIdx.x == 0) {
        const double *len = gt.len + gt.lenoff[g] + (size_t)t * C;
        double tot = 0.0;
        for (int k = 0; k < C; ++k) { f[k] = len[k] * exp(y[(size_t)k * Tc + tt]); tot += f[k]; }
        for (int k = 0; k < C; ++k) f[k] = f[k] / tot;
    }
    __syncthreads();
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x = 0.0;
        for (int k = 0; k < C; ++k) x = fma(__ldg(Wt + (size_t)k * padn + i), f[k], x);
        acc += log(x);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DG_EVAL_THREADS / 32; ++w) s += red[w];
        partial[(size_t)tt * gridDim.x + blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(done + tt, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            double tot = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) tot += __ldcg(partial + (size_t)tt * gridDim.x + b);
            lik_out[tt] = tot;
            done[tt] = 0;                   // ready for the next call
        }
    }
}

// ---------------------------------------------------------------------------------------
// Posterior summary of one chain per block (run_utils.py:14-30, 125-129):
//   burn-in b = int(m/4); psi_mean = Psi_all[b:].mean(0) (rows added sequentially, like
//   numpy's axis-0 reduction); 95% interval = order statistics sorted[-k], sorted[k],
//   k = int(n*(1-0.95)/2), found EXACTLY with an 8-bit radix select over the column held in
//   shared memory; lik_marg = log n + min - log(sum exp(min - Lik)).
// ---------------------------------------------------------------------------------------
#define DG_SUM_THREADS 256
// Two order statistics of the same n keys in one sweep: rank[0]-th and rank[1]-th smallest
// (0-based), by 8-bit radix selection from the top digit down.  Both selections share every
// pass over the keys (two histograms); the histogram updates are aggregated per warp, because
// the keys of a trace column share their sign / exponent digits and would otherwise all hit
// the same bin.  hist: 2 x 256 bins; sh: 4 ints.  Block-wide; every thread gets both results.
__device__ void dg_radix_select2(const unsigned long long *keys, int n, int rank0, int rank1, unsigned *hist, int *sh,
                                 unsigned long long &out0, unsigned long long &out1) {
    unsigned long long prefix[2] = {0, 0}, mask = 0;
    int rank[2] = {rank0, rank1};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_up = (n + 31) & ~31;                      // whole warps stay in the loop (match needs them)
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const bool same = prefix[0] == prefix[1];         // both selections still walk the same bucket
        for (int i = threadIdx.x; i < n_up; i += blockDim.x) {
            const bool ok = i < n;
            const unsigned long long k = ok ? keys[i] : 0ull;
            const unsigned d = (unsigned)(k >> shift) & 255u;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && same) break;
                const bool in = ok && (k & mask) == prefix[s];
                if (shift >= 48) {                        // sign / exponent digits: a handful of distinct values
                    const unsigned peers = __match_any_sync(0xffffffffu, in ? d : 256u + (unsigned)lane);
                    if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[s * 256 + d], (unsigned)__popc(peers));
                } else if (in) {
                    atomicAdd(&hist[s * 256 + d], 1u);
                }
            }
        }
        __syncthreads();
        if (warp < 2) {
            // warp s scans the 256 bins of selection s: 8 bins per lane
            const unsigned *h = hist + (same ? 0 : warp * 256);
            unsigned loc[8], tot = 0;
            for (int j = 0; j < 8; ++j) { loc[j] = h[lane * 8 + j]; tot += loc[j]; }
            unsigned incl = tot;
            for (int off = 1; off < 32; off <<= 1) {
                unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            const unsigned before = incl - tot;
            const unsigned r = (unsigned)rank[warp];
            if (r >= before && r < incl) {
                unsigned acc = before;
                for (int j = 0; j < 8; ++j) {
                    if (r < acc + loc[j]) { sh[2 * warp] = lane * 8 + j; sh[2 * warp + 1] = (int)(r - acc); break; }
                    acc += loc[j];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            prefix[s] |= (unsigned long long)(unsigned)sh[2 * s] << shift;
            rank[s] = sh[2 * s + 1];
        }
        mask |= 0xffull << shift;
        __syncthreads();
    }
    out0 = prefix[0];
    out1 = prefix[1];
}

// One block per chain, two phases.
//   1. The rows kept after the burn-in are read ONCE, coalesced (chunks of <= 64 trace columns through a
//      32-row shared-memory tile): the column means are accumulated in row order and the columns are written
//      back TRANSPOSED into a scratch (column-major: n contiguous doubles per column, the log-likelihood as one
//      more column).  A trace row is 2CT+2 doubles, so reading one column straight from the trace touches one
//      128-byte line per value; the round-1 kernel read 105 GB for 19 GB of trace that way.
//   2. Per column: the keys come from the scratch (contiguous), exact order statistics by radix select.
#define DG_SUM_TILE_COLS 64
__global__ void __launch_bounds__(DG_SUM_THREADS)
dicegp_summary_kernel(GeneTab gt, ChainTab ct, const int32_t *__restrict__ ids, double *__restrict__ mean_out,
                      double *__restrict__ ci_out, const int64_t *__restrict__ out_off, double *__restrict__ lik_out,
                      double *__restrict__ scratch, const int64_t *__restrict__ scr_off, int keys_in_smem, int parts) {
    extern __shared__ __align__(16) unsigned long long dg_keys_smem[];
    // phase 1 uses the dynamic shared memory as a 32 x 65 tile, phase 2 as one column of keys
    double (*tile)[DG_SUM_TILE_COLS + 1] = reinterpret_cast<double (*)[DG_SUM_TILE_COLS + 1]>(dg_keys_smem);
    __shared__ unsigned hist[512];
    __shared__ int sh[4];
    __shared__ double red[DG_SUM_THREADS / 32];
    // `parts` blocks share one chain (the long chains: a 20000-step chain is 15000 rows x CT columns of radix
    // selects, 9 ms in one block): block `part` owns the columns [c_lo, c_hi); the last one also the log-likelihood
    const int slot = blockIdx.x / parts, part = blockIdx.x - slot * parts;
    const int chain = ids[slot];
    const int C = gt.C[ct.gene[chain]], Tc = ct.Tc[chain], CT = C * Tc, rowlen = 2 * CT + 2;
    const int c_lo = (int)((long long)part * CT / parts), c_hi = (int)((long long)(part + 1) * CT / parts);
    const bool owns_lik = part == parts - 1;
    const int state = ct.state[chain], m_ret = ct.m_ret[chain];
    const int n_rows = state == dg::DICEGP_CHAIN_CONVERGED_ ? m_ret
                       : (state == dg::DICEGP_CHAIN_EXHAUSTED_ ? ct.M[chain] : ct.m_next[chain]);
    const int b = m_ret > 0 ? (m_ret >> 2) : 0;                   // int(m/4)
    const int n = n_rows - b;
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ct.ptab + ct.ptab_off[chain]; tv.shift = ct.pshift[chain]; tv.rowlen = rowlen;
    double *mean = mean_out + out_off[slot];
    double *ci = ci_out + 2 * out_off[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) {
        for (int i = c_lo + tid; i < c_hi; i += blockDim.x) { mean[i] = nan(""); ci[2 * i] = nan(""); ci[2 * i + 1] = nan(""); }
        if (tid == 0 && owns_lik) lik_out[slot] = nan("");
        return;
    }
    double *cols = scratch + scr_off[slot];                       // (CT + 1) columns of n doubles
    // ---- phase 1: means (rows in order, like numpy's axis-0 reduction) + transposed copy ----
    for (int c0 = c_lo; c0 < c_hi; c0 += DG_SUM_TILE_COLS) {
        const int w = min(DG_SUM_TILE_COLS, c_hi - c0);
        const bool last = owns_lik && c0 + w == CT;
        const int wl = w + (last && w < DG_SUM_TILE_COLS ? 1 : 0);   // the log-likelihood column rides along when there is room
        double a = 0.0;
        for (int r0 = 0; r0 < n; r0 += 32) {
            const int nr = min(32, n - r0);
            for (int idx = tid; idx < nr * wl; idx += DG_SUM_THREADS) {
                const int r = idx / wl, j = idx - r * wl;
                tile[r][j] = __ldcg(tv.row(b + r0 + r) + (j < w ? c0 + j : 2 * CT + 1));
            }
            __syncthreads();
            if (tid < w) for (int r = 0; r < nr; ++r) a += tile[r][tid];
            for (int idx = tid; idx < wl * 32; idx += DG_SUM_THREADS) {
                const int j = idx >> 5, r = idx & 31;
                if (r < nr) cols[(size_t)(j < w ? c0 + j : CT) * n + r0 + r] = tile[r][j];
            }
            __syncthreads();
        }
        if (tid < w) mean[c0 + tid] = a / (double)n;
        if (last && wl == w) {                                    // no room in the last chunk: the log-likelihood column alone
            for (int i = tid; i < n; i += DG_SUM_THREADS) cols[(size_t)CT * n + i] = __ldcg(tv.row(b + i) + 2 * CT + 1);
        }
    }
    __syncthreads();
    // ---- phase 2: per column radix select of ranks k and n-k (k = 0 -> both the minimum side quirk) ----
    const int k = (int)((double)n * (1.0 - 0.95) / 2.0);
    for (int col = c_lo; col < c_hi; ++col) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(cols + (size_t)col * n);
        const unsigned long long *keys = src;                     // long chains select in the scratch itself
        if (keys_in_smem) {
            __syncthreads();
            for (int i = tid; i < n; i += blockDim.x) dg_keys_smem[i] = __ldcg(src + i);
            __syncthreads();
            keys = dg_keys_smem;
        }
        unsigned long long lo, hi;
        dg_radix_select2(keys, n, k, k > 0 ? n - k : 0, hist, sh, lo, hi);
        if (tid == 0) {
            ci[2 * col] = __longlong_as_double((long long)hi);       // sorted[-k]
            ci[2 * col + 1] = __longlong_as_double((long long)lo);   // sorted[k]
        }
    }
    // harmonic-mean log-likelihood
    if (!owns_lik) return;
    if (c_lo == c_hi) {                                           // more parts than columns: this block copied nothing yet
        for (int i = tid; i < n; i += DG_SUM_THREADS) cols[(size_t)CT * n + i] = __ldcg(tv.row(b + i) + 2 * CT + 1);
        __syncthreads();
    }
    const double *lk = cols + (size_t)CT * n;
    double mn = INFINITY;
    for (int i = tid; i < n; i += blockDim.x) mn = fmin(mn, __ldcg(lk + i));
    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    __syncthreads();
    if (lane == 0) red[warp] = mn;
    __syncthreads();
    mn = red[0];
    for (int w = 1; w < DG_SUM_THREADS / 32; ++w) mn = fmin(mn, red[w]);
    double se = 0.0;
    for (int i = tid; i < n; i += blockDim.x) se += exp(mn - __ldcg(lk + i));
    for (int off = 16; off > 0; off >>= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
    __syncthreads();
    if (lane == 0) red[warp] = se;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < DG_SUM_THREADS / 32; ++w) t += red[w];
        lik_out[slot] = log((double)n) + mn - log(t);
    }
}

// Tiled copy of W for the wide-round kernel (GeneTab::Wt): one block per gene.
__global__ void dicegp_tile_kernel(GeneTab gt, int G, double *__restrict__ Wt) {
    const int g = blockIdx.x;
    if (g >= G) return;
    const int C = gt.C[g], T = gt.T;
    const double2 *Wg = reinterpret_cast<const double2 *>(gt.W + gt.woff[g]);
    double2 *out = reinterpret_cast<double2 *>(Wt + gt.wtoff[g]);
    const int32_t *roff = gt.roff + (size_t)g * (T + 1), *toff = gt.toff + (size_t)g * (T + 1);
    for (int t = 0; t < T; ++t) {
        const int np = (roff[t + 1] - roff[t]) >> 1;
        const double2 *src = Wg + (size_t)C * (roff[t] >> 1);
        double2 *dst = out + (size_t)toff[t] * C * 16;
        const int ntile = toff[t + 1] - toff[t];
        const int total = ntile * C * 16;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int tile = i / (C * 16), r = i - tile * (C * 16);
            const int k = r >> 4, w = (tile << 4) + (r & 15);
            dst[i] = w < np ? src[(size_t)k * np + w] : make_double2(0.0, 0.0);
        }
    }
}

// =======================================================================================
// host context
// =======================================================================================
namespace {

thread_local std::string g_last_error;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t resize(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

constexpr int kStreams = 16;

struct LaunchClass {
    int threads;
    size_t smem_limit;      // dynamic shared memory per block
    int blocks_per_sm;
    bool w_in_smem;
};

}  // namespace

struct dicegp_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t static_smem = 0;      // static shared memory of the chain kernel
    std::string err;
    cudaStream_t stream = nullptr;
    cudaStream_t class_stream[16] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join[128] = {}, ev_t0 = nullptr, ev_t1 = nullptr;
    int64_t total_launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    int use_tma = 1;

    // genes
    int G = 0, T = 0;
    std::vector<int32_t> hC, hncnt, hroff, htoff;
    std::vector<int64_t> hwtoff;
    bool tiles_valid = false;               // dWt holds the tiled copy of the uploaded W
    int64_t wt_total = 0;                   // doubles of the tiled copy
    std::vector<int64_t> hwoff, hlenoff;
    DevBuf<int32_t> dC, dncnt, droff, dtoff;
    DevBuf<int64_t> dwtoff;
    DevBuf<double> dWt;
    DevBuf<int64_t> dwoff, dlenoff;
    DevBuf<double> dW, dlen;

    // chains
    int n_chains = 0;
    std::vector<dicegp_chain_desc> hdesc;
    std::vector<int64_t> hptab_off, hst_off;
    std::vector<int32_t> hpshift;
    int64_t trace_budget_bytes = 0;      // 0 = default (a share of the free device memory)
    int64_t pool_cap = 0;
    std::vector<int32_t> hchainC;
    DevBuf<int32_t> cgene, ct0, cTc, cM, cinit, cgap, cm_next, ccnt, cstate, cm_ret, cflags;
    DevBuf<int64_t> ckinv, cpf, cjs, cjc, cy0, cptab_off, cptab, cst_off, cstream_id;
    DevBuf<int32_t> cpshift;
    DevBuf<unsigned long long> cpool_head;
    DevBuf<double> dgather;
    DevBuf<double> cprior, cpool, ctrace, cst;
    bool have_stream_id = false;
    bool chains_fresh = false;           // installed and not advanced yet
    // host staging of install_chains / set_prerun_chains: kept between calls (a batch installs 80k
    // pre-run chains and 10k joint chains in turn; fresh multi-MB vectors every time mean mmap /
    // page-fault churn, seen as sporadic 50 ms stalls)
    std::vector<int64_t> h_ptab;
    std::vector<int32_t> h_i32[7];
    std::vector<int64_t> h_i64[5];
    std::vector<double> h_prior;
    std::vector<dicegp_chain_desc> h_desc_tmp;

    // run scratch
    DevBuf<int32_t> dids;
    DevBuf<int64_t> ddoff;
    DevBuf<double> ddelta, du, dvar;
    DevBuf<int> dqueue;
    DevBuf<double> smean, sci, slik;       // posterior summary scratch
    DevBuf<unsigned long long> skeys;      // posterior summary: the kept trace rows transposed (column-major per chain)
    DevBuf<int64_t> sscr_off;              // ... first double of every chain's columns
    DevBuf<double> deval;                  // eval_loglik scratch: y | per-block partial sums | lik
    DevBuf<unsigned> deval_done;
    DevBuf<int64_t> soff;
    DevBuf<int32_t> sids;
    DevBuf<unsigned long long> dprof;
    DevBuf<unsigned long long> dctr;        // work counters of the chain kernels (ChainTab::ctr)
    size_t static_smem_wide = 0;
    int wide_cluster_cap[5] = {0, 0, 0, 0, 0};   // chains the wide-round kernel runs at once on clusters of 1 / 2 / 4 / 8 / 16 blocks
    double last_ms = 0.0;
    int64_t last_launches = 0, last_evals = 0, last_steps = 0;
    std::vector<int32_t> hm_next_before;

    int fail(int code, const std::string &msg) {
        err = msg;
        g_last_error = msg;
        return code;
    }
    int cuda_fail(cudaError_t e, const char *what) {
        return fail(DICEGP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    }
};

#define DG_CUDA(ctx, call)                                              \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call);    \
    } while (0)

namespace {

GeneTab make_gene_tab(dicegp_ctx *c) {
    GeneTab g;
    g.C = c->dC.p; g.woff = c->dwoff.p; g.roff = c->droff.p; g.ncnt = c->dncnt.p;
    g.lenoff = c->dlenoff.p; g.W = c->dW.p; g.len = c->dlen.p; g.T = c->T;
    g.Wt = c->dWt.p; g.wtoff = c->dwtoff.p; g.toff = c->dtoff.p;
    return g;
}
ChainTab make_chain_tab(dicegp_ctx *c) {
    ChainTab t;
    t.gene = c->cgene.p; t.t0 = c->ct0.p; t.Tc = c->cTc.p; t.M = c->cM.p; t.initial = c->cinit.p;
    t.gap = c->cgap.p; t.kinv_off = c->ckinv.p; t.pf_off = c->cpf.p; t.js_off = c->cjs.p;
    t.jc_off = c->cjc.p; t.y0_off = c->cy0.p; t.prior_const = c->cprior.p; t.pool = c->cpool.p;
    t.st_off = c->cst_off.p;
    t.ptab_off = c->cptab_off.p; t.pshift = c->cpshift.p; t.ptab = c->cptab.p;
    t.pool_head = c->cpool_head.p; t.pool_cap = c->pool_cap;
    t.stream_id = c->have_stream_id ? c->cstream_id.p : nullptr;
    t.trace = c->ctrace.p; t.st = c->cst.p; t.m_next = c->cm_next.p; t.cnt = c->ccnt.p;
    t.state = c->cstate.p; t.m_ret = c->cm_ret.p; t.flags = c->cflags.p;
    t.ctr = c->dctr.p;
    return t;
}

// Residency tiers: W resident in shared memory at 8/4/2/1 blocks per SM, then the tier
// that streams W from L2/HBM every step.  Each tier is further split by isoform count,
// because the likelihood loop is compiled per C (weights f[C] in registers).
// 0-3 resident, 4 streamed, 5-7 streamed over a cluster of 2/4/8 blocks; wide rounds (dicegp_wide.cuh): 8 = one block per
// chain (or the cluster size of a last-stragglers pass), 9-12 = large genes on a cluster of 2/4/8/16 blocks in the straggler passes
constexpr int kTiers = 13;
constexpr int kWideTier = 8;
constexpr int kKinds = 8;                      // kernel variants by C: generic, 2..8
constexpr int kGroups = kTiers * kKinds;
void launch_classes(const dicegp_ctx *c, LaunchClass out[kTiers]) {
    const size_t per_sm = 233472;  // 228 KB
    const int bps[4] = {8, 4, 2, 1};
    const int thr[4] = {64, 128, 256, 512};
    for (int i = 0; i < 4; ++i) {
        size_t lim = per_sm / bps[i] - 1024;
        if (lim > c->smem_optin - c->static_smem) lim = c->smem_optin - c->static_smem;
        out[i] = {thr[i], lim, bps[i], true};
    }
    out[4] = {256, 0, 2, false};
    for (int i = 5; i < kTiers; ++i) out[i] = {DG_LB_THREADS, 0, 1, false};
    out[kWideTier] = {DG_WIDE_THREADS, 0, 1, false};
    // development overrides: DICEGP_TIER_THREADS / DICEGP_TIER_BPS = five comma-separated ints
    const char *et = getenv("DICEGP_TIER_THREADS"), *eb = getenv("DICEGP_TIER_BPS");
    if (et) sscanf(et, "%d,%d,%d,%d,%d", &out[0].threads, &out[1].threads, &out[2].threads, &out[3].threads, &out[4].threads);
    if (eb) sscanf(eb, "%d,%d,%d,%d,%d", &out[0].blocks_per_sm, &out[1].blocks_per_sm, &out[2].blocks_per_sm, &out[3].blocks_per_sm, &out[4].blocks_per_sm);
    for (int i = 0; i < kWideTier; ++i) {
        if (out[i].threads > DG_LB_THREADS) out[i].threads = DG_LB_THREADS;
        // register file: DG_LB_THREADS threads per SM at the compiled register count
        if (out[i].threads * out[i].blocks_per_sm > DG_LB_THREADS) out[i].blocks_per_sm = DG_LB_THREADS / out[i].threads;
    }
}

// warps of a block that own a cp.async staging ring: all of them, unless the state warps take no likelihood items
inline int ring_warps(int threads, int C, int Tc) {
    return DG_STATE_LIK_WEIGHT > 0 ? threads / 32 : threads / 32 - (C * Tc + 31) / 32;
}
// shared memory of a chain besides W / the staging rings; `streamed`: tiers 4-7
size_t fixed_smem_bytes(int C, int Tc, int threads, bool streamed, int cs = 1) {
    const int nodes = streamed ? dg_nodes_streamed(C <= 8 ? C : 0) : DG_NODES_RESIDENT;
    return (size_t)dg_layout(C, Tc, (threads + 31) / 32, nodes, cs).total * sizeof(double);
}
inline int tier_cluster(int tier) { return (tier < 5 || tier >= kWideTier) ? 1 : (2 << (tier - 5)); }

int64_t chain_reads(const dicegp_ctx *c, int chain) {
    const dicegp_chain_desc &d = c->hdesc[chain];
    int64_t n = 0;
    for (int t = 0; t < d.Tc; ++t) n += c->hncnt[(size_t)d.gene * c->T + d.t0 + t];
    return n;
}
int64_t chain_wbytes(const dicegp_ctx *c, int chain) {
    const dicegp_chain_desc &d = c->hdesc[chain];
    const int32_t *ro = c->hroff.data() + (size_t)d.gene * (c->T + 1) + d.t0;
    return (int64_t)c->hC[d.gene] * (ro[d.Tc] - ro[0]) * 8;
}

inline int kind_of(int C) { return (C >= 2 && C <= 8) ? C - 1 : 0; }

// threads a chain needs at least: its state warps (one lane per latent value) plus one
// likelihood warp
inline int min_threads(int C, int Tc) { return ((C * Tc + 31) / 32) * 32 + 32; }

int pick_tier(const dicegp_ctx *c, const LaunchClass cls[kTiers], int chain, int latency_mode) {
    const dicegp_chain_desc &d = c->hdesc[chain];
    const int C = c->hC[d.gene];
    const int64_t wb = chain_wbytes(c, chain);
    // large genes: one chain over a cluster of blocks (the reads split over 2/4/8 SMs)
    static const int64_t wide_min = getenv("DICEGP_WIDE_MIN_BYTES") ? atoll(getenv("DICEGP_WIDE_MIN_BYTES")) : (1ll << 20);
    if (wb >= 16 * wide_min) return 7;
    if (wb >= 4 * wide_min) return 6;
    if (wb >= wide_min) return 5;
    // A resident chain that needs a whole SM leaves it idle during the serial part of every round;
    // four streamed chains per SM overlap each other's serial parts and are faster (measured:
    // 23.5 vs 29.7 ms for 592 two-isoform, 128 KB chains), so the one-block-per-SM resident class
    // is only taken in latency mode (one chain per SM anyway).
    static const int max_resident_env = getenv("DICEGP_MAX_RESIDENT_TIER") ? atoi(getenv("DICEGP_MAX_RESIDENT_TIER")) : -1;
    const int max_resident = max_resident_env >= 0 ? max_resident_env : (latency_mode ? 3 : 2);
    for (int i = 0; i <= max_resident && i < 4; ++i) {
        if (cls[i].threads < min_threads(C, d.Tc)) continue;
        if (fixed_smem_bytes(C, d.Tc, cls[i].threads, false) + (size_t)wb <= cls[i].smem_limit) return i;
    }
    return 4;
}

// ---- wide-round kernel (dicegp_wide.cuh): launch shape per isoform count ----------------------
struct WidePlan { int ncw, npw, bps, rj; };
inline int env_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }
inline size_t wide_fixed_bytes(int C, int Tc, const WidePlan &p) {
    return (size_t)dg::dg_wide_layout(C, Tc, p.ncw, 4 * p.rj, p.npw).ring * sizeof(double);
}
inline size_t wide_budget(const dicegp_ctx *c, const WidePlan &p) {
    return std::min<size_t>(233472 / p.bps - 1024, c->smem_optin - c->static_smem_wide);
}
WidePlan wide_plan(const dicegp_ctx *c, int C, int Tc, int cs = 1) {
    // One block of 16 warps per SM.  15 consumer warps + 1 producer warp for a whole chain (the stream of the ~13 steps
    // a round commits costs one warp ~55k cycles, a little less than the likelihood pass of a C = 8 chain); on a
    // cluster the likelihood pass shrinks with the cluster size while the stream does not: 11 consumers + 5
    // producers.  Every consumer warp needs two tasks (C x 256 bytes each) of ring: genes with many isoforms get
    // fewer consumer warps.  DICEGP_WIDE_NCW / _NPW / _BPS override.
    static const int e_ncw = env_int("DICEGP_WIDE_NCW", 0), e_bps = env_int("DICEGP_WIDE_BPS", 0), e_npw = env_int("DICEGP_WIDE_NPW", 0);
    WidePlan p;
    p.rj = 4;
    p.ncw = cs > 1 ? 11 : 15; p.npw = cs > 1 ? 5 : 1; p.bps = 1;
    if (e_ncw > 0) p.ncw = e_ncw;
    if (e_npw > 0) p.npw = e_npw;
    if (32 * (p.ncw + p.npw) > DG_WIDE_THREADS) p.ncw = DG_WIDE_THREADS / 32 - p.npw;
    if (e_bps > 0) p.bps = e_bps;
    if (32 * (p.ncw + p.npw) * p.bps > DG_WIDE_THREADS) p.bps = std::max(1, DG_WIDE_THREADS / (32 * (p.ncw + p.npw)));
    while (p.ncw > 4 && wide_fixed_bytes(C, Tc, p) + 2 * (size_t)p.ncw * C * 256 > wide_budget(c, p)) --p.ncw;
    return p;
}
bool wide_eligible(const dicegp_ctx *c, int chain, int latency_mode, int cs = 1) {
    // DICEGP_WIDE: 0 never, 1 (default) the straggler pass only, 2 every pass.  In the throughput pass two to
    // four chains of the general kernel per SM overlap each other's serial parts and spend a third less fp64
    // work per committed step (4 candidates per round instead of 16); with one chain per SM, W L2 resident
    // and acceptance rates of a few per cent (the chains still running after 4096 steps) the wide rounds win.
    static const int mode = env_int("DICEGP_WIDE", 1);
    if (mode == 0 || (mode == 1 && !latency_mode)) return false;
    const dicegp_chain_desc &d = c->hdesc[chain];
    const int C = c->hC[d.gene];
    if (C < 2 || C * d.Tc > DG_WIDE_MAXCT) return false;
    const WidePlan p = wide_plan(c, C, d.Tc, cs);
    const size_t stage = (size_t)p.ncw * C * 256;
    return wide_fixed_bytes(C, d.Tc, p) + 2 * stage <= wide_budget(c, p);
}

typedef void (*wide_kernel_t)(GeneTab, ChainTab, StreamArgs, const int32_t *, int, int, int, int, int *);
wide_kernel_t wide_kernel_for(int kind, bool cluster) {
    if (cluster) {
        switch (kind) {
            case 0: return dicegp_wide_kernel<0, 4, true>;
            case 1: return dicegp_wide_kernel<2, 4, true>;
            case 2: return dicegp_wide_kernel<3, 4, true>;
            case 3: return dicegp_wide_kernel<4, 4, true>;
            case 4: return dicegp_wide_kernel<5, 4, true>;
            case 5: return dicegp_wide_kernel<6, 4, true>;
            case 6: return dicegp_wide_kernel<7, 4, true>;
            default: return dicegp_wide_kernel<8, 4, true>;
        }
    }
    switch (kind) {
        case 0: return dicegp_wide_kernel<0, 4>;
        case 1: return dicegp_wide_kernel<2, 4>;
        case 2: return dicegp_wide_kernel<3, 4>;
        case 3: return dicegp_wide_kernel<4, 4>;
        case 4: return dicegp_wide_kernel<5, 4>;
        case 5: return dicegp_wide_kernel<6, 4>;
        case 6: return dicegp_wide_kernel<7, 4>;
        default: return dicegp_wide_kernel<8, 4>;
    }
}

// Chains that run on a cluster never allocate trace pages on the device (only the lead block
// writes the trace, and the blocks must stay in lockstep): give them all their pages now.
int preallocate_trace_pages(dicegp_ctx *c, const std::vector<int32_t> &ids, bool quiet = false) {
    unsigned long long head = 0;
    DG_CUDA(c, cudaMemcpy(&head, c->cpool_head.p, sizeof head, cudaMemcpyDeviceToHost));
    std::vector<std::vector<int64_t>> tabs(ids.size());
    unsigned long long need_all = 0;
    for (size_t i = 0; i < ids.size(); ++i) {                   // first see whether the pool has room for all of them
        const int32_t id = ids[i];
        const dicegp_chain_desc &d = c->hdesc[id];
        const int64_t rowlen = 2 * (int64_t)c->hC[d.gene] * d.Tc + 2;
        const int shift = c->hpshift[id];
        const int n_pages = (d.M + (1 << shift) - 1) >> shift;
        tabs[i].resize(n_pages);
        DG_CUDA(c, cudaMemcpy(tabs[i].data(), c->cptab.p + c->hptab_off[id], n_pages * sizeof(int64_t), cudaMemcpyDeviceToHost));
        for (int p = 0; p < n_pages; ++p)
            if (tabs[i][p] < 0) need_all += (unsigned long long)rowlen << shift;
    }
    if (head + need_all > (unsigned long long)c->pool_cap) {
        if (quiet) return DICEGP_ERR_NOMEM;
        return c->fail(DICEGP_ERR_NOMEM, "trace pool too small for the large genes of this batch; raise the trace budget");
    }
    for (size_t i = 0; i < ids.size(); ++i) {
        const int32_t id = ids[i];
        const dicegp_chain_desc &d = c->hdesc[id];
        const int64_t rowlen = 2 * (int64_t)c->hC[d.gene] * d.Tc + 2;
        const int shift = c->hpshift[id];
        bool changed = false;
        for (size_t p = 0; p < tabs[i].size(); ++p) {
            if (tabs[i][p] >= 0) continue;
            tabs[i][p] = (int64_t)head;
            head += (unsigned long long)rowlen << shift;
            changed = true;
        }
        if (changed)
            DG_CUDA(c, cudaMemcpy(c->cptab.p + c->hptab_off[id], tabs[i].data(), tabs[i].size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    DG_CUDA(c, cudaMemcpy(c->cpool_head.p, &head, sizeof head, cudaMemcpyHostToDevice));
    return DICEGP_OK;
}

typedef void (*chain_kernel_t)(GeneTab, ChainTab, StreamArgs, const int32_t *, int, int, int, int *);
chain_kernel_t chain_kernel_for(int kind, bool w_in_smem, bool cluster) {
    // streamed W: staged through cp.async for compiled isoform counts, plain loads otherwise;
    // cluster launches exist for the streamed variants only
    if (cluster) {
        switch (kind) {
            case 0: return dicegp_chain_kernel<0, dg::W_DIRECT, true, dg_nodes_streamed(0)>;
            case 1: return dicegp_chain_kernel<2, dg::W_STAGED, true, dg_nodes_streamed(2)>;
            case 2: return dicegp_chain_kernel<3, dg::W_STAGED, true, dg_nodes_streamed(3)>;
            case 3: return dicegp_chain_kernel<4, dg::W_STAGED, true, dg_nodes_streamed(4)>;
            case 4: return dicegp_chain_kernel<5, dg::W_STAGED, true, dg_nodes_streamed(5)>;
            case 5: return dicegp_chain_kernel<6, dg::W_STAGED, true, dg_nodes_streamed(6)>;
            case 6: return dicegp_chain_kernel<7, dg::W_STAGED, true, dg_nodes_streamed(7)>;
            default: return dicegp_chain_kernel<8, dg::W_STAGED, true, dg_nodes_streamed(8)>;
        }
    }
    (void)w_in_smem;
    switch (kind) {
        case 0: return dicegp_chain_kernel<0, dg::W_RESIDENT>;
        case 1: return dicegp_chain_kernel<2, dg::W_RESIDENT>;
        case 2: return dicegp_chain_kernel<3, dg::W_RESIDENT>;
        case 3: return dicegp_chain_kernel<4, dg::W_RESIDENT>;
        case 4: return dicegp_chain_kernel<5, dg::W_RESIDENT>;
        case 5: return dicegp_chain_kernel<6, dg::W_RESIDENT>;
        case 6: return dicegp_chain_kernel<7, dg::W_RESIDENT>;
        default: return dicegp_chain_kernel<8, dg::W_RESIDENT>;
    }
}

// How many chains the wide-round kernel can run at once with a cluster of `cs` blocks per chain (one block per SM;
// clusters do not span GPCs, so this is less than SMs / cs): asked of the driver once per context.
int wide_cluster_capacity(dicegp_ctx *c, int cs) {
    const int slot = cs >= 16 ? 4 : (cs >= 8 ? 3 : (cs >= 4 ? 2 : (cs >= 2 ? 1 : 0)));
    if (cs <= 1) return c->sm_count;
    if (c->wide_cluster_cap[slot] == 0) {                       // 0 = not asked yet, -1 = not available
        wide_kernel_t kern = wide_kernel_for(7, true);
        const int smem = (int)(c->smem_optin - c->static_smem_wide);
        int n = 0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(c->sm_count / cs * cs); cfg.blockDim = dim3(DG_WIDE_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            (cs > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ||
            cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = cs > 8 ? -1 : c->sm_count / cs * 3 / 4;        // conservative guess; no non-portable clusters on this device
        }
        c->wide_cluster_cap[slot] = n;
    }
    return c->wide_cluster_cap[slot];
}

// The tiled copy of W (wide-round kernel), built once per uploaded batch, when a wide launch first needs it.
int ensure_tiles(dicegp_ctx *c) {
    if (c->tiles_valid) return DICEGP_OK;
    if (c->dWt.resize(std::max<int64_t>(c->wt_total, 2)) != cudaSuccess)
        return c->fail(DICEGP_ERR_NOMEM, "device allocation failed (tiled copy of W)");
    dicegp_tile_kernel<<<c->G, 256, 0, c->stream>>>(make_gene_tab(c), c->G, c->dWt.p);
    DG_CUDA(c, cudaGetLastError());
    ++c->total_launches;
    c->tiles_valid = true;
    return DICEGP_OK;
}

// Common driver: advance `ids` (host order) by n_steps with stream args `sa`.
// The kernel's slot is the position in the per-group id list, so the per-slot host-stream
// offsets are permuted alongside.
// latency_mode: few chains are left (the stragglers of a batch); give each as many likelihood
// warps as fit instead of packing many chains per SM.
int run_chains(dicegp_ctx *c, const std::vector<int32_t> &ids, int n_steps, StreamArgs sa,
               const std::vector<int64_t> *delta_off, bool persistent, int latency_mode = 0) {
    c->chains_fresh = false;
    LaunchClass cls[kTiers];
    launch_classes(c, cls);
    if (latency_mode) {
        for (int i = 0; i < 4; ++i) { cls[i].threads = DG_LB_THREADS; cls[i].blocks_per_sm = 1; }
    }
    std::vector<int32_t> by_group[kGroups];
    std::vector<int32_t> slot_src[kGroups];
    // last stragglers (latency_mode 2): fewer chains than SMs, so every chain of the wide-round kernel gets a
    // thread-block cluster -- the largest of 8 / 4 / 2 blocks that still lets all chains run at once
    int wide_cs = 1;
    if (latency_mode == 2 && preallocate_trace_pages(c, ids, true) != DICEGP_OK)
        latency_mode = 1;         // no room to give every chain all its trace pages now: no clusters (pages on demand, NOMEM status)
    if (latency_mode == 2 && persistent) {
        const int n = (int)ids.size();
        const int e_cs = env_int("DICEGP_WIDE_CS", 0);
        wide_cs = n <= wide_cluster_capacity(c, 8) ? 8 : (n <= wide_cluster_capacity(c, 4) ? 4 : (n <= wide_cluster_capacity(c, 2) ? 2 : 1));
        if (e_cs == 1 || e_cs == 2 || e_cs == 4 || e_cs == 8) wide_cs = e_cs;
    }
    static const int wide_big = env_int("DICEGP_WIDE_BIG", 1);
    for (size_t i = 0; i < ids.size(); ++i) {
        const dicegp_chain_desc &d = c->hdesc[ids[i]];
        const int C = c->hC[d.gene];
        int tier = pick_tier(c, cls, ids[i], latency_mode);
        if (tier <= 4 && wide_eligible(c, ids[i], latency_mode)) tier = kWideTier;   // wide rejection-chain rounds
        else if (latency_mode && persistent && tier >= 5 && tier <= 7 && wide_big && wide_eligible(c, ids[i], latency_mode, tier_cluster(tier)))
        {
            tier = kWideTier + 1 + (tier - 5);                 // large genes: the same rounds on a cluster of 2 / 4 / 8 blocks
            // ... and of 16 (a non-portable cluster size) from 64 MB of W on: such a chain is bound by the likelihood
            // pass alone, which shrinks with the SMs it is spread over
            static const int64_t big16 = getenv("DICEGP_WIDE_CS16_BYTES") ? atoll(getenv("DICEGP_WIDE_CS16_BYTES")) : (64ll << 20);
            if (tier == kWideTier + 3 && big16 > 0 && chain_wbytes(c, ids[i]) >= big16 && wide_cluster_capacity(c, 16) > 0)
                tier = kWideTier + 4;
        }
        if (latency_mode == 2 && tier == 4) tier = 6;          // last stragglers: a cluster of 4 blocks each
        const int k = tier * kKinds + kind_of(C);
        by_group[k].push_back(ids[i]);
        slot_src[k].push_back((int32_t)i);
    }
    for (int k = kWideTier * kKinds; k < kGroups; ++k)
        if (!by_group[k].empty()) {
            int rc = ensure_tiles(c);
            if (rc != DICEGP_OK) return rc;
            break;
        }
    if (persistent) {
        // heavy chains first (cost ~ C * reads)
        for (int k = 0; k < kGroups; ++k) {
            std::vector<std::pair<int64_t, int32_t>> key;
            for (int32_t id : by_group[k]) key.push_back({-chain_wbytes(c, id), id});
            std::stable_sort(key.begin(), key.end());
            for (size_t i = 0; i < key.size(); ++i) by_group[k][i] = key[i].second;
        }
    }
    // flatten ids (and permuted host-stream offsets) into one device array
    std::vector<int32_t> flat;
    std::vector<int64_t> flat_doff;   // [0, n): delta offsets, [n, 2n): u offsets, kernel-slot order
    std::vector<int64_t> flat_uoff;
    size_t base[kGroups + 1] = {0};
    for (int k = 0; k < kGroups; ++k) {
        base[k] = flat.size();
        for (size_t i = 0; i < by_group[k].size(); ++i) {
            flat.push_back(by_group[k][i]);
            if (delta_off) {
                flat_doff.push_back((*delta_off)[slot_src[k][i]]);
                flat_uoff.push_back((int64_t)slot_src[k][i] * n_steps);
            }
        }
    }
    base[kGroups] = flat.size();
    DG_CUDA(c, c->dids.resize(flat.size()));
    DG_CUDA(c, cudaMemcpyAsync(c->dids.p, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (delta_off) {
        flat_doff.insert(flat_doff.end(), flat_uoff.begin(), flat_uoff.end());
        DG_CUDA(c, c->ddoff.resize(flat_doff.size()));
        DG_CUDA(c, cudaMemcpyAsync(c->ddoff.p, flat_doff.data(), flat_doff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    }
    DG_CUDA(c, c->dqueue.resize(kGroups));
    DG_CUDA(c, cudaMemsetAsync(c->dqueue.p, 0, kGroups * sizeof(int), c->stream));

    // steps bookkeeping for the stats (m_next before)
    c->hm_next_before.resize(c->n_chains);
    DG_CUDA(c, cudaMemcpyAsync(c->hm_next_before.data(), c->cm_next.p, c->n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));

#ifdef DG_PROFILE
    DG_CUDA(c, c->dprof.resize(16));
    DG_CUDA(c, cudaMemsetAsync(c->dprof.p, 0, 16 * sizeof(unsigned long long), c->stream));
    sa.prof = reinterpret_cast<const double *>(c->dprof.p);
#endif
    GeneTab gt = make_gene_tab(c);
    ChainTab ct = make_chain_tab(c);
    DG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    DG_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
    int64_t launches = 0;
    int next_stream = 0;
    // big tiers first: the 1-block-per-SM and streaming launches carry most of the work
    const int tier_order[kTiers] = {7, 6, 5, kWideTier + 4, kWideTier + 3, kWideTier + 2, kWideTier + 1, kWideTier, 4, 3, 2, 1, 0};
    for (int to = 0; to < kTiers; ++to) {
        for (int kind_i = 0; kind_i < kKinds; ++kind_i) {
            // many isoforms first: their steps are the longest, so their chains should not be the ones that start last
            // (in the straggler pass most of the chains that use all M steps are those; in pass 0 the order was
            // measured to make no difference)
            const int kind = kind_i == kKinds - 1 ? 0 : kKinds - 1 - kind_i;
            const int tier = tier_order[to];
            const int k = tier * kKinds + kind;
            const int n_ids = (int)by_group[k].size();
            if (n_ids == 0) continue;
            if (tier >= kWideTier) {
                int maxC = 2, maxTcw = 1;
                for (int32_t id : by_group[k]) {
                    maxC = std::max(maxC, c->hC[c->hdesc[id].gene]);
                    maxTcw = std::max(maxTcw, c->hdesc[id].Tc);
                }
                // cluster size: that of the pass (last stragglers), or the large gene's own tier
                const int wcs = tier == kWideTier ? wide_cs : (2 << (tier - kWideTier - 1));
                const WidePlan wp = wide_plan(c, kind == 0 ? maxC : kind + 1, maxTcw, wcs);
                if (tier > kWideTier) {
                    int rc = preallocate_trace_pages(c, by_group[k]);     // only the lead block writes the trace
                    if (rc != DICEGP_OK) return rc;
                }
                size_t fixed = 0;
                for (int32_t id : by_group[k]) {
                    const dicegp_chain_desc &d = c->hdesc[id];
                    fixed = std::max(fixed, wide_fixed_bytes(c->hC[d.gene], d.Tc, wp));
                }
                const size_t budget = wide_budget(c, wp);
                const size_t ring_bytes = (budget - fixed) & ~(size_t)127;
                const size_t smem = fixed + ring_bytes;
                cudaStream_t st = c->class_stream[next_stream];
                next_stream = (next_stream + 1) % kStreams;
                DG_CUDA(c, cudaStreamWaitEvent(st, c->ev_fork, 0));
                StreamArgs sak = sa;
                if (delta_off) {
                    sak.delta_off = c->ddoff.p + base[k];
                    sak.u_off = c->ddoff.p + flat.size() + base[k];
                }
                int grid = n_ids;                                      // (last-stragglers passes: the pages were pre-assigned above)
                int *queue = nullptr;
                if (persistent) {
                    grid = std::min(n_ids, wcs > 1 ? wide_cluster_capacity(c, wcs) : c->sm_count * wp.bps);
                    queue = c->dqueue.p + k;
                }
                grid *= wcs;
                wide_kernel_t kern = wide_kernel_for(kind, wcs > 1);
                DG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(c->smem_optin - c->static_smem_wide)));
                if (wcs > 8) DG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                if (getenv("DICEGP_VERBOSE"))
                    fprintf(stderr, "[dicegp]   launch wide C-kind %d: %d chains, %d + %d warps, cluster %d, %zu B smem, grid %d\n",
                            kind == 0 ? 0 : kind + 1, n_ids, wp.ncw, wp.npw, wcs, smem, grid);
                {
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof cfg);
                    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * (wp.ncw + wp.npw)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = wcs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = wcs > 1 ? 1 : 0;
                    const int32_t *ids_arg = c->dids.p + base[k];
                    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, gt, ct, sak, ids_arg, n_ids, n_steps, (int)ring_bytes, wp.npw, queue);
                    if (e == cudaSuccess) e = cudaGetLastError();
                    if (e != cudaSuccess) return c->cuda_fail(e, "wide chain kernel launch");
                }
                ++launches;
                ++c->total_launches;
                DG_CUDA(c, cudaEventRecord(c->ev_join[k], st));
                DG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
                continue;
            }
            // launch shape: resident tiers use the tier's block size; the streaming tier plans
            // (likelihood warps per chain, chains per SM) from the group's shapes so that as many
            // streaming warps as possible are resident: threads*k <= DG_LB_THREADS (registers),
            // smem*k <= what an SM offers, at most 4 chains per SM
            int threads = cls[tier].threads, bps = cls[tier].blocks_per_sm;
            int maxC = 1, maxTc = 1, maxS = 1;
            for (int32_t id : by_group[k]) {
                const dicegp_chain_desc &d = c->hdesc[id];
                maxC = std::max(maxC, c->hC[d.gene]);
                maxTc = std::max(maxTc, d.Tc);
                maxS = std::max(maxS, (c->hC[d.gene] * d.Tc + 31) / 32);
            }
            const int cs = tier_cluster(tier);
            if (tier == 4 && !getenv("DICEGP_TIER_THREADS")) {
                // every warp of a block streams W (the state warps join the likelihood pass after the prior terms),
                // so every warp owns a staging ring
                int best_score = -1;
                const int lb = kind != 0 ? DG_LB_THREADS_STAGED : DG_LB_THREADS;   // threads per SM the variant's registers allow
                for (int kk = 1; kk <= (latency_mode ? 1 : 4); ++kk) {
                    for (int nwl = 1; nwl <= lb / 32 - 1; ++nwl) {
                        const int thr = 32 * (maxS + nwl);
                        if (thr * kk > lb) continue;
                        size_t need = fixed_smem_bytes(maxC, maxTc, thr, true) + 1024;
                        if (kind != 0) need += (size_t)ring_warps(thr, maxC, maxTc) * dg::dg_stages(maxC) * dg::dg_stage_slot_bytes(maxC);
                        if (need * kk > 233472 || need - 1024 + 127 > c->smem_optin - c->static_smem) continue;
                        // measured per isoform count on the bench workload (scripts/class_timing.py): ten streaming warps
                        // per SM first, then three chains per SM rather than two or four (their serial phases overlap; a
                        // fourth chain only shrinks everybody's share of warps), then more warps per chain
                        const int score = latency_mode ? nwl
                                        : std::min(kk * ring_warps(thr, maxC, maxTc), 10) * 100 - 15 * std::abs(kk - 3) + nwl;
                        if (score > best_score) { best_score = score; threads = thr; bps = kk; }
                    }
                }
                if (best_score < 0) return c->fail(DICEGP_ERR_LIMIT, "chain has too many latent values for the streaming tier");
                // development override: DICEGP_PLAN_K<C> = "threads,blocks per SM" for the streamed launch of C isoforms
                char name[32];
                snprintf(name, sizeof name, "DICEGP_PLAN_K%d", kind == 0 ? 0 : kind + 1);
                if (const char *e = getenv(name)) sscanf(e, "%d,%d", &threads, &bps);
            }
            if (tier >= 5) {
                // cluster tier: as many likelihood warps per block as the staging ring allows
                int best = -1;
                const int lb = kind != 0 ? DG_LB_THREADS_STAGED : DG_LB_THREADS;
                for (int nwl = 2; nwl <= lb / 32 - 2; ++nwl) {
                    const int thr = 32 * (maxS + nwl);
                    if (thr > lb) break;
                    size_t need = fixed_smem_bytes(maxC, maxTc, thr, true, cs) + 1024;
                    if (kind != 0) need += (size_t)ring_warps(thr, maxC, maxTc) * dg::dg_stages(maxC) * dg::dg_stage_slot_bytes(maxC);
                    if (need > 233472 || need - 1024 + 127 > c->smem_optin - c->static_smem) break;
                    best = thr;
                }
                if (best < 0) return c->fail(DICEGP_ERR_LIMIT, "chain has too many latent values for the cluster tier");
                threads = best; bps = 1;
                int rc = preallocate_trace_pages(c, by_group[k]);
                if (rc != DICEGP_OK) return rc;
            }
            // dynamic shared memory of this launch: the largest need among its chains
            size_t smem = 0;
            for (int32_t id : by_group[k]) {
                const dicegp_chain_desc &d = c->hdesc[id];
                size_t need = fixed_smem_bytes(c->hC[d.gene], d.Tc, threads, tier >= 4, cs);
                if (cls[tier].w_in_smem) need += (size_t)chain_wbytes(c, id);
                else if (kind != 0) {
                    // per warp: kStages x C x 32 lanes x 16 bytes of cp.async staging
                    const int C = c->hC[d.gene];
                    need += (size_t)ring_warps(threads, C, d.Tc) * dg::dg_stages(C) * dg::dg_stage_slot_bytes(C);
                }
                smem = std::max(smem, need);
            }
            smem = (smem + 127) & ~(size_t)127;
            if (smem > c->smem_optin - c->static_smem) return c->fail(DICEGP_ERR_LIMIT, "chain needs more shared memory than the device offers");
            cudaStream_t st = c->class_stream[next_stream];
            next_stream = (next_stream + 1) % kStreams;
            DG_CUDA(c, cudaStreamWaitEvent(st, c->ev_fork, 0));
            StreamArgs sak = sa;
            if (delta_off) {
                sak.delta_off = c->ddoff.p + base[k];
                sak.u_off = c->ddoff.p + flat.size() + base[k];
            }
            int grid = n_ids;
            int *queue = nullptr;
            if (persistent) {
                grid = std::min(n_ids, c->sm_count * bps / cs);
                queue = c->dqueue.p + k;
            }
            grid *= cs;
            // the streamed tiers all use the cluster-capable variants (cluster size 1 for tier 4): their code
            // schedule is also the faster one for a single block (measured)
            if (getenv("DICEGP_VERBOSE"))
                fprintf(stderr, "[dicegp]   launch tier %d C-kind %d: %d chains, %d threads x %d blocks/SM (cluster %d), %zu B smem, grid %d\n",
                        tier, kind == 0 ? 0 : kind + 1, n_ids, threads, bps, cs, smem, grid);
            chain_kernel_t kern = chain_kernel_for(kind, cls[tier].w_in_smem, tier >= 4);
            DG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(c->smem_optin - c->static_smem)));
            {
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof cfg);
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                const int32_t *ids_arg = c->dids.p + base[k];
                cudaError_t e = cudaLaunchKernelEx(&cfg, kern, gt, ct, sak, ids_arg, n_ids, n_steps, c->use_tma, queue);
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e != cudaSuccess) return c->cuda_fail(e, "chain kernel launch");
            }
            ++launches;
            ++c->total_launches;
            DG_CUDA(c, cudaEventRecord(c->ev_join[k], st));
            DG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
        }
    }
    DG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return c->cuda_fail(e, "chain kernel");
    }
    float ms = 0.f;
    DG_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    // stats: steps advanced and read evaluations
    std::vector<int32_t> after(c->n_chains);
    DG_CUDA(c, cudaMemcpy(after.data(), c->cm_next.p, c->n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int64_t steps = 0, evals = 0;
    for (int32_t id : ids) {
        int64_t ds = (int64_t)after[id] - c->hm_next_before[id];
        int64_t init_eval = (c->hm_next_before[id] == 0 && after[id] > 0) ? 1 : 0;
        steps += ds;
        evals += (ds + init_eval) * chain_reads(c, id);
    }
    c->last_ms = ms; c->last_launches = launches; c->last_evals = evals; c->last_steps = steps;
#ifdef DG_PROFILE
    {
        unsigned long long h[16];
        DG_CUDA(c, cudaMemcpy(h, c->dprof.p, sizeof h, cudaMemcpyDeviceToHost));
        // general kernel: S, window, B2wait, decide, update, geweke; wide kernel: S, wait f, likelihood, wait
        // partials, decide + new state, wait state, rows (+ geweke), aux warp busy; h[11] = rounds (wide)
        if (h[11]) {
            double n = (double)h[11];
            fprintf(stderr, "[dg_prof] wide: steps=%llu rounds=%llu (%.2f steps/round) cycles/round:", h[10], h[11], h[10] / n);
            const char *nm[8] = {"S", "waitf", "lik", "waitp", "decide", "waits", "rows", "aux"};
            double tot = 0;
            for (int i = 0; i < 8; ++i) { fprintf(stderr, " %s=%.0f", nm[i], h[i] / n); if (i < 7) tot += h[i] / n; }
            fprintf(stderr, " total=%.0f | within lik (warp 0): barrier wait=%.0f task=%.0f normals=%.0f issue=%.0f flush=%.0f\n", tot, h[8] / n, h[9] / n, h[12] / n, h[13] / n, h[14] / n);
        } else {
            const char *nm[9] = {"S", "prior", "B2wait", "decide", "update", "geweke", "lik(warp0)", "-", "-"};
            double n = (double)std::max<unsigned long long>(h[10], 1);
            fprintf(stderr, "[dg_prof] steps=%llu cycles/step:", h[10]);
            double tot = 0;
            for (int i = 0; i < 7; ++i) { fprintf(stderr, " %s=%.0f", nm[i], h[i] / n); tot += h[i] / n; }
            fprintf(stderr, " total=%.0f\n", tot);
        }
    }
#endif
    return DICEGP_OK;
}

// Single-time-point chains (the pre-runs of get_psi, static mode) whose W fits shared memory: one
// warp per chain (dicegp_prerun.cuh), advanced by up to n_steps.
bool prerun_eligible(const dicegp_ctx *c, int chain, size_t *smem_bytes) {
    const dicegp_chain_desc &d = c->hdesc[chain];
    const int C = c->hC[d.gene];
    if (d.Tc != 1 || C < 2 || C > DG_MAXC || d.prop_factor_off < 0) return false;
    const size_t wd = (size_t)chain_wbytes(c, chain) / sizeof(double);
    // (the run-time-C group shares one launch: its scratch is sized for the largest isoform count)
    const size_t bytes = (wd + (size_t)dg::dg_prerun_scratch(kind_of(C) == 0 ? DG_MAXC : C)) * sizeof(double);
    if (bytes > 110 * 1024) return false;                                      // at least two chains per SM
    *smem_bytes = bytes;
    return true;
}

int run_prerun_warps(dicegp_ctx *c, const std::vector<int32_t> &ids, uint64_t seed, int n_steps) {
    c->chains_fresh = false;
    c->hm_next_before.resize(c->n_chains);
    DG_CUDA(c, cudaMemcpyAsync(c->hm_next_before.data(), c->cm_next.p, c->n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    std::vector<int32_t> by_kind[kKinds];
    for (int32_t id : ids) by_kind[kind_of(c->hC[c->hdesc[id].gene])].push_back(id);
    std::vector<int32_t> flat;
    size_t base[kKinds + 1];
    for (int k = 0; k < kKinds; ++k) { base[k] = flat.size(); flat.insert(flat.end(), by_kind[k].begin(), by_kind[k].end()); }
    base[kKinds] = flat.size();
    DG_CUDA(c, c->dids.resize(flat.size()));
    DG_CUDA(c, cudaMemcpyAsync(c->dids.p, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, c->dqueue.resize(kGroups));
    DG_CUDA(c, cudaMemsetAsync(c->dqueue.p, 0, kGroups * sizeof(int), c->stream));
    typedef void (*prerun_kernel_t)(GeneTab, ChainTab, uint64_t, const int32_t *, int, int *, int, int);
    static const prerun_kernel_t kernels[kKinds] = {
        dicegp_prerun_kernel<0>, dicegp_prerun_kernel<2>, dicegp_prerun_kernel<3>, dicegp_prerun_kernel<4>,
        dicegp_prerun_kernel<5>, dicegp_prerun_kernel<6>, dicegp_prerun_kernel<7>, dicegp_prerun_kernel<8>};
    for (int k = 0; k < kKinds; ++k)
        DG_CUDA(c, cudaFuncSetAttribute(kernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    GeneTab gt = make_gene_tab(c);
    ChainTab ct = make_chain_tab(c);
    DG_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    DG_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
    int64_t launches = 0, steps = 0, evals = 0;
    for (int k = kKinds - 1; k >= 0; --k) {                 // many isoforms first: fewest chains per SM, longest steps
        const int n_ids = (int)by_kind[k].size();
        if (n_ids == 0) continue;
        size_t wd = 0;
        int maxC = 2;
        for (int32_t id : by_kind[k]) {
            wd = std::max(wd, (size_t)chain_wbytes(c, id) / sizeof(double));
            maxC = std::max(maxC, c->hC[c->hdesc[id].gene]);
        }
        wd = (wd + 1) & ~(size_t)1;
        const size_t smem = (wd + (size_t)dg::dg_prerun_scratch(k == 0 ? DG_MAXC : maxC)) * sizeof(double);
        const int bps = (int)std::min<size_t>(32, 233472 / (smem + 1024));
        const int grid = std::min(n_ids, c->sm_count * std::max(bps, 1));
        cudaStream_t st = c->class_stream[k];
        DG_CUDA(c, cudaStreamWaitEvent(st, c->ev_fork, 0));
        kernels[k]<<<grid, 32, smem, st>>>(gt, ct, seed, c->dids.p + base[k], n_ids, c->dqueue.p + k, (int)wd, n_steps);
        DG_CUDA(c, cudaGetLastError());
        ++launches;
        ++c->total_launches;
        DG_CUDA(c, cudaEventRecord(c->ev_join[k], st));
        DG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
    }
    DG_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return c->cuda_fail(e, "pre-run kernel");
    }
    float ms = 0.f;
    DG_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    std::vector<int32_t> after(c->n_chains);
    DG_CUDA(c, cudaMemcpy(after.data(), c->cm_next.p, c->n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int32_t id : ids) {
        const int64_t ds = (int64_t)after[id] - c->hm_next_before[id];
        const int64_t init_eval = (c->hm_next_before[id] == 0 && after[id] > 0) ? 1 : 0;     // the start value
        steps += ds;
        evals += (ds + init_eval) * chain_reads(c, id);
    }
    c->last_ms = ms; c->last_launches = launches; c->last_evals = evals; c->last_steps = steps;
    return DICEGP_OK;
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

int dicegp_abi_version(void) { return DICEGP_ABI_VERSION; }

const char *dicegp_last_error(const dicegp_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

int dicegp_create(int device, dicegp_ctx **out) {
    if (!out) { g_last_error = "out is NULL"; return DICEGP_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_last_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return DICEGP_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_last_error = "device index out of range"; return DICEGP_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return DICEGP_ERR_CUDA; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return DICEGP_ERR_CUDA; }
    if (prop.major < 10) {
        g_last_error = "libdicegp is built for sm_100a (B200) only";
        return DICEGP_ERR_CUDA;
    }
    dicegp_ctx *c = new dicegp_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    {
        cudaFuncAttributes fa0, fa1;
        if (cudaFuncGetAttributes(&fa0, dicegp_chain_kernel<0, dg::W_DIRECT, false>) != cudaSuccess ||
            cudaFuncGetAttributes(&fa1, dicegp_chain_kernel<0, dg::W_RESIDENT, false>) != cudaSuccess) {
            g_last_error = std::string("kernel image not loadable on this device: ") + cudaGetErrorString(cudaGetLastError());
            delete c;
            return DICEGP_ERR_CUDA;
        }
        c->static_smem = std::max(fa0.sharedSizeBytes, fa1.sharedSizeBytes);
        cudaFuncAttributes fw;
        if (cudaFuncGetAttributes(&fw, dicegp_wide_kernel<0, 4>) != cudaSuccess) {
            g_last_error = std::string("kernel image not loadable on this device: ") + cudaGetErrorString(cudaGetLastError());
            delete c;
            return DICEGP_ERR_CUDA;
        }
        c->static_smem_wide = fw.sharedSizeBytes;
    }
    const char *no_tma = getenv("DICEGP_NO_TMA");
    c->use_tma = (no_tma && no_tma[0] == '1') ? 0 : 1;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) goto fail;
    for (int k = 0; k < kStreams; ++k)
        if (cudaStreamCreateWithFlags(&c->class_stream[k], cudaStreamNonBlocking) != cudaSuccess) goto fail;
    for (int k = 0; k < 128; ++k)
        if (cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) != cudaSuccess) goto fail;
    if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) goto fail;
    if (cudaEventCreate(&c->ev_t0) != cudaSuccess || cudaEventCreate(&c->ev_t1) != cudaSuccess) goto fail;
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) goto fail;
    if (c->dctr.resize(8) != cudaSuccess || cudaMemset(c->dctr.p, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) goto fail;
    *out = c;
    return DICEGP_OK;
fail:
    g_last_error = std::string("stream/event creation failed: ") + cudaGetErrorString(cudaGetLastError());
    delete c;
    return DICEGP_ERR_CUDA;
}

int dicegp_destroy(dicegp_ctx *c) {
    if (!c) return DICEGP_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->dC.release(); c->dncnt.release(); c->droff.release(); c->dwoff.release(); c->dlenoff.release();
    c->dtoff.release(); c->dwtoff.release(); c->dWt.release();
    c->dW.release(); c->dlen.release();
    c->cgene.release(); c->ct0.release(); c->cTc.release(); c->cM.release(); c->cinit.release();
    c->cgap.release(); c->cm_next.release(); c->ccnt.release(); c->cstate.release(); c->cm_ret.release();
    c->cflags.release(); c->ckinv.release(); c->cpf.release(); c->cjs.release(); c->cjc.release();
    c->cy0.release(); c->cptab_off.release(); c->cptab.release(); c->cpshift.release(); c->cpool_head.release();
    c->dgather.release(); c->cst_off.release(); c->cstream_id.release();
    c->cprior.release(); c->cpool.release(); c->ctrace.release(); c->cst.release();
    c->dids.release(); c->ddoff.release(); c->ddelta.release(); c->du.release(); c->dvar.release();
    c->dqueue.release(); c->deval.release(); c->deval_done.release(); c->dctr.release(); c->dprof.release();
    c->smean.release(); c->sci.release(); c->slik.release(); c->skeys.release(); c->sscr_off.release(); c->soff.release(); c->sids.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    for (int k = 0; k < 96; ++k)
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
    for (int k = 0; k < kStreams; ++k)
        if (c->class_stream[k]) cudaStreamDestroy(c->class_stream[k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return DICEGP_OK;
}

int dicegp_get_limits(const dicegp_ctx *ctx, dicegp_limits *out) {
    if (!out) return DICEGP_ERR_ARG;
    out->max_isoforms = DG_MAXC;
    out->max_times = DG_MAXT;
    out->abi_version = DICEGP_ABI_VERSION;
    out->sm_count = ctx ? ctx->sm_count : 0;
    out->max_latent = DG_LB_THREADS - 64;     // state warps (one lane per latent value) + two likelihood warps in one block
    out->reserved = 0;
    return DICEGP_OK;
}

int dicegp_upload_genes(dicegp_ctx *c, int32_t G, int32_t T, const int32_t *n_iso, const int32_t *n_reads,
                        const double *W, int64_t w_count, const double *len_iso, int64_t len_count) {
    if (!c) return DICEGP_ERR_ARG;
    if (G <= 0 || T <= 0 || !n_iso || !n_reads || !len_iso || (w_count > 0 && !W))
        return c->fail(DICEGP_ERR_ARG, "upload_genes: NULL or non-positive argument");
    if (T > DG_MAXT) return c->fail(DICEGP_ERR_LIMIT, "upload_genes: T beyond max_times");
    DG_CUDA(c, cudaSetDevice(c->device));
    c->G = 0; c->n_chains = 0;
    c->hC.assign(n_iso, n_iso + G);
    c->hncnt.assign(n_reads, n_reads + (size_t)G * T);
    c->hroff.assign((size_t)G * (T + 1), 0);
    c->htoff.assign((size_t)G * (T + 1), 0);
    c->hwtoff.assign(G, 0);
    c->tiles_valid = false;
    c->hwoff.assign(G, 0);
    c->hlenoff.assign(G, 0);
    int64_t w = 0, l = 0, wt_total = 0;
    for (int g = 0; g < G; ++g) {
        const int C = n_iso[g];
        if (C < 1 || C > DG_MAXC) return c->fail(DICEGP_ERR_LIMIT, "upload_genes: isoform count outside [1, max_isoforms]");
        if ((int64_t)C * T > DG_LB_THREADS - 64) return c->fail(DICEGP_ERR_LIMIT, "upload_genes: C*T beyond max_latent (dicegp_get_limits)");
        c->hwoff[g] = w; c->hlenoff[g] = l;
        int64_t r = 0;
        for (int t = 0; t < T; ++t) {
            const int n = n_reads[(size_t)g * T + t];
            if (n < 0) return c->fail(DICEGP_ERR_ARG, "upload_genes: negative read count");
            c->hroff[(size_t)g * (T + 1) + t] = (int32_t)r;
            r += n + (n & 1);
            if (r > 0x7fffffff) return c->fail(DICEGP_ERR_LIMIT, "upload_genes: more than 2^31 reads in one gene");
        }
        c->hroff[(size_t)g * (T + 1) + T] = (int32_t)r;
        {
            int64_t tiles = 0;                                   // 16-word tiles of the wide-round kernel
            for (int t = 0; t < T; ++t) {
                c->htoff[(size_t)g * (T + 1) + t] = (int32_t)tiles;
                const int64_t np = (c->hroff[(size_t)g * (T + 1) + t + 1] - c->hroff[(size_t)g * (T + 1) + t]) >> 1;
                tiles += (np + 15) >> 4;
            }
            c->htoff[(size_t)g * (T + 1) + T] = (int32_t)tiles;
            c->hwtoff[g] = wt_total;
            wt_total += tiles * C * 32;                          // doubles
        }
        w += (int64_t)C * r;
        l += (int64_t)C * T;
    }
    if (w != w_count) return c->fail(DICEGP_ERR_ARG, "upload_genes: w_count does not match the packed layout");
    if (l != len_count) return c->fail(DICEGP_ERR_ARG, "upload_genes: len_count does not match sum C*T");
    if (c->dC.resize(G) != cudaSuccess || c->dncnt.resize((size_t)G * T) != cudaSuccess ||
        c->droff.resize((size_t)G * (T + 1)) != cudaSuccess || c->dwoff.resize(G) != cudaSuccess ||
        c->dlenoff.resize(G) != cudaSuccess || c->dW.resize(std::max<int64_t>(w, 2)) != cudaSuccess ||
        c->dlen.resize(l) != cudaSuccess)
        return c->fail(DICEGP_ERR_NOMEM, "upload_genes: device allocation failed");
    DG_CUDA(c, cudaMemcpyAsync(c->dC.p, c->hC.data(), G * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, cudaMemcpyAsync(c->dncnt.p, c->hncnt.data(), (size_t)G * T * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, cudaMemcpyAsync(c->droff.p, c->hroff.data(), (size_t)G * (T + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, c->dtoff.resize((size_t)G * (T + 1)));
    DG_CUDA(c, c->dwtoff.resize(G));
    DG_CUDA(c, cudaMemcpyAsync(c->dtoff.p, c->htoff.data(), (size_t)G * (T + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, cudaMemcpyAsync(c->dwtoff.p, c->hwtoff.data(), G * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    c->wt_total = wt_total;
    DG_CUDA(c, cudaMemcpyAsync(c->dwoff.p, c->hwoff.data(), G * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, cudaMemcpyAsync(c->dlenoff.p, c->hlenoff.data(), G * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    // W goes up in 64 MB pieces, each submitted when the previous one has landed: another context
    // sampling on this GPU at the same time (PipelinedSampler) issues many small host->device copies,
    // which would otherwise queue behind one multi-GB transfer on the copy engine
    for (int64_t o = 0; o < w; o += (8 << 20)) {          // 64 MB pieces (~1.3 ms each): ~50 host syncs per 3 GB batch
        const int64_t cnt = std::min<int64_t>(w - o, 8 << 20);
        DG_CUDA(c, cudaMemcpyAsync(c->dW.p + o, W + o, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        DG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    DG_CUDA(c, cudaMemcpyAsync(c->dlen.p, len_iso, l * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d_bytes += (w + l) * (int64_t)sizeof(double) + ((int64_t)G * (2 * T + 2) ) * 4 + (int64_t)G * 16;
    c->G = G; c->T = T;
    return DICEGP_OK;
}

static int install_chains(dicegp_ctx *c, int32_t n, const dicegp_chain_desc *desc, const double *pool,
                          int64_t pool_count, bool pool_on_device) {
    c->n_chains = 0;
    c->hdesc.assign(desc, desc + n);
    c->hptab_off.assign(n, 0);
    c->hpshift.assign(n, 0);
    c->hst_off.assign(n, 0);
    std::vector<int64_t> &ptab = c->h_ptab;
    ptab.clear();
    std::vector<int32_t> &gene = c->h_i32[0], &t0 = c->h_i32[1], &Tc = c->h_i32[2], &M = c->h_i32[3], &ini = c->h_i32[4],
                         &gap = c->h_i32[5];
    std::vector<int64_t> &kinv = c->h_i64[0], &pf = c->h_i64[1], &js = c->h_i64[2], &jc = c->h_i64[3], &y0 = c->h_i64[4];
    std::vector<double> &prior = c->h_prior;
    for (std::vector<int32_t> *v : {&gene, &t0, &Tc, &M, &ini, &gap}) v->resize(n);
    for (std::vector<int64_t> *v : {&kinv, &pf, &js, &jc, &y0}) v->resize(n);
    prior.resize(n);
    int64_t tr = 0, st = 0, worst = 0;
    for (int i = 0; i < n; ++i) {
        const dicegp_chain_desc &d = desc[i];
        if (d.gene < 0 || d.gene >= c->G) return c->fail(DICEGP_ERR_ARG, "set_chains: gene index out of range");
        if (d.Tc < 1 || d.t0 < 0 || d.t0 + d.Tc > c->T) return c->fail(DICEGP_ERR_ARG, "set_chains: time range outside the gene");
        if (d.M < 1 || d.gap < 1 || d.initial < 0) return c->fail(DICEGP_ERR_ARG, "set_chains: M/initial/gap invalid");
        const int C = c->hC[d.gene];
        const int64_t need[4] = {d.kinv_off + (int64_t)d.Tc * d.Tc, d.jump_scale_off + C - 1, d.jump_const_off + C - 1,
                                 d.y0_off >= 0 ? d.y0_off + (int64_t)C * d.Tc : 0};
        if (d.kinv_off < 0 || d.jump_scale_off < 0 || d.jump_const_off < 0) return c->fail(DICEGP_ERR_ARG, "set_chains: negative pool offset");
        for (int k = 0; k < 4; ++k)
            if (need[k] > pool_count) return c->fail(DICEGP_ERR_ARG, "set_chains: pool offset beyond pool_count");
        if (d.prop_factor_off >= 0 && d.prop_factor_off + (int64_t)d.Tc * d.Tc > pool_count)
            return c->fail(DICEGP_ERR_ARG, "set_chains: prop_factor offset beyond pool_count");
        gene[i] = d.gene; t0[i] = d.t0; Tc[i] = d.Tc; M[i] = d.M; ini[i] = d.initial; gap[i] = d.gap;
        kinv[i] = d.kinv_off; pf[i] = d.prop_factor_off; js[i] = d.jump_scale_off; jc[i] = d.jump_const_off;
        y0[i] = d.y0_off; prior[i] = d.prior_const;
        c->hst_off[i] = st;
        const int64_t rowlen = 2 * (int64_t)C * d.Tc + 2;
        // pages of 2^shift rows, at most 1024; the first page is allocated here
        int shift = 0;
        while ((1 << shift) < d.M && shift < 10) ++shift;
        const int n_pages = (d.M + (1 << shift) - 1) >> shift;
        c->hpshift[i] = shift;
        c->hptab_off[i] = (int64_t)ptab.size();
        ptab.push_back(tr);
        for (int k = 1; k < n_pages; ++k) ptab.push_back(-1);
        tr += rowlen << shift;
        worst += (rowlen << shift) * n_pages;
        st += rowlen + dg::dg_geweke_state_len((C - 1) * d.Tc);     // resume state + running Geweke sums
    }
#define DG_RESIZE(buf, cnt) if ((buf).resize(cnt) != cudaSuccess) return c->fail(DICEGP_ERR_NOMEM, "set_chains: device allocation failed (" #buf ")")
    DG_RESIZE(c->cgene, n); DG_RESIZE(c->ct0, n); DG_RESIZE(c->cTc, n); DG_RESIZE(c->cM, n);
    DG_RESIZE(c->cinit, n); DG_RESIZE(c->cgap, n); DG_RESIZE(c->cm_next, n); DG_RESIZE(c->ccnt, n);
    DG_RESIZE(c->cstate, n); DG_RESIZE(c->cm_ret, n); DG_RESIZE(c->cflags, n);
    DG_RESIZE(c->ckinv, n); DG_RESIZE(c->cpf, n); DG_RESIZE(c->cjs, n); DG_RESIZE(c->cjc, n);
    DG_RESIZE(c->cy0, n); DG_RESIZE(c->cst_off, n); DG_RESIZE(c->cstream_id, n);
    DG_RESIZE(c->cprior, n); DG_RESIZE(c->cst, st);
    if (!pool_on_device) DG_RESIZE(c->cpool, std::max<int64_t>(pool_count, 1));
    DG_RESIZE(c->cptab_off, n); DG_RESIZE(c->cpshift, n); DG_RESIZE(c->cptab, ptab.size()); DG_RESIZE(c->cpool_head, 1);
    {
        // trace pool: first pages + room for growth, bounded by the budget
        // (cudaMemGetInfo goes through the resource manager and was seen to take tens of ms now and then on a box
        // whose GPUs are being monitored: only asked when no budget was set)
        int64_t budget = c->trace_budget_bytes;
        if (budget <= 0) {
            size_t free_b = 0, total_b = 0;
            DG_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
            free_b += c->ctrace.n * sizeof(double);                  // the old pool is reused / replaced
            budget = (int64_t)(free_b * 0.6);
        }
        int64_t cap = std::min<int64_t>(worst, std::max<int64_t>(budget / 8, tr));
        if (c->ctrace.resize(cap) != cudaSuccess) {
            char msg[200];
            snprintf(msg, sizeof msg, "set_chains: cannot allocate %.1f GB of trace memory (first pages need %.1f GB); use fewer chains per batch",
                     cap * 8.0 / 1e9, tr * 8.0 / 1e9);
            return c->fail(DICEGP_ERR_NOMEM, msg);
        }
        c->pool_cap = cap;
    }
#undef DG_RESIZE
    cudaStream_t s = c->stream;
#define DG_UP(dst, src, type) DG_CUDA(c, cudaMemcpyAsync((dst).p, (src).data(), n * sizeof(type), cudaMemcpyHostToDevice, s))
    DG_UP(c->cgene, gene, int32_t); DG_UP(c->ct0, t0, int32_t); DG_UP(c->cTc, Tc, int32_t); DG_UP(c->cM, M, int32_t);
    DG_UP(c->cinit, ini, int32_t); DG_UP(c->cgap, gap, int32_t);
    DG_UP(c->ckinv, kinv, int64_t); DG_UP(c->cpf, pf, int64_t); DG_UP(c->cjs, js, int64_t); DG_UP(c->cjc, jc, int64_t);
    DG_UP(c->cy0, y0, int64_t); DG_UP(c->cptab_off, c->hptab_off, int64_t); DG_UP(c->cst_off, c->hst_off, int64_t);
    DG_UP(c->cpshift, c->hpshift, int32_t);
    DG_CUDA(c, cudaMemcpyAsync(c->cptab.p, ptab.data(), ptab.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    {
        unsigned long long head = (unsigned long long)tr;
        DG_CUDA(c, cudaMemcpyAsync(c->cpool_head.p, &head, sizeof head, cudaMemcpyHostToDevice, s));
    }
    DG_UP(c->cprior, prior, double);
#undef DG_UP
    if (!pool_on_device && pool_count > 0)
        DG_CUDA(c, cudaMemcpyAsync(c->cpool.p, pool, pool_count * sizeof(double), cudaMemcpyHostToDevice, s));
    DG_CUDA(c, cudaMemsetAsync(c->cm_next.p, 0, n * sizeof(int32_t), s));
    DG_CUDA(c, cudaMemsetAsync(c->ccnt.p, 0, n * sizeof(int32_t), s));
    DG_CUDA(c, cudaMemsetAsync(c->cstate.p, 0, n * sizeof(int32_t), s));
    DG_CUDA(c, cudaMemsetAsync(c->cm_ret.p, 0, n * sizeof(int32_t), s));
    DG_CUDA(c, cudaMemsetAsync(c->cflags.p, 0, n * sizeof(int32_t), s));
    DG_CUDA(c, cudaStreamSynchronize(s));
    c->have_stream_id = false;
    c->n_chains = n;
    c->chains_fresh = true;
    return DICEGP_OK;
}

int dicegp_set_chains(dicegp_ctx *c, int32_t n, const dicegp_chain_desc *desc, const double *pool,
                      int64_t pool_count) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->G == 0) return c->fail(DICEGP_ERR_STATE, "set_chains: upload genes first");
    if (n <= 0 || !desc || (pool_count > 0 && !pool)) return c->fail(DICEGP_ERR_ARG, "set_chains: NULL or empty argument");
    DG_CUDA(c, cudaSetDevice(c->device));
    return install_chains(c, n, desc, pool, pool_count, false);
}

int dicegp_main_chains_from_prerun(dicegp_ctx *c, int32_t M, int32_t initial, int32_t gap, double theta1,
                                   const double *kinv, const double *prop_factor, double prior_const,
                                   double *var_out) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->G == 0 || c->n_chains != c->G * c->T)
        return c->fail(DICEGP_ERR_STATE, "main_chains_from_prerun: expects G*T finished pre-run chains");
    if (!kinv) return c->fail(DICEGP_ERR_ARG, "main_chains_from_prerun: kinv is NULL");
    DG_CUDA(c, cudaSetDevice(c->device));
    const int G = c->G, T = c->T;
    for (int i = 0; i < c->n_chains; ++i) {
        const dicegp_chain_desc &d = c->hdesc[i];
        if (d.gene != i / T || d.t0 != i % T || d.Tc != 1)
            return c->fail(DICEGP_ERR_STATE, "main_chains_from_prerun: chain g*T+j must be the pre-run of gene g, time j");
    }
    std::vector<int32_t> st(c->n_chains);
    DG_CUDA(c, cudaMemcpy(st.data(), c->cstate.p, c->n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int32_t s : st)
        if (s == 0) return c->fail(DICEGP_ERR_STATE, "main_chains_from_prerun: a pre-run chain has not finished");
    DG_CUDA(c, c->dvar.resize((size_t)G * DG_MAXC));
    {
        const int threads = 128, total = G * DG_MAXC;
        dicegp_prerun_var_kernel<<<(total + threads - 1) / threads, threads, 0, c->stream>>>(
            make_gene_tab(c), make_chain_tab(c), G, c->dvar.p);
        DG_CUDA(c, cudaGetLastError());
        ++c->total_launches;
    }
    std::vector<double> var((size_t)G * DG_MAXC);
    DG_CUDA(c, cudaMemcpyAsync(var.data(), c->dvar.p, var.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    if (var_out) memcpy(var_out, var.data(), var.size() * sizeof(double));
    // new chain set: pool = [kinv TT][pfac TT] + per gene [js C-1][jc C-1]
    const int TT = T * T;
    std::vector<double> pool(2 * (size_t)TT);
    memcpy(pool.data(), kinv, TT * sizeof(double));
    if (prop_factor) memcpy(pool.data() + TT, prop_factor, TT * sizeof(double));
    std::vector<dicegp_chain_desc> desc(G);
    for (int g = 0; g < G; ++g) {
        const int C = c->hC[g];
        dicegp_chain_desc &d = desc[g];
        d.gene = g; d.t0 = 0; d.Tc = T; d.M = M; d.initial = initial; d.gap = gap;
        d.kinv_off = 0; d.prop_factor_off = prop_factor ? TT : -1; d.y0_off = -1;
        d.prior_const = prior_const;
        d.jump_scale_off = (int64_t)pool.size();
        for (int k = 0; k < C - 1; ++k) pool.push_back(var[(size_t)g * DG_MAXC + k] * 5 / ((double)(T * C) * theta1));
        d.jump_const_off = (int64_t)pool.size();
        for (int k = 0; k < C - 1; ++k) {
            const double js = pool[d.jump_scale_off + k];
            pool.push_back(prior_const - (double)T * std::log(js) / 2.0);
        }
    }
    return install_chains(c, G, desc.data(), pool.data(), (int64_t)pool.size(), false);
}

int dicegp_run_host_stream(dicegp_ctx *c, int32_t n, const int32_t *chain_ids, int32_t n_steps,
                           const double *delta, const int64_t *delta_off, int64_t delta_count,
                           const double *u) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->n_chains == 0) return c->fail(DICEGP_ERR_STATE, "run_host_stream: no chains");
    if (n <= 0 || n_steps <= 0 || !chain_ids || !delta_off || !u || (delta_count > 0 && !delta))
        return c->fail(DICEGP_ERR_ARG, "run_host_stream: NULL or non-positive argument");
    DG_CUDA(c, cudaSetDevice(c->device));
    std::vector<int32_t> ids(chain_ids, chain_ids + n);
    std::vector<int64_t> doff(delta_off, delta_off + n);
    std::vector<char> seen(c->n_chains, 0);
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || ids[i] >= c->n_chains) return c->fail(DICEGP_ERR_ARG, "run_host_stream: chain id out of range");
        if (seen[ids[i]]) return c->fail(DICEGP_ERR_ARG, "run_host_stream: chain listed twice");
        seen[ids[i]] = 1;
        const dicegp_chain_desc &d = c->hdesc[ids[i]];
        const int64_t need = (int64_t)n_steps * (c->hC[d.gene] - 1) * d.Tc;
        if (doff[i] < 0 || doff[i] + need > delta_count) return c->fail(DICEGP_ERR_ARG, "run_host_stream: delta block outside delta_count");
    }
    DG_CUDA(c, c->ddelta.resize(std::max<int64_t>(delta_count, 1)));
    DG_CUDA(c, c->du.resize((size_t)n * n_steps));
    if (delta_count > 0)
        DG_CUDA(c, cudaMemcpyAsync(c->ddelta.p, delta, delta_count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DG_CUDA(c, cudaMemcpyAsync(c->du.p, u, (size_t)n * n_steps * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    StreamArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.mode = 0;
    sa.delta = c->ddelta.p;
    sa.u = c->du.p;
    return run_chains(c, ids, n_steps, sa, &doff, false);
}

int dicegp_run_device_stream(dicegp_ctx *c, uint64_t seed, const int64_t *stream_id) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->n_chains == 0) return c->fail(DICEGP_ERR_STATE, "run_device_stream: no chains");
    DG_CUDA(c, cudaSetDevice(c->device));
    for (int i = 0; i < c->n_chains; ++i)
        if (c->hdesc[i].prop_factor_off < 0)
            return c->fail(DICEGP_ERR_ARG, "run_device_stream: chain without prop_factor");
    if (stream_id) {
        DG_CUDA(c, cudaMemcpyAsync(c->cstream_id.p, stream_id, c->n_chains * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        c->have_stream_id = true;
    } else {
        c->have_stream_id = false;
    }
    std::vector<int32_t> ids;
    double fast_ms = 0.0;
    int64_t fast_launches = 0, fast_evals = 0, fast_steps = 0;
    {
        // single-time-point chains (get_psi's pre-runs, static mode) whose W fits shared memory take the
        // warp-per-chain kernel and run to their end in one launch (a step costs them ~2.5 us, so
        // even a chain that uses all M steps is short); DICEGP_PRERUN_KERNEL=0 sends them through
        // the general kernel; =1 forces the warp kernel.  Unset: the warp kernel needs a few chains per
        // SM to pay (a lone warp per chain is a throughput shape: 400 mixed chains ran 20 % slower
        // with it than with the general kernel, 10k chains 2.4 x faster)
        const char *e = getenv("DICEGP_PRERUN_KERNEL");
        const bool use_fast = !(e && e[0] == '0');
        std::vector<int32_t> fast;
        for (int i = 0; i < c->n_chains; ++i) {
            size_t bytes = 0;
            if (use_fast && prerun_eligible(c, i, &bytes)) fast.push_back(i);
            else ids.push_back(i);
        }
        if (!(e && e[0] == '1') && (int)fast.size() < 2 * c->sm_count) {
            ids.resize(c->n_chains);
            for (int i = 0; i < c->n_chains; ++i) ids[i] = i;
            fast.clear();
        }
        if (!fast.empty()) {
            int fast_M = 0;
            for (int32_t i : fast) fast_M = std::max(fast_M, c->hdesc[i].M);
            int rc = run_prerun_warps(c, fast, seed, fast_M);
            if (rc != DICEGP_OK) return rc;
            fast_ms = c->last_ms; fast_launches = c->last_launches; fast_evals = c->last_evals; fast_steps = c->last_steps;
            if (getenv("DICEGP_VERBOSE"))
                fprintf(stderr, "[dicegp] warp-per-chain kernel: %.1f ms, %lld steps, %zu chains\n", fast_ms, (long long)fast_steps, fast.size());
            if (ids.empty()) return DICEGP_OK;
        }
    }
    struct AddFast {                                             // the general path's stats + the fast path's
        dicegp_ctx *c; double ms; int64_t l, e, s;
        ~AddFast() { c->last_ms += ms; c->last_launches += l; c->last_evals += e; c->last_steps += s; }
    } add_fast{c, fast_ms, fast_launches, fast_evals, fast_steps};
    int max_M = 0;
    for (int32_t i : ids) max_M = std::max(max_M, c->hdesc[i].M);
    StreamArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.mode = 1;
    sa.seed = seed;
    // Pass 1 (throughput shapes): every chain, capped at kStragglerSteps.  Pass 2 (latency
    // shapes): the few chains still running -- they would otherwise crawl along with the two or
    // three likelihood warps of the throughput shape while the rest of the GPU idles.  The
    // stream is counter based, so a resumed chain continues bit-identically.
    // Pass 1 (throughput shapes): every chain, capped at 4096 steps.  Pass 2 (latency shapes,
    // one chain per SM with as many likelihood warps as fit): the ~10% of chains still running.
    // Otherwise a handful of non-converging chains would crawl along with two likelihood warps
    // each while the rest of the GPU idles.  The stream is counter based, so a resumed chain
    // continues bit-identically.  (An optional third pass on clusters of 4 blocks exists for
    // experiments -- DICEGP_STRAGGLER_STEPS=a,b -- but measured slower on the bench workload.)
    // Further passes when the stragglers are few (a rank of a multi-GPU run, a small batch): with fewer chains than
    // SMs what is left is bound by the chains that use all M steps, so the survivors of every 4096 more steps are
    // re-launched on thread-block clusters of 2 / 4 / 8 SMs per chain (wide-round kernel, DSMEM exchange) as soon as
    // that many SMs per chain are free; once the clusters have their full size the chains run to the end.
    // DICEGP_STRAGGLER_STEPS=a[,b] fixes the plan instead (a: cap of pass 0; b: cap of a pass 1 before the cluster pass).
    // Cap of pass 0, measured on the time-series workload (scripts/exp_r02i.sh): with 10k chains the throughput shapes
    // of the general kernel are as good as the wide rounds up to ~4096 steps; with few chains per SM the chains that
    // are still running soon leave SMs half empty (a throughput shape takes 5-8 us per step, so 4096 steps of ONE
    // chain are 20-33 ms) and the wide-round kernel (one chain per SM, 3-5 us per step) should take over early:
    // 1250 chains 137 -> 101 ms, 2500 chains 202 -> 165 ms, 5000 chains 347 -> 323 ms.
    int cap0 = ids.size() <= 3000 ? 1024 : (ids.size() <= 7000 ? 2048 : 4096), cap1 = 0;
    const bool fixed_plan = getenv("DICEGP_STRAGGLER_STEPS") != nullptr;
    if (const char *e = getenv("DICEGP_STRAGGLER_STEPS")) sscanf(e, "%d,%d", &cap0, &cap1);
    if (const char *e = getenv("DICEGP_CAP0")) cap0 = atoi(e);            // cap of pass 0 only (the staged plan stays)
    if (cap0 <= 0 || max_M <= cap0) return run_chains(c, ids, max_M, sa, nullptr, true, 0);
    double ms = 0.0;
    int64_t launches = 0, evals = 0, steps = 0;
    std::vector<int32_t> todo = ids;
    int prev = 0;                                               // steps every chain in `todo` has done
    for (int pass = 0; !todo.empty() && prev < max_M; ++pass) {
        int cap, mode;
        const int n = (int)todo.size();
        if (pass == 0) { cap = cap0; mode = 0; }
        else if (fixed_plan) {
            if (pass == 1 && cap1 > cap0 && cap1 < max_M) { cap = cap1; mode = 1; }
            else { cap = max_M; mode = (cap1 > cap0 && cap1 < max_M) ? 2 : 1; }
        } else if (n > 4 * c->sm_count) { cap = max_M; mode = 1; }       // throughput bound: one chain per SM to the end
        else {
            mode = n <= wide_cluster_capacity(c, 2) ? 2 : 1;              // clusters as soon as two SMs per chain are free
            cap = n <= wide_cluster_capacity(c, 8) ? max_M : std::min(max_M, prev + 4096);
        }
        int rc = run_chains(c, todo, cap - prev, sa, nullptr, true, mode);
        if (rc != DICEGP_OK) return rc;
        prev = cap;
        if (getenv("DICEGP_VERBOSE"))
            fprintf(stderr, "[dicegp] pass %d (cap %d, mode %d): %.1f ms, %lld steps, %zu chains\n", pass, cap,
                    mode, c->last_ms, (long long)c->last_steps, todo.size());
        ms += c->last_ms; launches += c->last_launches; evals += c->last_evals; steps += c->last_steps;
        if (cap >= max_M) break;
        std::vector<int32_t> st(c->n_chains);
        DG_CUDA(c, cudaMemcpy(st.data(), c->cstate.p, c->n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost));
        std::vector<int32_t> rest;
        for (int32_t id : todo) if (st[id] == DICEGP_CHAIN_RUNNING) rest.push_back(id);
        todo.swap(rest);
    }
    c->last_ms = ms; c->last_launches = launches; c->last_evals = evals; c->last_steps = steps;
    return DICEGP_OK;
}

int dicegp_dump_stream(dicegp_ctx *c, int32_t chain, uint64_t seed, int64_t stream_id, int32_t step0, int32_t n_steps,
                       double *delta, double *u) {
    if (!c) return DICEGP_ERR_ARG;
    if (chain < 0 || chain >= c->n_chains) return c->fail(DICEGP_ERR_ARG, "dump_stream: chain id out of range");
    if (step0 < 0 || n_steps <= 0 || !delta || !u) return c->fail(DICEGP_ERR_ARG, "dump_stream: NULL buffer or empty step range");
    const dicegp_chain_desc &d = c->hdesc[chain];
    if (d.prop_factor_off < 0) return c->fail(DICEGP_ERR_ARG, "dump_stream: chain without prop_factor");
    DG_CUDA(c, cudaSetDevice(c->device));
    const size_t ctm = (size_t)(c->hC[d.gene] - 1) * d.Tc;
    DG_CUDA(c, c->dgather.resize((size_t)n_steps * (ctm + 1)));
    double *dd = c->dgather.p, *du = dd + (size_t)n_steps * ctm;
    dicegp_stream_dump_kernel<<<1, 128, 0, c->stream>>>(make_gene_tab(c), make_chain_tab(c), chain, seed, stream_id, step0,
                                                        n_steps, dd, du);
    DG_CUDA(c, cudaGetLastError());
    ++c->total_launches;
    if (ctm > 0) DG_CUDA(c, cudaMemcpyAsync(delta, dd, (size_t)n_steps * ctm * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    DG_CUDA(c, cudaMemcpyAsync(u, du, (size_t)n_steps * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    return DICEGP_OK;
}

int dicegp_eval_loglik(dicegp_ctx *c, int32_t gene, int32_t t0, int32_t Tc, const double *y, double *lik) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->G == 0) return c->fail(DICEGP_ERR_STATE, "eval_loglik: no genes uploaded");
    if (gene < 0 || gene >= c->G || t0 < 0 || Tc <= 0 || t0 + Tc > c->T || !y || !lik)
        return c->fail(DICEGP_ERR_ARG, "eval_loglik: gene / time range outside the batch, or NULL buffer");
    DG_CUDA(c, cudaSetDevice(c->device));
    const int C = c->hC[gene];
    int max_n = 1;
    for (int t = t0; t < t0 + Tc; ++t) max_n = std::max(max_n, c->hncnt[(size_t)gene * c->T + t]);
    // enough blocks to cover the reads four deep, at most two waves of the device over the time points
    int nblk = (max_n + 4 * DG_EVAL_THREADS - 1) / (4 * DG_EVAL_THREADS);
    nblk = std::max(1, std::min(nblk, std::max(1, 2 * c->sm_count / Tc)));
    const size_t need = (size_t)C * Tc + (size_t)Tc * nblk + Tc;      // y | partial sums | lik
    DG_CUDA(c, c->deval.resize(need));
    if (c->deval_done.n < (size_t)c->T) {
        DG_CUDA(c, c->deval_done.resize(c->T));
        DG_CUDA(c, cudaMemsetAsync(c->deval_done.p, 0, c->deval_done.n * sizeof(unsigned), c->stream));
    }
    double *dy = c->deval.p, *dpart = dy + (size_t)C * Tc, *dlik = dpart + (size_t)Tc * nblk;
    DG_CUDA(c, cudaMemcpyAsync(dy, y, (size_t)C * Tc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dicegp_loglik_kernel<<<dim3(nblk, Tc), DG_EVAL_THREADS, 0, c->stream>>>(make_gene_tab(c), gene, t0, dy, Tc, dpart,
                                                                            c->deval_done.p, dlik);
    DG_CUDA(c, cudaGetLastError());
    ++c->total_launches;
    DG_CUDA(c, cudaMemcpyAsync(lik, dlik, (size_t)Tc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    return DICEGP_OK;
}

int dicegp_chain_status(dicegp_ctx *c, int32_t n, const int32_t *chain_ids, int32_t *m_ret, int32_t *n_rows,
                        int32_t *cnt, int32_t *state, int32_t *flags) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->n_chains == 0) return c->fail(DICEGP_ERR_STATE, "chain_status: no chains");
    if (n <= 0 || !chain_ids) return c->fail(DICEGP_ERR_ARG, "chain_status: NULL or empty id list");
    DG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->n_chains;
    std::vector<int32_t> a(N), b(N), s(N), f(N), mn(N);
    DG_CUDA(c, cudaMemcpy(a.data(), c->cm_ret.p, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    DG_CUDA(c, cudaMemcpy(b.data(), c->ccnt.p, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    DG_CUDA(c, cudaMemcpy(s.data(), c->cstate.p, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    DG_CUDA(c, cudaMemcpy(f.data(), c->cflags.p, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    DG_CUDA(c, cudaMemcpy(mn.data(), c->cm_next.p, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const int id = chain_ids[i];
        if (id < 0 || id >= N) return c->fail(DICEGP_ERR_ARG, "chain_status: chain id out of range");
        if (m_ret) m_ret[i] = a[id];
        if (cnt) cnt[i] = b[id];
        if (state) state[i] = s[id];
        if (flags) flags[i] = f[id];
        if (n_rows) n_rows[i] = s[id] == DICEGP_CHAIN_CONVERGED ? a[id] : (s[id] == DICEGP_CHAIN_EXHAUSTED ? c->hdesc[id].M : mn[id]);
    }
    return DICEGP_OK;
}

int dicegp_download_trace(dicegp_ctx *c, int32_t chain, int32_t first_row, int32_t n_rows, double *psi,
                          double *y, double *pro, double *lik) {
    if (!c) return DICEGP_ERR_ARG;
    if (chain < 0 || chain >= c->n_chains) return c->fail(DICEGP_ERR_ARG, "download_trace: chain id out of range");
    const dicegp_chain_desc &d = c->hdesc[chain];
    if (first_row < 0 || n_rows < 0 || first_row + n_rows > d.M) return c->fail(DICEGP_ERR_ARG, "download_trace: row range outside the trace");
    if (n_rows == 0) return DICEGP_OK;
    DG_CUDA(c, cudaSetDevice(c->device));
    {
        // rows beyond those written have no page: bound the range by the rows the chain holds
        int32_t written = 0;
        int rc = dicegp_chain_status(c, 1, &chain, nullptr, &written, nullptr, nullptr, nullptr);
        if (rc != DICEGP_OK) return rc;
        if (first_row + n_rows > written) return c->fail(DICEGP_ERR_ARG, "download_trace: row range beyond the rows written so far");
    }
    const int CT = c->hC[d.gene] * d.Tc;
    const size_t rowlen = 2 * (size_t)CT + 2;
    DG_CUDA(c, c->dgather.resize((size_t)n_rows * rowlen));
    dicegp_gather_rows_kernel<<<std::min(n_rows, 1184), 128, 0, c->stream>>>(make_chain_tab(c), chain, (int)rowlen,
                                                                            first_row, n_rows, c->dgather.p);
    DG_CUDA(c, cudaGetLastError());
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    const double *src = c->dgather.p;
    const size_t pitch = rowlen * sizeof(double);
    if (psi) DG_CUDA(c, cudaMemcpy2D(psi, CT * sizeof(double), src, pitch, CT * sizeof(double), n_rows, cudaMemcpyDeviceToHost));
    if (y) DG_CUDA(c, cudaMemcpy2D(y, CT * sizeof(double), src + CT, pitch, CT * sizeof(double), n_rows, cudaMemcpyDeviceToHost));
    if (pro) DG_CUDA(c, cudaMemcpy2D(pro, sizeof(double), src + 2 * CT, pitch, sizeof(double), n_rows, cudaMemcpyDeviceToHost));
    if (lik) DG_CUDA(c, cudaMemcpy2D(lik, sizeof(double), src + 2 * CT + 1, pitch, sizeof(double), n_rows, cudaMemcpyDeviceToHost));
    return DICEGP_OK;
}

int dicegp_posterior_summary(dicegp_ctx *c, int32_t n, const int32_t *chain_ids, double *psi_mean, double *psi_ci,
                             const int64_t *out_off, double *lik_marg) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->n_chains == 0) return c->fail(DICEGP_ERR_STATE, "posterior_summary: no chains");
    if (n <= 0 || !chain_ids || !psi_mean || !psi_ci || !out_off || !lik_marg)
        return c->fail(DICEGP_ERR_ARG, "posterior_summary: NULL or empty argument");
    DG_CUDA(c, cudaSetDevice(c->device));
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        const int id = chain_ids[i];
        if (id < 0 || id >= c->n_chains) return c->fail(DICEGP_ERR_ARG, "posterior_summary: chain id out of range");
        const dicegp_chain_desc &d = c->hdesc[id];
        total = std::max<int64_t>(total, out_off[i] + (int64_t)c->hC[d.gene] * d.Tc);
    }
    // Launch groups by rows kept: the radix select holds one column in shared memory, so
    // short chains (the bulk) run 8 blocks per SM and only the long ones need a whole SM.
    std::vector<int32_t> ids(chain_ids, chain_ids + n), nr(n), mr(n);
    {
        int rc = dicegp_chain_status(c, n, ids.data(), mr.data(), nr.data(), nullptr, nullptr, nullptr);
        if (rc != DICEGP_OK) return rc;
    }
    // classes 0-3 by rows written; class 4: chains that keep more rows after the burn-in than one column of
    // keys fits shared memory (any --mcmc max_run works: those select in a global scratch row instead)
    constexpr int kClasses = 5;
    const int bounds[4] = {3072, 7168, 14336, 0x7fffffff};
    const int64_t smem_rows = (int64_t)(c->smem_optin - 4096) / (int64_t)sizeof(unsigned long long);
    std::vector<int32_t> order;            // positions grouped by size class
    int gstart[kClasses + 1] = {0, 0, 0, 0, 0, 0}, gmax[kClasses] = {1, 1, 1, 1, 1};
    for (int k = 0; k < kClasses; ++k) {
        gstart[k] = (int)order.size();
        for (int i = 0; i < n; ++i) {
            // the keys are the rows kept after the burn-in int(m/4) (run_utils.py:125)
            const int keep = nr[i] - (mr[i] > 0 ? mr[i] >> 2 : 0);
            const bool huge = keep > smem_rows;
            const int lo = k == 0 ? -1 : bounds[k - 1];
            const bool mine = k == 4 ? huge : (!huge && nr[i] > lo && nr[i] <= bounds[k]);
            if (mine) { order.push_back(i); gmax[k] = std::max(gmax[k], keep); }
        }
    }
    gstart[kClasses] = (int)order.size();
    std::vector<int32_t> ids_sorted(n);
    std::vector<int64_t> off_sorted(n), scr_sorted(n);
    int64_t scr_total = 0;                 // transposed copy of the kept rows: (CT + 1) columns of `keep` doubles per chain
    for (int i = 0; i < n; ++i) {
        ids_sorted[i] = chain_ids[order[i]]; off_sorted[i] = out_off[order[i]];
        const dicegp_chain_desc &d = c->hdesc[ids_sorted[i]];
        const int keep = std::max(0, nr[order[i]] - (mr[order[i]] > 0 ? mr[order[i]] >> 2 : 0));
        scr_sorted[i] = scr_total;
        scr_total += (int64_t)keep * ((int64_t)c->hC[d.gene] * d.Tc + 1);
    }
    // (grown with 25 % headroom: the kept rows differ from batch to batch by a few per cent, and a cudaFree + cudaMalloc of
    // several GB inside a step costs more than the summary kernels)
    if (((size_t)scr_total > c->skeys.n && c->skeys.resize((size_t)(scr_total + scr_total / 4 + 1)) != cudaSuccess) ||
        c->sscr_off.resize(n) != cudaSuccess)
        return c->fail(DICEGP_ERR_NOMEM, "posterior_summary: device allocation failed (transposed trace scratch)");
    DevBuf<double> &dmean = c->smean, &dci = c->sci, &dlik = c->slik;
    DevBuf<int64_t> &doff = c->soff;
    DevBuf<int32_t> &dids = c->sids;
    if (dmean.resize(total) != cudaSuccess || dci.resize(2 * total) != cudaSuccess || dlik.resize(n) != cudaSuccess ||
        doff.resize(n) != cudaSuccess || dids.resize(n) != cudaSuccess)
        return c->fail(DICEGP_ERR_NOMEM, "posterior_summary: device allocation failed");
    int rc = DICEGP_OK;
    cudaError_t e;
    std::vector<double> lik_sorted(n);
    do {
        if ((e = cudaMemcpyAsync(doff.p, off_sorted.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dids.p, ids_sorted.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(c->sscr_off.p, scr_sorted.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) break;
        if ((e = cudaFuncSetAttribute(dicegp_summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->smem_optin - 4096))) != cudaSuccess) break;
        // the size classes run side by side (the few long chains keep only some SMs busy, for as long
        // as the thousands of short ones take all together): longest first, one stream each.
        // (Measured and dropped: fetching four trace columns per sweep over the rows -- a quarter of the
        // DRAM traffic, but four columns of keys in shared memory leave too few blocks per SM; and
        // finishing a selection early once its bucket holds <= 32 keys -- no gain.)
        if ((e = cudaEventRecord(c->ev_fork, c->stream)) != cudaSuccess) break;
        for (int k = kClasses - 1; k >= 0 && e == cudaSuccess; --k) {
            const int cnt = gstart[k + 1] - gstart[k];
            if (cnt == 0) continue;
            const size_t sm = std::max<size_t>(k == 4 ? 0 : (size_t)gmax[k] * sizeof(unsigned long long),
                                               32 * (DG_SUM_TILE_COLS + 1) * sizeof(double));
            cudaStream_t st = c->class_stream[k];
            if ((e = cudaStreamWaitEvent(st, c->ev_fork, 0)) != cudaSuccess) break;
            // the long classes are few chains of many rows: eight blocks per chain, each a share of the columns
            const int parts = k >= 2 ? 8 : 1;
            dicegp_summary_kernel<<<cnt * parts, DG_SUM_THREADS, sm, st>>>(make_gene_tab(c), make_chain_tab(c), dids.p + gstart[k],
                                                                           dmean.p, dci.p, doff.p + gstart[k], dlik.p + gstart[k],
                                                                           reinterpret_cast<double *>(c->skeys.p), c->sscr_off.p + gstart[k],
                                                                           k == 4 ? 0 : 1, parts);
            ++c->total_launches;
            if ((e = cudaGetLastError()) != cudaSuccess) break;
            if ((e = cudaEventRecord(c->ev_join[k], st)) != cudaSuccess) break;
            e = cudaStreamWaitEvent(c->stream, c->ev_join[k], 0);
        }
        if (e != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(psi_mean, dmean.p, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(psi_ci, dci.p, 2 * total * sizeof(double), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(lik_sorted.data(), dlik.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
        e = cudaStreamSynchronize(c->stream);
    } while (0);
    if (e != cudaSuccess) rc = c->cuda_fail(e, "posterior_summary");
    else {
        for (int i = 0; i < n; ++i) lik_marg[order[i]] = lik_sorted[i];
        c->d2h_bytes += (3 * total + n) * (int64_t)sizeof(double);
    }
    return rc;
}

int dicegp_timer_start(dicegp_ctx *c) {
    if (!c) return DICEGP_ERR_ARG;
    DG_CUDA(c, cudaSetDevice(c->device));
    DG_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
    return DICEGP_OK;
}

int dicegp_timer_stop(dicegp_ctx *c, double *ms) {
    if (!c || !ms) return DICEGP_ERR_ARG;
    DG_CUDA(c, cudaSetDevice(c->device));
    DG_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
    DG_CUDA(c, cudaEventSynchronize(c->ev_t1));
    float f = 0.f;
    DG_CUDA(c, cudaEventElapsedTime(&f, c->ev_t0, c->ev_t1));
    *ms = f;
    return DICEGP_OK;
}

int dicegp_counters(dicegp_ctx *c, int64_t *kernel_launches, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!c) return DICEGP_ERR_ARG;
    if (kernel_launches) *kernel_launches = c->total_launches;
    if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
    return DICEGP_OK;
}

int dicegp_work_counters(dicegp_ctx *c, int64_t *out, int32_t n_out, int32_t reset) {
    if (!c) return DICEGP_ERR_ARG;
    if (!out || n_out < 0) return c->fail(DICEGP_ERR_ARG, "work_counters: NULL buffer");
    DG_CUDA(c, cudaSetDevice(c->device));
    unsigned long long h[8];
    DG_CUDA(c, cudaMemcpyAsync(h, c->dctr.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    if (reset) DG_CUDA(c, cudaMemsetAsync(c->dctr.p, 0, sizeof h, c->stream));
    DG_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n_out && i < 8; ++i) out[i] = (int64_t)h[i];
    return DICEGP_OK;
}

int dicegp_set_prerun_chains(dicegp_ctx *c, double theta1, double var0, int32_t M, int32_t initial, int32_t gap) {
    if (!c) return DICEGP_ERR_ARG;
    if (c->G == 0) return c->fail(DICEGP_ERR_STATE, "set_prerun_chains: upload genes first");
    if (!(theta1 > 0) || !(var0 > 0)) return c->fail(DICEGP_ERR_ARG, "set_prerun_chains: theta1 and var0 must be positive");
    DG_CUDA(c, cudaSetDevice(c->device));
    // K = [[theta1]] (GP_K of one time point): inverse, determinant and svd factor are scalars
    const double two_pi_log = std::log(2.0 * 3.14159265358979323846);
    std::vector<double> pool;
    pool.push_back(1.0 / theta1);                       // kinv
    pool.push_back(std::sqrt(theta1));                  // z * sqrt(theta1) ~ N(0, K)
    const double prior_const = -(two_pi_log + std::log(theta1)) / 2.0;
    int64_t js_off[DG_MAXC + 1], jc_off[DG_MAXC + 1];
    for (int C = 0; C <= DG_MAXC; ++C) js_off[C] = jc_off[C] = -1;
    const int G = c->G, T = c->T;
    std::vector<dicegp_chain_desc> &desc = c->h_desc_tmp;
    desc.resize((size_t)G * T);
    for (int g = 0; g < G; ++g) {
        const int C = c->hC[g];
        if (js_off[C] < 0) {
            const double cov = theta1 * var0 * 5 / (1 * C * theta1);      // model_GP.py:328 with T = 1
            js_off[C] = (int64_t)pool.size();
            for (int k = 0; k < C - 1; ++k) pool.push_back(var0 * 5 / (1 * C * theta1));
            jc_off[C] = (int64_t)pool.size();
            for (int k = 0; k < C - 1; ++k) pool.push_back(-(two_pi_log + std::log(cov)) / 2.0);
        }
        for (int j = 0; j < T; ++j) {
            dicegp_chain_desc &d = desc[(size_t)g * T + j];
            d.gene = g; d.t0 = j; d.Tc = 1; d.M = M; d.initial = initial; d.gap = gap;
            d.kinv_off = 0; d.prop_factor_off = 1; d.jump_scale_off = js_off[C]; d.jump_const_off = jc_off[C];
            d.y0_off = -1; d.prior_const = prior_const;
        }
    }
    return install_chains(c, G * T, desc.data(), pool.data(), (int64_t)pool.size(), false);
}

int dicegp_set_trace_budget(dicegp_ctx *c, int64_t bytes) {
    if (!c) return DICEGP_ERR_ARG;
    if (bytes < 0) return c->fail(DICEGP_ERR_ARG, "set_trace_budget: negative budget");
    c->trace_budget_bytes = bytes;
    return DICEGP_OK;
}

int dicegp_last_run_stats(dicegp_ctx *c, double *kernel_ms, int64_t *launches, int64_t *read_evals, int64_t *steps) {
    if (!c) return DICEGP_ERR_ARG;
    if (kernel_ms) *kernel_ms = c->last_ms;
    if (launches) *launches = c->last_launches;
    if (read_evals) *read_evals = c->last_evals;
    if (steps) *steps = c->last_steps;
    return DICEGP_OK;
}

}  // extern "C"
