// dicegp_chain.cuh -- the per-chain MH loop (one thread block per chain), kernel v2.
//
// 1. Exact speculative rounds.  A Metropolis-Hastings step either keeps Y_now or moves to
//    Y_try, so the states the chain can be in after D steps form a binary tree with
//    J = 2^D - 1 proposal nodes, all of which are known BEFORE the first decision because
//    the proposal increments and uniforms come from a stream that does not depend on the
//    chain (model_GP.py:329,351).  One round evaluates the posterior of all J candidate
//    states in ONE pass over the read matrix W (J dot products per read instead of 1),
//    then walks the tree with the ordinary accept tests.  D steps are committed per round;
//    the arithmetic of every committed step is identical to running the steps one by one.
//    W traffic per step drops to 1/D and so does the serial latency per step.
// 2. Warp specialisation.  STATE warps hold one thread per (node, isoform, time) latent
//    value and own the serial part (proposal, softmax / length reweighting, GP prior,
//    decisions, trace rows).  LIKELIHOOD warps own fixed read slices, accumulate prod(x_n)
//    as (mantissa, exponent) per node, and refill the random stream ring while the state
//    warps decide.  Three block barriers per round: B1 (f ready), B2 (partial likelihoods
//    ready), B3 (decisions known).
#pragma once
#include "dicegp_kernels.cuh"

#ifdef DG_PROFILE
#define DG_TICK(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); dg_prof[i] += t__ - dg_t0; dg_t0 = t__; } } while (0)
#else
#define DG_TICK(i) do { } while (0)
#endif

namespace dg {

// named barriers: 0 = the whole block, 2 = state warps, 3 = likelihood warps
__device__ __forceinline__ void dg_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void dg_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_state(int n) { asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory"); }
__device__ __forceinline__ void bar_lik(int n) { asm volatile("bar.sync 3, %0;" ::"r"(n) : "memory"); }

// ---- thread-block cluster helpers (large genes: one chain over several SMs) ---------------
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the shared memory of block `rank` of the cluster (DSMEM), same offset as `p` here
__device__ __forceinline__ void st_cluster(double *p, unsigned rank, double v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(dg_smem_u32(p)), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cluster(int *p, unsigned rank, int v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(dg_smem_u32(p)), "r"(rank));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

// mantissa in [1,2) of a positive normal double whose high word is `hi`
__device__ __forceinline__ double dg_mant(double x, int hi) {
    return __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
}

// How a likelihood warp gets at W:
//   W_RESIDENT  the chain's W sits in shared memory (staged once per launch by a TMA bulk copy)
//   W_STAGED    W streams from L2/HBM through a per-warp ring of cp.async stages (each lane
//               copies exactly the 16-byte words it will consume: no barrier, deep prefetch)
//   W_DIRECT    plain global loads (isoform counts too large for the staging ring)
enum { W_RESIDENT = 0, W_STAGED = 1, W_DIRECT = 2 };
// share of the likelihood items a state warp takes, in sixteenths of a likelihood warp's share
#ifndef DG_STATE_LIK_WEIGHT
#define DG_STATE_LIK_WEIGHT 12
#endif
// Shape of a likelihood item, per isoform count C:
//   words   16-byte words (read pairs) a lane handles per item.  The words of a lane are independent
//           dependency chains in one basic block, which the scheduler interleaves (a likelihood
//           warp is bound by instruction latency, not by a throughput limit), and their reads
//           share one exponent extraction.
//   stages  depth of the per-warp cp.async ring of the streamed tiers (stages-1 items in flight);
//           any depth >= 2 (ring slots are counted, not masked).
// Measured on 592-chain groups of the bench workload (scripts/stream_sweep.py); DG_ITEM_WORDS /
// DG_STAGE_BYTES / DG_MIN_STAGES override the table for experiments.
__host__ __device__ constexpr int dg_item_words(int C) {
#ifdef DG_ITEM_WORDS
    return DG_ITEM_WORDS;
#else
    return C <= 2 ? 4 : (C <= 8 ? 2 : 1);
#endif
}
__host__ __device__ constexpr int dg_stage_slot_bytes(int C) { return C * 512 * dg_item_words(C); }
__host__ __device__ constexpr int dg_stages(int C) {
#ifdef DG_STAGE_BYTES
    return DG_STAGE_BYTES / dg_stage_slot_bytes(C) < DG_MIN_STAGES ? DG_MIN_STAGES
         : (DG_STAGE_BYTES / dg_stage_slot_bytes(C) > 12 ? 12 : DG_STAGE_BYTES / dg_stage_slot_bytes(C));
#else
    return C <= 3 ? 4 : (C <= 5 ? 3 : (C <= 8 ? 2 : 4));     // 16 / 12 / 12 / 15 / 12 / 14 / 16 KB per warp for C = 2..8
#endif
}

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dg_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Likelihood pass of one warp.  The chain's reads are cut into ITEMS of 32 consecutive read
// pairs of one time point (cum[t] = first item of time t); a warp takes the contiguous run of
// items [q0, q1) (fixed for the chain), so it changes time point (and reloads f) only a few times per pass.
//   x_n^(j) = sum_k W[n,k] f^(j)[k]                            model_GP.py:338
// W block of a time point: isoform-major, row k = npairs double2 (two reads per 16-byte
// word; consecutive lanes take consecutive words: conflict-free LDS.128 / coalesced
// LDG.128).  f of the J candidates sits in shared memory as ff[t][k][JP] and is read as
// broadcast LDS.128.  prod_n x_n is carried per candidate as a mantissa product and an
// exponent sum; anything that is not a positive normal double raises `bad` (the round is
// then redone with real logarithms, LOGPATH).
// W_STAGED: the warp's ring is a CYCLIC stream over its items -- after the last item of a pass the
// first items of the NEXT pass are already on their way (every pass reads the same items), so a
// round does not start with an empty pipeline and the copies overlap the serial part of the round.
// ring_pos: -1 before the first pass of a chain (the pass primes the ring), then the slot of the
// next item; the caller drains the ring (cp.async.wait_group 0) when the chain leaves the block.
template <int KC, int J, int WMODE, bool LOGPATH>
__device__ __forceinline__ void lik_pass(const double *__restrict__ Wc, int C, int Tc, const int *rel, const int *ncnt,
                                         const int *cum, const double *ff, double *stage, int lane, int q0, int q1,
                                         int &ring_pos, double (&mm)[J], int (&ee)[J], unsigned &bad, double (&ls)[J]) {
    constexpr int JP = (J + 1) & ~1;
    constexpr int IW = dg_item_words(KC > 0 ? KC : DG_MAXC), IP = 32 * IW;   // 16-byte words per lane, read pairs per item
    constexpr int kStages = dg_stages(KC > 0 ? KC : DG_MAXC);
    const int kmax = KC > 0 ? KC : C;
    if (q0 >= q1) return;
    double2 *mystage = reinterpret_cast<double2 *>(stage) + lane;     // [stage][k][word][32 lanes]
    // issue cursor (W_STAGED): item qi of time point ti, lane's first pair ip of inp, source word isrc
    int qi = q0, ti = 0, ip = 0, inp = 0;
    const double2 *isrc = nullptr;
    auto issue = [&](int slot) {
        if (isrc == nullptr || qi >= cum[ti + 1]) {
            while (qi >= cum[ti + 1]) ++ti;
            const int r0 = rel[ti];
            inp = (rel[ti + 1] - r0) >> 1;
            ip = (qi - cum[ti]) * IP + lane;
            isrc = reinterpret_cast<const double2 *>(Wc + (size_t)C * r0) + ip;
        }
#pragma unroll
        for (int w = 0; w < IW; ++w) {
            if (ip + 32 * w < inp) {
#pragma unroll
                for (int k = 0; k < kmax; ++k)
                    cp_async16(mystage + ((slot * kmax + k) * IW + w) * 32, isrc + 32 * w + (size_t)k * inp);
            }
        }
        ++qi; ip += IP; isrc += IP;
        if (qi == q1) { qi = q0; ti = 0; isrc = nullptr; }               // the stream is cyclic
        cp_async_commit();
    };
    int islot = 0, cslot = 0;                                        // ring slots of the next issue / the next item
    if (WMODE == W_STAGED) {
        if (ring_pos < 0) {                                          // first pass of the chain: prime the ring
#pragma unroll
            for (int s = 0; s < kStages - 1; ++s) issue(s);
            ring_pos = 0;
        } else {                                                     // kStages-1 items are in flight: move the cursor past them
            qi = q0 + (kStages - 1) % (q1 - q0);
        }
        cslot = ring_pos;
        islot = cslot + kStages - 1;
        if (islot >= kStages) islot -= kStages;
    }
#ifndef DG_FREG_MAX
#define DG_FREG_MAX 12
#endif
    constexpr bool FREG = (KC > 0 && KC * J <= (J > 3 ? 18 : DG_FREG_MAX));    // f of the time point held in registers
    constexpr int FR = FREG ? KC : 1;
    int it = 0, q = q0, tt = 0;
    while (q < q1) {
        while (q >= cum[tt + 1]) ++tt;
        const int c0 = cum[tt], c1 = min(cum[tt + 1], q1);
        const int r0 = rel[tt], npairs = (rel[tt + 1] - r0) >> 1, n = ncnt[tt];
        int p = (q - c0) * IP + lane;
        const double2 *Wt = reinterpret_cast<const double2 *>(Wc + (size_t)C * r0) + p;
        const double *ffcol = ff + (size_t)tt * C * JP;
        double fr[FR][J];
        if (FREG) {
#pragma unroll
            for (int k = 0; k < FR; ++k) {
#pragma unroll
                for (int j = 0; j < J; ++j) fr[k][j] = ffcol[k * JP + j];
            }
        }
#pragma unroll((WMODE == W_STAGED && J <= 3 && IW == 1) ? 2 : 1)
        for (; q < c1; ++q, ++it, p += IP, Wt += IP) {
            if (WMODE == W_STAGED) {
                issue(islot);
                islot = (islot + 1 == kStages) ? 0 : islot + 1;
                cp_async_wait<kStages - 1>();
            }
            if (p < npairs) {
                const double2 *st = mystage + (cslot * kmax * IW) * 32;
                // the IW words of this lane are independent dependency chains in one basic block; a word
                // past the end of the time point recomputes word 0 (valid address) and is dropped below
                bool val[IW];
                double x0[IW][J], x1[IW][J];
#pragma unroll
                for (int w = 0; w < IW; ++w) {
                    val[w] = (p + 32 * w) < npairs;
#pragma unroll
                    for (int j = 0; j < J; ++j) { x0[w][j] = 0.0; x1[w][j] = 0.0; }
                }
#pragma unroll
                for (int k = 0; k < kmax; ++k) {
                    double2 wv[IW];
#pragma unroll
                    for (int w = 0; w < IW; ++w) {
                        const int wo = val[w] ? w : 0;
                        wv[w] = (WMODE == W_STAGED) ? st[(k * IW + wo) * 32] : Wt[(size_t)k * npairs + 32 * wo];
                    }
                    double fj[JP];
                    if (!FREG) {
                        const double2 *f2 = reinterpret_cast<const double2 *>(ffcol + k * JP);
#pragma unroll
                        for (int u = 0; u < JP / 2; ++u) { const double2 t = f2[u]; fj[2 * u] = t.x; fj[2 * u + 1] = t.y; }
                    }
#pragma unroll
                    for (int w = 0; w < IW; ++w) {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const double fv = FREG ? fr[k < FR ? k : 0][j] : fj[j];
                            x0[w][j] = fma(wv[w].x, fv, x0[w][j]);
                            x1[w][j] = fma(wv[w].y, fv, x1[w][j]);
                        }
                    }
                }
                if (LOGPATH) {
#pragma unroll
                    for (int w = 0; w < IW; ++w) {
                        if (val[w]) {
                            const bool v1 = (2 * (p + 32 * w) + 1) < n;      // the pad slot of an odd block
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                ls[j] += log(x0[w][j]);
                                if (v1) ls[j] += log(x1[w][j]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        // the reads of this lane's words are multiplied first: scaling by powers of two is
                        // exact, so mant(product) carries the same bits as the product of the mantissas; a
                        // product that leaves the positive normal range (a zero, negative or non-finite x,
                        // or an under/overflow) sends the round to the log path
                        double pr2 = 1.0;
                        int sg = 0;
#pragma unroll
                        for (int w = 0; w < IW; ++w) {
                            const bool v1 = (2 * (p + 32 * w) + 1) < n;      // the pad slot of an odd block
                            const double a0 = val[w] ? x0[w][j] : 1.0;
                            const double a1 = (val[w] && v1) ? x1[w][j] : 1.0;
                            sg |= __double2hiint(a0) | __double2hiint(a1);       // two negative x would hide each other
                            pr2 = (w == 0) ? a0 * a1 : pr2 * (a0 * a1);
                        }
                        const int h = __double2hiint(pr2);
                        bad |= ((unsigned)(h - 0x00100000) >= 0x7fe00000u) ? 1u : 0u;
                        bad |= (unsigned)sg >> 31;
                        ee[j] += (h >> 20) - 1023;
                        mm[j] *= dg_mant(pr2, h);
                    }
                }
            }
            if (WMODE == W_STAGED) cslot = (cslot + 1 == kStages) ? 0 : cslot + 1;
            if (!LOGPATH && (it & 255) == 255) {               // keep the running products in range
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int g = __double2hiint(mm[j]);
                    ee[j] += (g >> 20) - 1023;
                    mm[j] = dg_mant(mm[j], g);
                }
            }
        }
    }
    if (WMODE == W_STAGED) ring_pos = cslot;
}

// Slow path (some x_n is 0 / negative / non-finite): the same sums with real logarithms so
// that -inf and NaN propagate like np.log(...).sum() does (model_GP.py:344).
template <int KC, int J, int WMODE>
__device__ __noinline__ void lik_warp_pass_log(const double *Wc, const int *rel, const int *ncnt, const int *cum,
                                               const double *ff, double *stage, double *red, int C, int Tc, int q0, int q1,
                                               int &ring_pos, int lane, int wl, int slots, int cs) {
    double mm[J], ls[J];
    int ee[J];
    unsigned bad = 0u;
#pragma unroll
    for (int j = 0; j < J; ++j) { mm[j] = 1.0; ls[j] = 0.0; ee[j] = 0; }
    lik_pass<KC, J, WMODE, true>(Wc, C, Tc, rel, ncnt, cum, ff, stage, lane, q0, q1, ring_pos, mm, ee, bad, ls);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        double v = ls[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) {
            if (cs == 1) red[j * slots + wl] = v;
            else for (int r = 0; r < cs; ++r) st_cluster(red + j * slots + wl, r, v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Geweke convergence rule over trace rows [0, m)          model_GP.py:156-185, :369-391
// Series = Psi_all[:m, c, t] for c < C-1 (columns 0..CTm-1 of a trace row); window A = rows
// [0, int(0.1 m)), window B = rows [int(0.5 m), m); converged when every column has
// |mean_A - mean_B| / sqrt(var_A + var_B) <= 2 (population variances, np.var).
//
// The windows only grow and slide between checks, so the chain keeps, per column, running
// sums over both windows of d = x - K and d^2 (K = the column's value in row int(0.5 m) of
// the first check: a post-burn-in value, which keeps var = Q/n - (S/n)^2 free of
// cancellation) and a check touches only the rows that entered or left a window
// (~1.6 gap rows) instead of all 0.6 m rows twice.  Lanes walk columns (coalesced), warps
// walk rows.  The sums live in global memory behind the chain's resume state, so a chain
// cut into several launches sees the same operations as an uncut one.
//   gst[0] rows in window A, gst[1] / gst[2] first / end row of window B, gst[3] K is set,
//   then K, S_A, Q_A, S_B, Q_B (CTm doubles each).
// Returns 1 to every thread of the block when all columns pass.
// ---------------------------------------------------------------------------------------
// numpy's pairwise sum (np.add.reduce of a 1-D series) of n copies of v: blocks of <= 128 values with 8
// accumulators (all equal here, so one is carried: their tree sum 8 r is exact), split in halves above that.
__device__ __noinline__ double dg_np_const_sum(double v, int n) {
    auto block = [&](int k) {
        if (k < 8) {
            double r = 0.0;
            for (int i = 0; i < k; ++i) r += v;
            return r;
        }
        double r = v;
        int i = 8;
        for (; i < k - (k % 8); i += 8) r += v;
        double res = 8.0 * r;
        for (; i < k; ++i) res += v;
        return res;
    };
    if (n <= 128) return block(n);
    int f_n[32], f_state[32], fs = 0, rs = 0;
    double res[34];
    f_n[0] = n; f_state[0] = 0; fs = 1;
    while (fs > 0) {
        const int k = fs - 1;
        if (f_n[k] <= 128) { res[rs++] = block(f_n[k]); --fs; continue; }
        int n2 = f_n[k] / 2;
        n2 -= n2 % 8;
        if (f_state[k] == 0) { f_state[k] = 1; f_n[fs] = n2; f_state[fs] = 0; ++fs; }
        else if (f_state[k] == 1) { f_state[k] = 2; f_n[fs] = f_n[k] - n2; f_state[fs] = 0; ++fs; }
        else { const double r = res[rs - 1], l = res[rs - 2]; rs -= 2; res[rs++] = l + r; --fs; }
    }
    return res[0];
}
// np.mean and np.var (two-pass, population) of n copies of v
__device__ __noinline__ void dg_np_const_stats(double v, int n, double &mean, double &var) {
    mean = dg_np_const_sum(v, n) / (double)n;
    const double e = v - mean;
    var = dg_np_const_sum(__dmul_rn(e, e), n) / (double)n;
}

__host__ __device__ inline int dg_geweke_state_len(int CTm) { return 4 + 5 * CTm; }

// Running sums of d = x - K and d^2 over rows first + i (i = warp, warp + nwarps, ... < n) of one column, added
// (or, SUB, removed) in row order.  Eight rows are loaded at a time: their page-table and value loads are
// independent, so a lone warp (the warp-per-chain kernels) pays the L2 latency once per eight rows, not twice
// per row; the arithmetic and its order are unchanged.
template <bool SUB>
__device__ __forceinline__ void gw_accum(const TraceView &tv, int first, int n, int warp, int nwarps, int col, double K,
                                         double &s, double &q) {
    int i = warp;
    for (; i + 7 * nwarps < n; i += 8 * nwarps) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(tv.row(first + i + u * nwarps) + col);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double d = v[u] - K;
            if (!SUB) { s += d; q = fma(d, d, q); } else { s -= d; q = fma(-d, d, q); }
        }
    }
    for (; i < n; i += nwarps) {
        const double d = __ldcg(tv.row(first + i) + col) - K;
        if (!SUB) { s += d; q = fma(d, d, q); } else { s -= d; q = fma(-d, d, q); }
    }
}

// Runs on the `nwarps` warps (index `warp`) that synchronise on named barrier `bar_id`.
__device__ __noinline__ int geweke_converged(double *gw, double *gst, int CTm, int nwarps, int warp, int lane,
                                             TraceView tv, int m, int bar_id = 0) {
    const int nbar = nwarps * 32;
    if (m <= 0) return 1;                                // empty series: Z is NaN, which is not > 2
    const int GC = CTm < DG_GW_COLS ? (CTm > 0 ? CTm : 1) : DG_GW_COLS;   // scratch columns per warp
    const int nA = (int)(0.1 * (double)m);
    const int iB = (int)(0.5 * (double)m);
    const int nB = m - iB;
    const int nA0 = (int)gst[0];
    int iB0 = (int)gst[1], mB0 = (int)gst[2];
    const bool haveK = gst[3] != 0.0;
    const bool resetB = iB >= mB0;                        // first check: window B starts empty at iB
    if (resetB) { iB0 = iB; mB0 = iB; }
    const int addA = nA - nA0, addB = m - mB0, subB = iB - iB0;
    double *Kc = gst + 4, *SA = Kc + CTm, *QA = SA + CTm, *SB = QA + CTm, *QB = SB + CTm;
    double *g0 = gw, *g1 = gw + nwarps * GC;
    int ok = 1;
    for (int c0 = 0; c0 < CTm; c0 += DG_GW_COLS) {
        const int col = c0 + lane;
        const bool act = col < CTm;
        double K = 0.0;
        if (act) K = haveK ? Kc[col] : __ldcg(tv.row(iB) + col);
        // ---- window A: rows [nA0, nA) enter ----
        double s = 0.0, q = 0.0;
        if (act) gw_accum<false>(tv, nA0, addA, warp, nwarps, col, K, s, q);
        if (lane < GC) { g0[warp * GC + lane] = s; g1[warp * GC + lane] = q; }
        dg_bar(bar_id, nbar);
        double sa = 0.0, qa = 0.0;
        if (warp == 0 && act) {
            if (nA0 > 0) { sa = SA[col]; qa = QA[col]; }
            for (int w = 0; w < nwarps; ++w) { sa += g0[w * GC + lane]; qa += g1[w * GC + lane]; }
        }
        dg_bar(bar_id, nbar);
        // ---- window B: rows [mB0, m) enter, rows [iB0, iB) leave ----
        s = 0.0; q = 0.0;
        if (act) {
            gw_accum<false>(tv, mB0, addB, warp, nwarps, col, K, s, q);
            gw_accum<true>(tv, iB0, subB, warp, nwarps, col, K, s, q);
        }
        if (lane < GC) { g0[warp * GC + lane] = s; g1[warp * GC + lane] = q; }
        dg_bar(bar_id, nbar);
        if (warp == 0 && act) {
            double sb = 0.0, qb = 0.0;
            if (!resetB) { sb = SB[col]; qb = QB[col]; }
            for (int w = 0; w < nwarps; ++w) { sb += g0[w * GC + lane]; qb += g1[w * GC + lane]; }
            Kc[col] = K; SA[col] = sa; QA[col] = qa; SB[col] = sb; QB[col] = qb;
            const double meanA = sa / (double)nA, meanB = sb / (double)nB;
            double va = qa / (double)nA - meanA * meanA, vb = qb / (double)nB - meanB * meanB;
            if (va < 0.0) va = 0.0;                       // rounding of an (almost) constant window
            if (vb < 0.0) vb = 0.0;
            double den = sqrt(va + vb), dmean = meanA - meanB;
            if (sa == 0.0 && qa == 0.0 && sb == 0.0 && qb == 0.0) {
                // a series that is constant over both windows (a chain that never moved: -inf likelihood,
                // SURVEY Appendix B #9).  numpy's two-pass variance of n equal values is the square of the
                // rounding error of its pairwise mean -- usually tiny but not zero, and then Z <= sqrt(2)
                // and the reference calls the chain converged; only when both means are exact is Z None.
                // Reproduce numpy's arithmetic for this case (O(n) adds; it ends the chain at the first check).
                double mA, vA, mB, vB;
                dg_np_const_stats(K, nA, mA, vA);
                dg_np_const_stats(K, nB, mB, vB);
                den = sqrt(vA + vB);
                dmean = mA - mB;
            }
            // reference: Z None (den == 0) or Z > 2 -> not converged; NaN compares false
            if (den == 0.0) ok = 0;
            else if (fabs(dmean) / den > 2.0) ok = 0;
        }
        dg_bar(bar_id, nbar);
    }
    // every column was judged by warp 0; its verdict goes to all warps through the (now free) scratch
    if (warp == 0) {
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) { gst[0] = (double)nA; gst[1] = (double)iB; gst[2] = (double)m; gst[3] = 1.0; gw[0] = (double)ok; }
    }
    dg_bar(bar_id, nbar);
    ok = (int)gw[0];
    dg_bar(bar_id, nbar);
    return ok;
}

// Out-of-line copies of the long math routines used on the round loop: the loop body has
// to stay small enough for the instruction cache (the serial part runs in one warp, which
// cannot hide instruction-fetch latency).
__device__ __noinline__ double dg_log(double x) { return log(x); }

// Fill stream ring entries for steps [s0, s1).  A ring row holds CT+2 doubles: the proposal
// increments delta[c,t] (c < C-1), log(u) at [CT] and the proposal log-density Q at [CT+1].
//   device stream: Philox normals z (Box-Muller, both outputs used), then
//                  delta = sqrt(jump_scale[c]) * (z[c,:] @ A_K)
//   host stream:   delta and u come from the uploaded chunk
//   Q = sum_c ( jump_const[c] - (d_c K^-1 d_c) / (2 jump_scale[c]) ),  d = Y_now - Y_try = -delta
//       (model_GP.py:330-331: Q_now and Q_try are the same number; it only enters the MH
//       exponent as ((P_try + Q) - P_now) - Q, so it is formed here, off the serial path)
// Called by `nprod` threads (index `pt`) which must ALL call it: they synchronise on the named
// barrier `bar_id` (0 with nprod = blockDim.x is __syncthreads).
__device__ __noinline__ void produce_stream(bool dev_stream, int s0, int s1, int pt, int nprod, int bar_id,
                                            double *zring, double *dring, int RING, int C, int Tc,
                                            const double *pfac, const double *sqs, const double *kinv,
                                            const double *jcon, const double *ijs, uint64_t seed, int64_t sid,
                                            const double *dblk, const double *ublk, int m_start) {
    const int CT = C * Tc, CTm = CT - Tc, RL = CT + 2;
    if (dev_stream) {
        const int npair = (CTm + 1) >> 1;
        for (int s = s0; s < s1; ++s) {
            double *z = zring + (s % DG_STREAM_BLOCK) * RL;
            for (int i = pt; i < npair; i += nprod) {
                double z0, z1;
                dg_normal_pair(seed, sid, (uint32_t)s, (uint32_t)i, z0, z1);
                z[2 * i] = z0;
                if (2 * i + 1 < CTm) z[2 * i + 1] = z1;
            }
            if (pt == nprod - 1) dring[(s % RING) * RL + CT] = dg_log(dg_uniform(seed, sid, (uint32_t)s));
        }
        dg_bar(bar_id, nprod);
        for (int s = s0; s < s1; ++s) {
            const double *z = zring + (s % DG_STREAM_BLOCK) * RL;
            double *d = dring + (s % RING) * RL;
            for (int i = pt; i < CTm; i += nprod) {
                const int c = i / Tc, t = i - c * Tc;
                double a = 0.0;
                for (int j = 0; j < Tc; ++j) a = fma(z[c * Tc + j], pfac[j * Tc + t], a);
                d[i] = sqs[c] * a;
            }
        }
    } else {
        for (int s = s0; s < s1; ++s) {
            double *d = dring + (s % RING) * RL;
            const double *src = dblk + (size_t)(s - m_start) * CTm;
            for (int i = pt; i < CTm; i += nprod) d[i] = src[i];
            if (pt == nprod - 1) d[CT] = dg_log(ublk[s - m_start]);
        }
    }
    dg_bar(bar_id, nprod);
    // per (step, c): d K^-1 d into the z ring (free now)
    const int nq = (s1 - s0) * (C - 1);
    for (int q = pt; q < nq; q += nprod) {
        const int s = s0 + q / (C - 1), c = q - (s - s0) * (C - 1);
        const double *d = dring + (s % RING) * RL + c * Tc;
        double acc = 0.0;
        for (int t = 0; t < Tc; ++t) {
            double r = 0.0;
            for (int j = 0; j < Tc; ++j) r = fma(d[j], kinv[j * Tc + t], r);
            acc += r * d[t];
        }
        zring[(s % DG_STREAM_BLOCK) * RL + c] = jcon[c] - (acc * ijs[c]) / 2.0;
    }
    dg_bar(bar_id, nprod);
    for (int s = s0 + pt; s < s1; s += nprod) {
        double Q = 0.0;
        for (int c = 0; c < C - 1; ++c) Q += zring[(s % DG_STREAM_BLOCK) * RL + c];
        dring[(s % RING) * RL + CT + 1] = Q;
    }
}

// Trace page of row `step` (first touch allocates from the pool).  Returns -1 when the
// pool is exhausted.  Rare: one call per 2^pshift steps.
__device__ __noinline__ long long trace_page(int64_t *ptab, int step, int pshift, int rowlen,
                                             unsigned long long *pool_head, int64_t pool_cap) {
    long long pg = __ldcg(ptab + (step >> pshift));
    if (pg < 0) {
        const unsigned long long need = (unsigned long long)rowlen << pshift;
        const unsigned long long off = atomicAdd(pool_head, need);
        if (off + need <= (unsigned long long)pool_cap) {
            pg = (long long)off;
            ptab[step >> pshift] = pg;
        }
    }
    return pg;
}

// ---------------------------------------------------------------------------------------
// one chain, up to n_steps steps.  `slot` indexes the host stream blocks.
// ---------------------------------------------------------------------------------------
template <int KC, int WMODE, int J, bool CLUSTER>
__device__ void run_chain(const GeneTab &gt, const ChainTab &ct, const StreamArgs &sa, int chain,
                          int slot, int n_steps, double *smem, int use_tma, int cs_rt, int crank_rt) {
    const int cs = CLUSTER ? cs_rt : 1, crank = CLUSTER ? crank_rt : 0;   // compile-time 1 / 0 for ordinary launches
    constexpr int JP = (J + 1) & ~1;
    constexpr int IPAIRS = 32 * dg_item_words(KC > 0 ? KC : DG_MAXC);   // read pairs per likelihood item
    constexpr int DT = J >= 4 ? 3 : (J >= 2 ? 2 : 1);     // levels of the speculation tree (J <= 7 nodes, heap order)
    static_assert(J >= 1 && J <= 7, "1..7 candidate nodes per round");
    constexpr int SB = DG_STREAM_BLOCK;                   // stream steps produced per batch
    constexpr int RING = dg_ring_rows(J);                 // ring of 2-3 batches
    static_assert(RING >= SB + 2 * J - 1, "the refilled batch must not overlap the rows of the next round");
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;

    const int g = ct.gene[chain], t0 = ct.t0[chain];
    const int C = KC > 0 ? KC : gt.C[g];
    const int Tc = ct.Tc[chain], T = gt.T;
    const int CT = C * Tc, CTm = (C - 1) * Tc, RL = CT + 2;     // RL: doubles per stream ring row
    const int S = (CT + 31) >> 5, nstate = S << 5, NWL = nwarps - S;
    // a cluster of cs blocks works on ONE chain: every block runs the same serial part (same
    // inputs, same instructions, same results), the reads are split over all likelihood warps
    // of the cluster, and the per-warp partial likelihoods are exchanged through distributed
    // shared memory once per round.  Only block 0 writes the trace and the chain status.
    const int SLOTS = cs > 1 ? DG_CLUSTER_SLOTS : DG_MAXWARPS;
    const bool lead = crank == 0;
    const SmLayout L = dg_layout(C, Tc, nwarps, J, cs);
    double *ycand = smem + L.ycand, *psic = smem + L.psic, *gs = smem + L.gs, *ex = smem + L.ex;
    double *pr = smem + L.pr, *ff = smem + L.ff, *zring = smem + L.zring, *dring = smem + L.dring;
    double *kinv = smem + L.kinv, *pfac = smem + L.pfac, *ijs = smem + L.ijs, *jcon = smem + L.jcon, *sqs = smem + L.sqs;
    double *red = smem + L.red, *bc = smem + L.bc, *gw = smem + L.gw;
    int *redi = reinterpret_cast<int *>(smem + L.redi);
    unsigned *redb = reinterpret_cast<unsigned *>(smem + L.redb);
    int *rel = reinterpret_cast<int *>(smem + L.rel), *ncnt = rel + (Tc + 1);
    int *cum = reinterpret_cast<int *>(smem + L.seg);      // first 32-pair item of every time point
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + L.mbar);
    double *Wsm = smem + L.w;
    const double prior_const = ct.prior_const[chain];
    const bool is_state = tid < nstate;
    // EVERY warp takes part in the likelihood pass: the dedicated likelihood warps (index 0 .. NWL-1) start right
    // after B1, the state warps (NWL .. nwarps-1) join once the prior terms of the round are done and therefore get
    // a smaller share of the items (DG_STATE_LIK_WEIGHT sixteenths of a likelihood warp's; 0 = none)
    const int wl = is_state ? NWL + warp : warp - S;      // likelihood index of this warp within the block
    const int gwl = crank * nwarps + wl, GW = cs * nwarps;   // ... within the whole cluster
    int xbuf = 0;                                         // exchange buffer (clusters double-buffer)

    const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t0;
    const int r_first = roff[0];
    const int r_total = roff[Tc] - r_first;
    const double *Wg = gt.W + gt.woff[g] + (size_t)C * r_first;
    const size_t wbytes = (size_t)C * r_total * sizeof(double);
    constexpr bool W_IN_SMEM = (WMODE == W_RESIDENT);
    const double *Wc = W_IN_SMEM ? Wsm : Wg;
    double *stage = Wsm + (size_t)wl * dg_stages(KC > 0 ? KC : DG_MAXC) * (dg_stage_slot_bytes(KC > 0 ? KC : DG_MAXC) / 8);   // W_STAGED: per-warp cp.async ring

    // ---- stage W (TMA bulk copy, asynchronous) and the small per-chain constants --------
    if (W_IN_SMEM) {
        if (use_tma) {
            if (tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dg_smem_u32(mbar)));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0 && wbytes > 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                                 dg_smem_u32(mbar)), "r"((uint32_t)wbytes) : "memory");
                const uint32_t kChunk = 32768u;
                size_t done = 0;
                while (done < wbytes) {
                    const uint32_t sz = (uint32_t)((wbytes - done) < kChunk ? (wbytes - done) : kChunk);
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(dg_smem_u32(reinterpret_cast<char *>(Wsm) + done)),
                          "l"(reinterpret_cast<const char *>(Wg) + done), "r"(sz), "r"(dg_smem_u32(mbar))
                        : "memory");
                    done += sz;
                }
            }
        } else {
            const double2 *src = reinterpret_cast<const double2 *>(Wg);
            double2 *dst = reinterpret_cast<double2 *>(Wsm);
            const size_t n2 = wbytes / sizeof(double2);
            for (size_t i = tid; i < n2; i += nthreads) dst[i] = src[i];
        }
    }
    for (int i = tid; i <= Tc; i += nthreads) rel[i] = roff[i] - r_first;
    for (int i = tid; i < Tc; i += nthreads) ncnt[i] = gt.ncnt[(size_t)g * T + t0 + i];
    {
        const double *pool = ct.pool;
        const double *kv = pool + ct.kinv_off[chain];
        for (int i = tid; i < Tc * Tc; i += nthreads) kinv[i] = kv[i];
        if (ct.pf_off[chain] >= 0) {
            const double *pf = pool + ct.pf_off[chain];
            for (int i = tid; i < Tc * Tc; i += nthreads) pfac[i] = pf[i];
        }
        const double *js = pool + ct.js_off[chain], *jc = pool + ct.jc_off[chain];
        for (int i = tid; i < C - 1; i += nthreads) {
            ijs[i] = 1.0 / js[i];
            sqs[i] = sqrt(js[i]);
            jcon[i] = jc[i];
        }
    }
    const int M = ct.M[chain], initial = ct.initial[chain], gap = ct.gap[chain];
    int m = ct.m_next[chain];
    double *st = ct.st + ct.st_off[chain];
    const int rowlen = 2 * CT + 2;
    double *gst = st + rowlen;                            // running Geweke sums (geweke_converged)
    if (m == 0 && lead && tid < 4) gst[tid] = 0.0;
    int64_t *ptab = ct.ptab + ct.ptab_off[chain];
    const int pshift = ct.pshift[chain];
    const int pmask = (1 << pshift) - 1;
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
    const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;
    const bool dev_stream = sa.mode != 0;

    // ---- state lanes: one thread per latent value (c, t); the J candidate nodes of a round
    //      are walked in registers (independent exp / divide chains overlap) -------------------
    const bool s_act = is_state && tid < CT;              // owns a latent value
    const int si = s_act ? tid : 0;                       // c*Tc + t
    const int sc = si / Tc, stt = si - sc * Tc;
    const bool s_free = s_act && sc < C - 1;              // value is sampled (c < C-1)
    double y_now = 0.0, psi_now = 0.0;
    double P_now = 0.0, L_now = 0.0;                      // thread 0
    int cnt = 0, flags = 0;
    if (tid == 0) { cnt = ct.cnt[chain]; flags = ct.flags[chain]; }
    const double len_own = s_act ? gt.len[gt.lenoff[g] + (size_t)(t0 + stt) * C + sc] : 0.0;
    const double *kinv_col = kinv + stt;                  // column t of K^-1
    double *ff_own = ff + (stt * C + sc) * JP;            // f of this (t, c), all candidates
    const double *y0p = ct.y0_off[chain] >= 0 ? ct.pool + ct.y0_off[chain] : nullptr;

    __syncthreads();                                      // rel / ncnt visible
    if (tid == 0) {
        int a = 0;
        for (int t = 0; t < Tc; ++t) { cum[t] = a; a += (((rel[t + 1] - rel[t]) >> 1) + IPAIRS - 1) / IPAIRS; }
        cum[Tc] = a;
    }
    if (m > 0 && s_act) { y_now = st[si]; psi_now = st[CT + si]; }
    if (m > 0 && tid == 0) { P_now = st[2 * CT]; L_now = st[2 * CT + 1]; }
    if (W_IN_SMEM && use_tma && wbytes > 0) {
        uint32_t ok = 0;                                  // wait for the bulk copy (phase 0)
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(dg_smem_u32(mbar)), "r"(0u) : "memory");
        }
    }
    __syncthreads();
    // this warp's run of likelihood items: the items of the chain split over all warps of the block (cluster) in
    // proportion to their weights
    int lq0, lq1, ring_pos = -1;
    {
        const long long wblock = 16ll * NWL + (long long)DG_STATE_LIK_WEIGHT * S, wall_ = wblock * cs;
        const long long pre = (long long)crank * wblock + (wl < NWL ? 16ll * wl : 16ll * NWL + (long long)DG_STATE_LIK_WEIGHT * (wl - NWL));
        const long long own = wl < NWL ? 16 : DG_STATE_LIK_WEIGHT;
        const long long nitems = cum[Tc];
        lq0 = (int)((nitems * pre) / wall_);
        lq1 = (int)((nitems * (pre + own)) / wall_);
    }

    const int m_start = m;
    const int m_end = (m_start + n_steps < M) ? (m_start + n_steps) : M;
    int state = 0, m_ret = 0;
    long long page_cur = -1;                              // thread 0: page of the last row written
    const double *dblk = dev_stream ? nullptr : sa.delta + sa.delta_off[slot];
    const double *ublk = dev_stream ? nullptr : sa.u + sa.u_off[slot];
    // stream ring: two batches of SB steps; a new batch is produced (by the likelihood warps,
    // with all their lanes busy) whenever the chain is within one batch of the end of the ring
    int produced = m;                                     // ring holds steps [.., produced)
    for (int b = 0; b < RING / SB && produced < m_end; ++b) {
        const int upto = min(produced + SB, m_end);
        produce_stream(dev_stream, produced, upto, tid, nthreads, 0, zring, dring, RING, C, Tc, pfac, sqs, kinv,
                       jcon, ijs, sa.seed, sid, dblk, ublk, m_start);
        produced = upto;
        __syncthreads();
    }
    __syncthreads();
#ifdef DG_PROFILE
    long long dg_prof[10];
    long long dg_t0 = clock64();
    for (int i = 0; i < 10; ++i) dg_prof[i] = 0;
#endif

    // Shape of the J candidate nodes of a round (uniform over the block, may change between
    // rounds).  0 = binary tree in heap order (node 2j+1 follows a rejection of node j, 2j+2 an
    // acceptance; J = 3: always 2 steps per round, J = 6: three levels without the accept-accept
    // leaf).  1 = rejection chain: node d proposes step m+d from the unchanged state, the round
    // ends at the first acceptance: (1 - (1-p)^J) / p steps per round, the better shape when the
    // acceptance rate p is low (time-series genes run at ~0.16).
#ifdef DG_NO_CHAIN
    constexpr bool kChain = false;
#else
    constexpr bool kChain = (J >= 3);
#endif
    constexpr int kThr20 = (J == 3) ? 7 : 6;              // chain when p < kThr20 / 20
    constexpr int DM = kChain ? J : DT;                   // most steps a round can commit
    int shape = (kChain && m > 0 && ct.cnt[chain] * 20 < m * kThr20) ? 1 : 0;
    bool init = (m == 0);                                 // evaluate the start value first (model_GP.py:275-292)
    // next convergence-check step at or after m (m >= initial and m % gap == 0)
    int next_check = m > initial ? m : initial;
    next_check += (gap - next_check % gap) % gap;
    int rbase = m % RING;                                 // ring slot of step m
    unsigned long long nrounds = 0;                       // passes over W (thread 0 reports them)
    while (m < m_end || init) {
        ++nrounds;
        // steps this round may commit: up to D, not past the chunk, not past a check step
        const int nd_max = init ? 0 : min(shape ? DM : DT, min(m_end - m, next_check - m + 1));
        // ================= S phase (state warps): candidates -> f =================
        double yj[J], bj[J], ej[J], gj[J];
        if (is_state) {
            if (s_act) {
                if (init) {
                    const double v = y0p ? y0p[si] : 0.0;                    // Y_now = Ymean (:275)
#pragma unroll
                    for (int j = 0; j < J; ++j) { yj[j] = v; bj[j] = v; }
                } else if (s_free) {
                    double dl[DM];                                           // increments of steps m .. m+DM-1
#pragma unroll
                    for (int d = 0; d < DM; ++d) {
                        int sl = rbase + d;
                        if (sl >= RING) sl -= RING;
                        dl[d] = d < nd_max ? dring[sl * RL + si] : 0.0;
                    }
                    // state before node j proposes = Y_now moved through its accepted ancestors;
                    // Y_try = delta + that state (:329, numpy adds the mean last)
                    if (kChain && shape) {
#pragma unroll
                        for (int j = 0; j < J; ++j) { bj[j] = y_now; yj[j] = dl[j < DM ? j : 0] + y_now; }
                    } else {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const int dep = (j >= 3) ? 2 : (j >= 1 ? 1 : 0);
                            if (j == 0) bj[j] = y_now;
                            else bj[j] = (j & 1) ? bj[(j - 1) / 2] : yj[(j - 1) / 2];   // odd = parent rejected
                            yj[j] = dl[dep] + bj[j];
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) { yj[j] = 0.0; bj[j] = 0.0; }    // Y_try[C-1,:] stays 0 (:295)
                }
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    ej[j] = exp(yj[j]);
                    gj[j] = len_own * ej[j];
                    ycand[j * CT + si] = yj[j];
                    ex[j * CT + si] = ej[j];
                    gs[j * CT + si] = gj[j];
                }
            }
            bar_state(nstate);
            if (s_act) {
                // fsi = len*psi / sum(len*psi) with psi = e/sum(e): the common 1/sum(e) cancels
                // (model_GP.py:335-337; one division on the critical path instead of two)
#pragma unroll
                for (int j = 0; j < J; ++j) ff_own[j] = gj[j] / dg_np_sum(gs + j * CT + stt, C, Tc);
            }
        }
        __syncthreads();                                                     // B1: f ready
        DG_TICK(0);
        // ================= rest of the serial part (state warps), then the likelihood pass (every warp) ==========
        if (is_state) {
            if (s_act) {
#pragma unroll
                for (int j = 0; j < J; ++j) psic[j * CT + si] = ej[j] / dg_np_sum(ex + j * CT + stt, C, Tc);   // (:335)
            }
            // GP prior (:332): x K^-1 x as sum_t (x K^-1)[t] * x[t]
            if (s_free) {
                double r[J];
#pragma unroll
                for (int j = 0; j < J; ++j) r[j] = 0.0;
                for (int t2 = 0; t2 < Tc; ++t2) {
                    const double kv = kinv_col[t2 * Tc];
#pragma unroll
                    for (int j = 0; j < J; ++j) r[j] = fma(ycand[j * CT + sc * Tc + t2], kv, r[j]);
                }
#pragma unroll
                for (int j = 0; j < J; ++j) pr[j * CT + si] = r[j] * yj[j];
            }
            bar_state(nstate);
            for (int q = tid; q < J * (C - 1); q += nstate) {                // (node, c): sums over t
                const int jn = q / (C - 1), c = q - jn * (C - 1);
                const double *prc = pr + jn * CT + c * Tc;
                double a = 0.0;
                for (int t = 0; t < Tc; ++t) a += prc[t];
                gs[jn * CT + c] = prior_const - a / 2.0;                     // gs is free after B1
            }
            bar_state(nstate);
            for (int q = tid; q < J; q += nstate) {
                double prior = 0.0;                                          // c ascending (:312-332)
                for (int c = 0; c < C - 1; ++c) prior += gs[q * CT + c];
                bc[16 + q] = prior;
            }
        }
        DG_TICK(1);
        {
            double mm[J], lsd[J];
            int ee[J];
            unsigned bad = 0u;
#pragma unroll
            for (int j = 0; j < J; ++j) { mm[j] = 1.0; ee[j] = 0; lsd[j] = 0.0; }
            lik_pass<KC, J, WMODE, false>(Wc, C, Tc, rel, ncnt, cum, ff, stage, lane, lq0, lq1, ring_pos, mm, ee, bad, lsd);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int gg = __double2hiint(mm[j]);
                ee[j] += (gg >> 20) - 1023;
                mm[j] = dg_mant(mm[j], gg);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    mm[j] *= __shfl_xor_sync(0xffffffffu, mm[j], off);       // < 2^32
                    ee[j] += __shfl_xor_sync(0xffffffffu, ee[j], off);
                }
                bad |= __shfl_xor_sync(0xffffffffu, bad, off);
            }
            if (lane == 0) {
                double *rd = red + xbuf * J * SLOTS;
                int *ri = redi + xbuf * J * SLOTS;
                int *rb = reinterpret_cast<int *>(redb) + xbuf * SLOTS;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int gg = __double2hiint(mm[j]);
                    const double mv = dg_mant(mm[j], gg);
                    const int ev = ee[j] + (gg >> 20) - 1023;
                    if (cs == 1) { rd[j * SLOTS + wl] = mv; ri[j * SLOTS + wl] = ev; }
                    else for (int r = 0; r < cs; ++r) { st_cluster(rd + j * SLOTS + gwl, r, mv); st_cluster(ri + j * SLOTS + gwl, r, ev); }
                }
                if (cs == 1) rb[wl] = (int)bad;
                else for (int r = 0; r < cs; ++r) st_cluster(rb + gwl, r, (int)bad);
            }
        }
        DG_TICK(6);
        if (cs == 1) __syncthreads(); else cluster_barrier();                // B2: partial likelihoods ready
        DG_TICK(2);
        // ================= decisions (warp 0; thread 0 walks the tree) =================
        bool redo = false;
        do {
            if (warp == 0) {
                double lik[J];
                unsigned bad = 0u;
                if (!redo) {
                    double mm[J];
                    int ee[J];
                    const double *rd = red + xbuf * J * SLOTS;
                    const int *ri = redi + xbuf * J * SLOTS;
                    const unsigned *rb = redb + xbuf * SLOTS;
#pragma unroll
                    for (int j = 0; j < J; ++j) { mm[j] = 1.0; ee[j] = 0; }
                    for (int w = lane; w < GW; w += 32) {                    // <= 4 slots per lane: product < 16
#pragma unroll
                        for (int j = 0; j < J; ++j) { mm[j] *= rd[j * SLOTS + w]; ee[j] += ri[j * SLOTS + w]; }
                        bad |= rb[w];
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            mm[j] *= __shfl_xor_sync(0xffffffffu, mm[j], off);   // < 2^32
                            ee[j] += __shfl_xor_sync(0xffffffffu, ee[j], off);
                        }
                        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
                    }
                    // lane j takes the logarithm of candidate j (one log latency for all of them)
                    double mine = mm[0];
                    int eme = ee[0];
#pragma unroll
                    for (int j = 1; j < J; ++j) if (lane == j) { mine = mm[j]; eme = ee[j]; }
                    const double lk = fma((double)eme, 0.693147180559945309417232121458, log(mine));
#pragma unroll
                    for (int j = 0; j < J; ++j) lik[j] = __shfl_sync(0xffffffffu, lk, j);
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        double a = 0.0;
                        for (int w = 0; w < GW; ++w) a += red[xbuf * J * SLOTS + j * SLOTS + w];
                        lik[j] = a;
                    }
                }
                if (lane == 0) {
                    bc[3] = bad ? 1.0 : 0.0;
                    if (!bad) {
                        if (init) {
                            L_now = lik[0];
                            P_now = bc[16] + lik[0];                         // (:284-292)
                            if (!isfinite(P_now)) flags |= 1;
                            bc[4] = 0.0;                                     // steps committed
                            bc[5] = 0.0;                                     // final node
                            bc[6] = 0.0;
                        } else {
                            // trace pages of the rows this round may write (first touch allocates)
                            int nd_ok = nd_max;
                            for (int d = 0; d < nd_max && lead; ++d) {
                                const int step = m + d;
                                if (page_cur < 0 || (step & pmask) == 0)
                                    page_cur = trace_page(ptab, step, pshift, rowlen, ct.pool_head, ct.pool_cap);
                                if (page_cur < 0) { nd_ok = d; break; }
                                bc[32 + d] = __longlong_as_double(page_cur);
                            }
                            // walk the tree: alpha = exp(min(a, 0)), accept iff u < alpha (:348-351),
                            // tested in the log domain with log(u) taken off the critical path
                            int node = 0, cur = -1, nd_done = nd_ok;
                            for (int d = 0; d < nd_ok; ++d) {
                                double likn = lik[0], priorn = bc[16 + node];
#pragma unroll
                                for (int j = 1; j < J; ++j) if (j == node) likn = lik[j];
                                const double P_try = priorn + likn;
                                int sl = rbase + d;
                                if (sl >= RING) sl -= RING;
                                const double lu = dring[sl * RL + CT], Q = dring[sl * RL + CT + 1];
                                const double a = ((P_try + Q) - P_now) - Q;
                                const bool acc = (a >= 0.0) || (lu < a);     // NaN a: both false
                                if (!isfinite(P_try)) flags |= 1;
                                if (acc) { ++cnt; P_now = P_try; L_now = likn; cur = node; }
                                bc[40 + d] = (double)cur;                    // state after step m+d
                                bc[48 + d] = P_now;                          // Pro_all[m+d]  (:362)
                                bc[56 + d] = L_now;                          // Lik_all[m+d]  (:363)
                                if (shape) {                                 // chain: later candidates assumed this rejection
                                    if (acc) { nd_done = d + 1; break; }
                                    node = d + 1;
                                } else {
                                    node = 2 * node + (acc ? 2 : 1);
                                    if (node >= J) { nd_done = d + 1; break; }   // leaf not evaluated this round
                                }
                            }
                            // a chain round that stopped early leaves the page cursor on its last
                            // committed row (the pages of the unused rows stay allocated for later)
                            if (nd_done > 0 && nd_done < nd_ok && lead) page_cur = __double_as_longlong(bc[32 + nd_done - 1]);
                            bc[4] = (double)nd_done;
                            bc[5] = (double)cur;
                            bc[6] = (nd_ok < nd_max) ? 1.0 : 0.0;            // trace pool exhausted
                            // shape of the next round from the running acceptance rate (p < 0.35 -> chain)
                            const int m_new = m + nd_done;
                            bc[7] = (kChain && m_new >= 64 && cnt * 20 < m_new * kThr20) ? 1.0 : 0.0;
                        }
                    }
                }
            }
            __syncthreads();                                                 // B3: decisions (or redo request)
            redo = bc[3] != 0.0;
            if (cs > 1) xbuf ^= 1;
            if (redo) {                                                      // rare: some x_n was not a positive normal
                lik_warp_pass_log<KC, J, WMODE>(Wc, rel, ncnt, cum, ff, stage, red + xbuf * J * SLOTS, C, Tc, lq0, lq1,
                                                ring_pos, lane, gwl, SLOTS, cs);
                if (cs == 1) __syncthreads(); else cluster_barrier();
            }
        } while (redo);
        DG_TICK(3);
        // ================= update + trace rows (state warps) || stream refill (others) ======
        const int nd = (int)bc[4];
        const int fin = (int)bc[5];
        const int shape_next = init ? shape : (int)bc[7];
        const bool nomem = bc[6] != 0.0;
        if (is_state) {
            if (s_act) {
                for (int d = 0; d < nd && lead; ++d) {
                    const int node = (int)bc[40 + d];
                    const int step = m + d;
                    double *row = ct.trace + __double_as_longlong(bc[32 + d]) + (size_t)(step & pmask) * rowlen;
                    row[si] = node < 0 ? psi_now : psic[node * CT + si];     // Psi_all[m]  (:365)
                    row[CT + si] = node < 0 ? y_now : ycand[node * CT + si]; // Y_all[m]    (:364)
                    if (tid == 0) { row[2 * CT] = bc[48 + d]; row[2 * CT + 1] = bc[56 + d]; }
                }
                if (init || fin >= 0) {
                    const int f0 = init ? 0 : fin;
                    y_now = ycand[f0 * CT + si];
                    psi_now = psic[f0 * CT + si];
                }
            }
            bar_state(nstate);                                               // candidates may be overwritten now
        } else if (!init && produced < m_end && produced - (m + nd) <= RING - SB) {
            // the batch behind the chain is consumed: overwrite it with the next SB steps
            produce_stream(dev_stream, produced, min(produced + SB, m_end), tid - nstate, nthreads - nstate,
                           3, zring, dring, RING, C, Tc, pfac, sqs, kinv, jcon, ijs, sa.seed, sid, dblk, ublk, m_start);
        }
        if (!init && produced < m_end && produced - (m + nd) <= RING - SB) produced = min(produced + SB, m_end);
        if (init) { init = false; continue; }
        shape = shape_next;
        m += nd;
        rbase += nd;
        if (rbase >= RING) rbase -= RING;
        if (nomem) {                                                         // trace pool exhausted (uniform)
            state = DICEGP_CHAIN_NOMEM_;
            m_ret = m - 1;
            break;
        }
        // convergence diagnostics (:369-391) after step m-1 if it is a check step; the rows
        // [0, m-1) exclude the one just written
        const int mlast = m - 1;
        if (nd > 0 && mlast >= initial && (mlast % gap) == 0) {
            if (cs == 1) __syncthreads(); else cluster_barrier();        // the lead block's rows are visible
            DG_TICK(4);
            int conv;
            if (cs == 1) {
                conv = geweke_converged(gw, gst, CTm, nwarps, warp, lane, tv, mlast);
            } else {                                                     // the lead block decides for the cluster
                if (lead) {
                    conv = geweke_converged(gw, gst, CTm, nwarps, warp, lane, tv, mlast);
                    if (tid == 0) for (int r = 0; r < cs; ++r) st_cluster(bc + 8, r, (double)conv);
                }
                cluster_barrier();
                conv = (int)bc[8];
            }
            DG_TICK(5);
            if (conv) {
                state = DICEGP_CHAIN_CONVERGED_;
                m_ret = mlast;
                break;
            }
            next_check += gap;
        }
        DG_TICK(4);
    }
#ifdef DG_PROFILE
    if (threadIdx.x == 0 && sa.prof) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(const_cast<double *>(sa.prof));
        for (int i = 0; i < 10; ++i) atomicAdd(o + i, (unsigned long long)dg_prof[i]);
        atomicAdd(o + 10, (unsigned long long)(m - m_start));
    }
#endif
    if (state == 0 && m >= M) { state = DICEGP_CHAIN_EXHAUSTED_; m_ret = M - 1; }

    // ---- save state ------------------------------------------------------------------------
    if (WMODE == W_STAGED) cp_async_wait<0>();            // the copies of the pass that will not run
    __syncthreads();
    if (s_act && lead) { st[si] = y_now; st[CT + si] = psi_now; }
    if (tid == 0) {
        if (lead) {
            st[2 * CT] = P_now; st[2 * CT + 1] = L_now;
            ct.m_next[chain] = m;
            ct.cnt[chain] = cnt;
            ct.flags[chain] = flags;
            ct.m_ret[chain] = m_ret;
            ct.state[chain] = state;
            if (ct.ctr) {                                 // work counters: W bytes brought on chip, rounds, candidate evaluations, steps
                atomicAdd(ct.ctr + 0, W_IN_SMEM ? (unsigned long long)wbytes : nrounds * (unsigned long long)wbytes);
                atomicAdd(ct.ctr + 1, nrounds);
                atomicAdd(ct.ctr + 2, nrounds * (unsigned long long)J * (unsigned long long)r_total);
                atomicAdd(ct.ctr + 3, (unsigned long long)(m - m_start));
                atomicAdd(ct.ctr + 4, nrounds * (unsigned long long)J * (unsigned long long)r_total * (unsigned long long)C);
            }
        }
        if (W_IN_SMEM && use_tma)
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(dg_smem_u32(mbar)) : "memory");
    }
    __syncthreads();
}

}  // namespace dg
