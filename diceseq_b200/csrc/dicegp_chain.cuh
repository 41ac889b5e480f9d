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

__device__ __forceinline__ void bar_state(int n) { asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory"); }
__device__ __forceinline__ void bar_lik(int n) { asm volatile("bar.sync 3, %0;" ::"r"(n) : "memory"); }

// mantissa in [1,2) of a positive normal double whose high word is `hi`
__device__ __forceinline__ double dg_mant(double x, int hi) {
    return __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
}

// One likelihood segment: pairs p = p0, p0+stride, ... of one time point, J candidates.
//   x_n^(j) = sum_k W[n,k] f^(j)[k]                            model_GP.py:338
// W block: isoform-major, row k = npairs double2 (two reads per 16-byte load; consecutive
// lanes take consecutive pairs: conflict-free LDS.128 / coalesced LDG.128).  f of the J
// candidates sits in shared memory as ff[k][JP] and is read as broadcast LDS.128.
// prod_n x_n is carried per candidate as a mantissa product and an exponent sum; anything
// that is not a positive normal double raises `bad` (the round is then redone with logs).
template <int KC, int J, bool LOGPATH>
__device__ __forceinline__ void lik_segment(const double2 *__restrict__ Wt, int npairs, int n, int C,
                                            const double *ffcol, int p0, int stride, double (&mm)[J],
                                            int (&ee)[J], unsigned &bad, double (&ls)[J]) {
    constexpr int JP = J + 1;
    int it = 0;
    for (int p = p0; p < npairs; p += stride) {
        double x0[J], x1[J];
#pragma unroll
        for (int j = 0; j < J; ++j) { x0[j] = 0.0; x1[j] = 0.0; }
        const int kmax = KC > 0 ? KC : C;
#pragma unroll
        for (int k = 0; k < kmax; ++k) {
            const double2 w = Wt[(size_t)k * npairs + p];
            const double2 *f2 = reinterpret_cast<const double2 *>(ffcol + k * JP);
            double fj[JP];
#pragma unroll
            for (int q = 0; q < JP / 2; ++q) { const double2 t = f2[q]; fj[2 * q] = t.x; fj[2 * q + 1] = t.y; }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                x0[j] = fma(w.x, fj[j], x0[j]);
                x1[j] = fma(w.y, fj[j], x1[j]);
            }
        }
        const bool v1 = (2 * p + 1) < n;              // the pad slot of an odd block
        if (LOGPATH) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                ls[j] += log(x0[j]);
                if (v1) ls[j] += log(x1[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double y1 = v1 ? x1[j] : 1.0;
                const int h0 = __double2hiint(x0[j]), h1 = __double2hiint(y1);
                bad |= (((unsigned)(h0 - 0x00100000) >= 0x7fe00000u) || ((unsigned)(h1 - 0x00100000) >= 0x7fe00000u)) ? 1u : 0u;
                ee[j] += (h0 >> 20) + (h1 >> 20) - 2046;
                mm[j] *= dg_mant(x0[j], h0) * dg_mant(y1, h1);
            }
            if (((++it) & 255) == 0) {                 // keep the running products in range
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int g = __double2hiint(mm[j]);
                    ee[j] += (g >> 20) - 1023;
                    mm[j] = dg_mant(mm[j], g);
                }
            }
        }
    }
}

// Slow path (some x_n is 0 / negative / non-finite): the same sums with real logarithms so
// that -inf and NaN propagate like np.log(...).sum() does (model_GP.py:344).
template <int KC, int J>
__device__ __noinline__ void lik_warp_pass_log(const double *Wc, const int *seg, const int *rel, const int *ncnt,
                                               const double *ff, double *red, int C, int NWL, int lane, int wl,
                                               int nseg) {
    constexpr int JP = J + 1;
    double mm[J], ls[J];
    int ee[J];
    unsigned bad = 0u;
#pragma unroll
    for (int j = 0; j < J; ++j) { mm[j] = 1.0; ls[j] = 0.0; ee[j] = 0; }
    for (int sg = wl; sg < nseg; sg += NWL) {
        const int tt = seg[3 * sg], first = seg[3 * sg + 1], stride = seg[3 * sg + 2];
        const int r0 = rel[tt], npairs = (rel[tt + 1] - r0) >> 1;
        const double2 *Wt = reinterpret_cast<const double2 *>(Wc + (size_t)C * r0);
        lik_segment<KC, J, true>(Wt, npairs, ncnt[tt], C, ff + (size_t)tt * C * JP, first + lane, stride, mm, ee, bad, ls);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        double v = ls[j];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[j * DG_MAXWARPS + wl] = v;
    }
}

// ---------------------------------------------------------------------------------------
// Geweke convergence rule over trace rows [0, m)          model_GP.py:156-185, :369-391
// Series = Psi_all[:m, c, t] for c < C-1 (columns 0..CTm-1 of a trace row).
// Two-pass mean / population variance like np.mean / np.var.  Lanes walk columns
// (coalesced), warps walk rows.  Returns 1 to every thread when all |Z| <= 2.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ int geweke_converged(double *gw, int CTm, int nwarps, int warp, int lane, TraceView tv, int m) {
    const int nA = (int)(0.1 * (double)m);
    const int iB = (int)(0.5 * (double)m);
    const int nB = m - iB;
    int ok = 1;
    for (int c0 = 0; c0 < CTm; c0 += DG_GW_COLS) {
        const int col = c0 + lane;
        const bool act = col < CTm;
        double *gA = gw, *gB = gw + nwarps * DG_GW_COLS;
        double sa = 0.0, sb = 0.0;
        if (act) {
            for (int r = warp; r < nA; r += nwarps) sa += __ldcg(tv.row(r) + col);
            for (int r = iB + warp; r < m; r += nwarps) sb += __ldcg(tv.row(r) + col);
        }
        gA[warp * DG_GW_COLS + lane] = sa;
        gB[warp * DG_GW_COLS + lane] = sb;
        __syncthreads();
        double meanA = 0.0, meanB = 0.0;
        for (int w = 0; w < nwarps; ++w) {
            meanA += gA[w * DG_GW_COLS + lane];
            meanB += gB[w * DG_GW_COLS + lane];
        }
        meanA /= (double)nA;
        meanB /= (double)nB;
        __syncthreads();
        double qa = 0.0, qb = 0.0;
        if (act) {
            for (int r = warp; r < nA; r += nwarps) {
                const double d = __ldcg(tv.row(r) + col) - meanA;
                qa = fma(d, d, qa);
            }
            for (int r = iB + warp; r < m; r += nwarps) {
                const double d = __ldcg(tv.row(r) + col) - meanB;
                qb = fma(d, d, qb);
            }
        }
        gA[warp * DG_GW_COLS + lane] = qa;
        gB[warp * DG_GW_COLS + lane] = qb;
        __syncthreads();
        if (warp == 0 && act) {
            double va = 0.0, vb = 0.0;
            for (int w = 0; w < nwarps; ++w) {
                va += gA[w * DG_GW_COLS + lane];
                vb += gB[w * DG_GW_COLS + lane];
            }
            va /= (double)nA;
            vb /= (double)nB;
            const double den = sqrt(va + vb);
            // reference: Z None (den == 0) or Z > 2 -> not converged; NaN compares false
            if (den == 0.0) ok = 0;
            else if (fabs(meanA - meanB) / den > 2.0) ok = 0;
        }
        __syncthreads();
    }
    return __syncthreads_and(ok);
}

// Which likelihood warp reads which slice: segment = (time point, first pair, stride).
// Warps are dealt to time points in proportion to their read pairs; with fewer warps
// than time points every time point is one segment and warps take several.
__device__ __noinline__ int build_segments(int *seg, const int *rel, int Tc, int NWL) {
    if (Tc >= NWL) {
        for (int t = 0; t < Tc; ++t) { seg[3 * t] = t; seg[3 * t + 1] = 0; seg[3 * t + 2] = 32; }
        return Tc;
    }
    int nw[DG_MAXT];
    for (int t = 0; t < Tc; ++t) nw[t] = 1;
    for (int w = Tc; w < NWL; ++w) {
        int best = 0;
        double load = -1.0;
        for (int t = 0; t < Tc; ++t) {
            const double l = (double)(rel[t + 1] - rel[t]) / (double)nw[t];
            if (l > load) { load = l; best = t; }
        }
        ++nw[best];
    }
    int k = 0;
    for (int t = 0; t < Tc; ++t)
        for (int j = 0; j < nw[t]; ++j) { seg[3 * k] = t; seg[3 * k + 1] = 32 * j; seg[3 * k + 2] = 32 * nw[t]; ++k; }
    return k;
}

// Out-of-line copies of the long math routines used on the round loop: the loop body has
// to stay small enough for the instruction cache (the serial part runs in one warp, which
// cannot hide instruction-fetch latency).
__device__ __noinline__ double dg_log(double x) { return log(x); }

// Fill stream ring entries for steps [s0, s1): increments delta[c,t] and log(u).
// Device stream: Philox normals z (Box-Muller, both outputs of a pair used), then
// delta = sqrt(jump_scale[c]) * (z[c,:] @ A_K).  Host stream: delta and u come from the
// uploaded chunk.  Called by `nprod` threads (index `pt`) which must ALL call it: they
// synchronise on __syncthreads (all_threads) or on the likelihood-warp barrier.
__device__ __noinline__ void produce_stream(bool dev_stream, int s0, int s1, int pt, int nprod, bool all_threads,
                                            double *zring, double *dring, int RING, int CT, int CTm, int Tc,
                                            const double *pfac, const double *sqs, uint64_t seed, int64_t sid,
                                            const double *dblk, const double *ublk, int m_start) {
    if (dev_stream) {
        const int npair = (CTm + 1) >> 1;
        for (int s = s0; s < s1; ++s) {
            double *z = zring + (s % RING) * CT;
            for (int i = pt; i < npair; i += nprod) {
                double z0, z1;
                dg_normal_pair(seed, sid, (uint32_t)s, (uint32_t)i, z0, z1);
                z[2 * i] = z0;
                if (2 * i + 1 < CTm) z[2 * i + 1] = z1;
            }
            if (pt == nprod - 1) dring[(s % RING) * CT + CT - 1] = dg_log(dg_uniform(seed, sid, (uint32_t)s));
        }
        if (all_threads) __syncthreads(); else bar_lik(nprod);
        for (int s = s0; s < s1; ++s) {
            const double *z = zring + (s % RING) * CT;
            double *d = dring + (s % RING) * CT;
            for (int i = pt; i < CTm; i += nprod) {
                const int c = i / Tc, t = i - c * Tc;
                double a = 0.0;
                for (int j = 0; j < Tc; ++j) a = fma(z[c * Tc + j], pfac[j * Tc + t], a);
                d[i] = sqs[c] * a;
            }
        }
    } else {
        for (int s = s0; s < s1; ++s) {
            double *d = dring + (s % RING) * CT;
            const double *src = dblk + (size_t)(s - m_start) * CTm;
            for (int i = pt; i < CTm; i += nprod) d[i] = src[i];
            if (pt == nprod - 1) d[CT - 1] = dg_log(ublk[s - m_start]);
        }
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
template <int KC, bool W_IN_SMEM, int D>
__device__ void run_chain(const GeneTab &gt, const ChainTab &ct, const StreamArgs &sa, int chain,
                          int slot, int n_steps, double *smem, int use_tma) {
    constexpr int J = (1 << D) - 1;
    constexpr int JP = J + 1;
    constexpr int RING = 2 * D;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    if (ct.state[chain] != 0) return;                    // already finished (block-uniform)

    const int g = ct.gene[chain], t0 = ct.t0[chain];
    const int C = KC > 0 ? KC : gt.C[g];
    const int Tc = ct.Tc[chain], T = gt.T;
    const int CT = C * Tc, CTm = (C - 1) * Tc;
    const int S = (J * CT + 31) >> 5, nstate = S << 5, NWL = nwarps - S;
    const SmLayout L = dg_layout(C, Tc, nwarps, D);
    double *ycand = smem + L.ycand, *psic = smem + L.psic, *gs = smem + L.gs, *ex = smem + L.ex;
    double *pr = smem + L.pr, *qq = smem + L.qq, *ff = smem + L.ff, *zring = smem + L.zring, *dring = smem + L.dring;
    double *kinv = smem + L.kinv, *pfac = smem + L.pfac, *ijs = smem + L.ijs, *jcon = smem + L.jcon, *sqs = smem + L.sqs;
    double *red = smem + L.red, *bc = smem + L.bc, *gw = smem + L.gw;
    int *redi = reinterpret_cast<int *>(smem + L.redi);
    unsigned *redb = reinterpret_cast<unsigned *>(smem + L.redb);
    int *rel = reinterpret_cast<int *>(smem + L.rel), *ncnt = rel + (Tc + 1);
    int *seg = reinterpret_cast<int *>(smem + L.seg);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + L.mbar);
    double *Wsm = smem + L.w;
    const double prior_const = ct.prior_const[chain];
    const bool is_state = tid < nstate;
    const int wl = warp - S;                              // likelihood-warp index (>= 0 for those)

    const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t0;
    const int r_first = roff[0];
    const int r_total = roff[Tc] - r_first;
    const double *Wg = gt.W + gt.woff[g] + (size_t)C * r_first;
    const size_t wbytes = (size_t)C * r_total * sizeof(double);
    const double *Wc = W_IN_SMEM ? Wsm : Wg;

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
    int64_t *ptab = ct.ptab + ct.ptab_off[chain];
    const int pshift = ct.pshift[chain];
    const int pmask = (1 << pshift) - 1;
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
    const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;
    const bool dev_stream = sa.mode != 0;

    // ---- state lanes: lane s = node*CT + (c*Tc + t) ----------------------------------------
    const bool s_act = is_state && tid < J * CT;          // owns a latent value of a candidate node
    const int nd_j = s_act ? tid / CT : 0;                // candidate node (heap order)
    const int si = s_act ? tid - nd_j * CT : 0;           // c*Tc + t
    const int sc = si / Tc, stt = si - sc * Tc;
    const bool s_free = s_act && sc < C - 1;              // value is sampled (c < C-1)
    int nd_depth = 0;                                     // depth of the node = step offset it proposes
    while ((2 << nd_depth) <= nd_j + 1) ++nd_depth;
    const int nd_path = (nd_j + 1) - (1 << nd_depth);     // edge bits root->node, MSB first: 1 = accepted parent
    double y_now = 0.0, psi_now = 0.0;                    // replicated over the J node groups
    double P_now = 0.0, L_now = 0.0;                      // thread 0
    int cnt = 0, flags = 0;
    if (tid == 0) { cnt = ct.cnt[chain]; flags = ct.flags[chain]; }
    const double len_own = s_act ? gt.len[gt.lenoff[g] + (size_t)(t0 + stt) * C + sc] : 0.0;

    __syncthreads();                                      // rel / ncnt visible
    if (tid == 0) bc[15] = (double)build_segments(seg, rel, Tc, NWL);
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
    const int nseg = (int)bc[15];

    const int m_start = m;
    const int m_end = (m_start + n_steps < M) ? (m_start + n_steps) : M;
    int state = 0, m_ret = 0;
    long long page_cur = -1;                              // thread 0: page of the last row written
    const double *dblk = dev_stream ? nullptr : sa.delta + sa.delta_off[slot];
    const double *ublk = dev_stream ? nullptr : sa.u + sa.u_off[slot];
    // stream ring: entries for steps [m, m + RING) (clipped to the chunk)
    produce_stream(dev_stream, m, min(m + RING, m_end), tid, nthreads, true, zring, dring, RING, CT, CTm, Tc, pfac, sqs,
                   sa.seed, sid, dblk, ublk, m_start);
    __syncthreads();
#ifdef DG_PROFILE
    long long dg_prof[10];
    long long dg_t0 = clock64();
    for (int i = 0; i < 10; ++i) dg_prof[i] = 0;
#endif

    bool init = (m == 0);                                 // evaluate the start value first (model_GP.py:275-292)
    while (m < m_end || init) {
        // steps this round may commit: up to D, not past the chunk, not past a check step
        int nd_max = 0;
        if (!init) {
            int mc = m > initial ? m : initial;
            mc += (gap - mc % gap) % gap;                 // next step with m >= initial and m % gap == 0
            nd_max = min(D, min(m_end - m, mc - m + 1));
        }
        // ================= S phase (state warps): candidates -> f =================
        double y_c = 0.0, base_c = 0.0, e_c = 1.0, g_c = 0.0, psi_c = 0.0;
        if (is_state) {
            if (s_act) {
                if (init) {
                    const double *y0 = ct.y0_off[chain] >= 0 ? ct.pool + ct.y0_off[chain] : nullptr;
                    y_c = y0 ? y0[si] : 0.0;                                 // Y_now = Ymean (:275)
                    base_c = y_c;
                } else if (s_free) {
                    // state before this node proposes: Y_now moved through its accepted ancestors
                    double cur = y_now;
#pragma unroll
                    for (int i = 0; i < D - 1; ++i) {
                        if (i < nd_depth && i < nd_max && ((nd_path >> (nd_depth - 1 - i)) & 1))
                            cur = dring[((m + i) % RING) * CT + si] + cur;
                    }
                    const double dl = (nd_depth < nd_max) ? dring[((m + nd_depth) % RING) * CT + si] : 0.0;
                    y_c = dl + cur;                                          // (:329) numpy adds the mean last
                    base_c = cur;
                }                                                            // else Y_try[C-1,:] stays 0 (:295)
                e_c = exp(y_c);
                g_c = len_own * e_c;
                ycand[tid] = y_c;
                ex[tid] = e_c;
                gs[tid] = g_c;
            }
            bar_state(nstate);
            if (s_act) {
                // fsi = len*psi / sum(len*psi) with psi = e/sum(e): the common 1/sum(e) cancels
                // (model_GP.py:335-337; one division on the critical path instead of two)
                const double sg = dg_np_sum(gs + nd_j * CT + stt, C, Tc);
                ff[(stt * C + sc) * JP + nd_j] = g_c / sg;
            }
        }
        __syncthreads();                                                     // B1: f ready
        DG_TICK(0);
        // ================= likelihood warps || rest of the serial part =================
        if (!is_state) {
            double mm[J], lsd[J];
            int ee[J];
            unsigned bad = 0u;
#pragma unroll
            for (int j = 0; j < J; ++j) { mm[j] = 1.0; ee[j] = 0; lsd[j] = 0.0; }
            for (int sg = wl; sg < nseg; sg += NWL) {
                const int tt = seg[3 * sg], first = seg[3 * sg + 1], stride = seg[3 * sg + 2];
                const int r0 = rel[tt], npairs = (rel[tt + 1] - r0) >> 1;
                const double2 *Wt = reinterpret_cast<const double2 *>(Wc + (size_t)C * r0);
                lik_segment<KC, J, false>(Wt, npairs, ncnt[tt], C, ff + (size_t)tt * C * JP, first + lane, stride, mm, ee, bad, lsd);
#pragma unroll
                for (int j = 0; j < J; ++j) {                                // a warp walking many segments stays in range
                    const int gg = __double2hiint(mm[j]);
                    ee[j] += (gg >> 20) - 1023;
                    mm[j] = dg_mant(mm[j], gg);
                }
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
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int gg = __double2hiint(mm[j]);
                    red[j * DG_MAXWARPS + wl] = dg_mant(mm[j], gg);
                    redi[j * DG_MAXWARPS + wl] = ee[j] + (gg >> 20) - 1023;
                }
                redb[wl] = bad;
            }
        } else {
            if (s_act) {
                const double se = dg_np_sum(ex + nd_j * CT + stt, C, Tc);
                psi_c = e_c / se;                                            // (:335)
                psic[tid] = psi_c;
            }
            // GP prior (:332) and proposal density (:330-331): x K^-1 x as (x K^-1)[t] * x[t]
            if (s_free) {
                const double *yc = ycand + nd_j * CT + sc * Tc;
                double r = 0.0;
                for (int j = 0; j < Tc; ++j) r = fma(yc[j], kinv[j * Tc + stt], r);
                pr[tid] = r * y_c;
                qq[tid] = base_c - y_c;                                      // d = Y_now - Y_try
            }
            bar_state(nstate);
            double qpart = 0.0;
            if (s_free) {
                const double *dc = qq + nd_j * CT + sc * Tc;
                double dq = 0.0;
                for (int j = 0; j < Tc; ++j) dq = fma(dc[j], kinv[j * Tc + stt], dq);
                qpart = dq * (base_c - y_c);
            }
            bar_state(nstate);                                               // everyone has read qq
            if (s_free) qq[tid] = qpart;
            bar_state(nstate);
            if (tid < J * (C - 1)) {                                         // lane (node, c): sums over t
                const int jn = tid / (C - 1), c = tid - jn * (C - 1);
                const double *prc = pr + jn * CT + c * Tc, *qc = qq + jn * CT + c * Tc;
                double a = 0.0, b = 0.0;
                for (int t = 0; t < Tc; ++t) { a += prc[t]; b += qc[t]; }
                gs[jn * CT + c] = prior_const - a / 2.0;                     // gs / ex are free after B1
                ex[jn * CT + c] = jcon[c] - (b * ijs[c]) / 2.0;
            }
            bar_state(nstate);
            if (tid < J) {
                double prior = 0.0, Q = 0.0;                                 // c ascending (:312-332)
                for (int c = 0; c < C - 1; ++c) { prior += gs[tid * CT + c]; Q += ex[tid * CT + c]; }
                bc[16 + tid] = prior;
                bc[16 + J + tid] = init ? 0.0 : Q;
            }
        }
        DG_TICK(1);
        __syncthreads();                                                     // B2: partial likelihoods ready
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
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        mm[j] = lane < NWL ? red[j * DG_MAXWARPS + lane] : 1.0;
                        ee[j] = lane < NWL ? redi[j * DG_MAXWARPS + lane] : 0;
                    }
                    bad = lane < NWL ? redb[lane] : 0u;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            mm[j] *= __shfl_xor_sync(0xffffffffu, mm[j], off);   // < 2^32
                            ee[j] += __shfl_xor_sync(0xffffffffu, ee[j], off);
                        }
                        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
                    }
#pragma unroll
                    for (int j = 0; j < J; ++j) lik[j] = fma((double)ee[j], 0.693147180559945309417232121458, dg_log(mm[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        double a = 0.0;
                        for (int w = 0; w < NWL; ++w) a += red[j * DG_MAXWARPS + w];
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
                            for (int d = 0; d < nd_max; ++d) {
                                const int step = m + d;
                                if (page_cur < 0 || (step & pmask) == 0)
                                    page_cur = trace_page(ptab, step, pshift, rowlen, ct.pool_head, ct.pool_cap);
                                if (page_cur < 0) { nd_ok = d; break; }
                                bc[32 + d] = __longlong_as_double(page_cur);
                            }
                            // walk the tree: alpha = exp(min(a, 0)), accept iff u < alpha (:348-351),
                            // tested in the log domain with log(u) taken off the critical path
                            int node = 0, cur = -1;
                            for (int d = 0; d < nd_ok; ++d) {
                                double likn = lik[0], priorn = bc[16 + node], Q = bc[16 + J + node];
#pragma unroll
                                for (int j = 1; j < J; ++j) if (j == node) likn = lik[j];
                                const double P_try = priorn + likn;
                                const double a = ((P_try + Q) - P_now) - Q;
                                const double lu = dring[((m + d) % RING) * CT + CT - 1];
                                const bool acc = (a >= 0.0) || (lu < a);     // NaN a: both false
                                if (!isfinite(P_try)) flags |= 1;
                                if (acc) { ++cnt; P_now = P_try; L_now = likn; cur = node; }
                                bc[40 + d] = (double)cur;                    // state after step m+d
                                bc[48 + d] = P_now;                          // Pro_all[m+d]  (:362)
                                bc[56 + d] = L_now;                          // Lik_all[m+d]  (:363)
                                node = 2 * node + (acc ? 2 : 1);
                            }
                            bc[4] = (double)nd_ok;
                            bc[5] = (double)cur;
                            bc[6] = (nd_ok < nd_max) ? 1.0 : 0.0;            // trace pool exhausted
                        }
                    }
                }
            }
            __syncthreads();                                                 // B3: decisions (or redo request)
            redo = bc[3] != 0.0;
            if (redo) {                                                      // rare: some x_n was not a positive normal
                if (!is_state) lik_warp_pass_log<KC, J>(Wc, seg, rel, ncnt, ff, red, C, NWL, lane, wl, nseg);
                __syncthreads();
            }
        } while (redo);
        DG_TICK(3);
        // ================= update + trace rows (state warps) || stream refill (others) ======
        const int nd = (int)bc[4];
        const int fin = (int)bc[5];
        const bool nomem = bc[6] != 0.0;
        if (is_state) {
            if (s_act && nd_j == 0) {                                        // primary lanes write the rows
                for (int d = 0; d < nd; ++d) {
                    const int node = (int)bc[40 + d];
                    const int step = m + d;
                    double *row = ct.trace + __double_as_longlong(bc[32 + d]) + (size_t)(step & pmask) * rowlen;
                    row[si] = node < 0 ? psi_now : psic[node * CT + si];     // Psi_all[m]  (:365)
                    row[CT + si] = node < 0 ? y_now : ycand[node * CT + si]; // Y_all[m]    (:364)
                    if (tid == 0) { row[2 * CT] = bc[48 + d]; row[2 * CT + 1] = bc[56 + d]; }
                }
            }
            if (s_act && (init || fin >= 0)) {
                const int f0 = init ? 0 : fin;
                y_now = ycand[f0 * CT + si];
                psi_now = psic[f0 * CT + si];
            }
            bar_state(nstate);                                               // candidates may be overwritten now
        } else if (!init && nd > 0) {
            // ring entries of the steps just consumed are free: refill them RING steps ahead
            produce_stream(dev_stream, min(m + RING, m_end), min(m + nd + RING, m_end), tid - nstate, nthreads - nstate,
                           false, zring, dring, RING, CT, CTm, Tc, pfac, sqs, sa.seed, sid, dblk, ublk, m_start);
        }
        if (init) { init = false; continue; }
        m += nd;
        if (nomem) {                                                         // trace pool exhausted (uniform)
            state = DICEGP_CHAIN_NOMEM_;
            m_ret = m - 1;
            break;
        }
        // convergence diagnostics (:369-391) after step m-1 if it is a check step; the rows
        // [0, m-1) exclude the one just written
        const int mlast = m - 1;
        if (nd > 0 && mlast >= initial && (mlast % gap) == 0) {
            __syncthreads();
            DG_TICK(4);
            const int conv = geweke_converged(gw, CTm, nwarps, warp, lane, tv, mlast);
            DG_TICK(5);
            if (conv) {
                state = DICEGP_CHAIN_CONVERGED_;
                m_ret = mlast;
                break;
            }
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
    __syncthreads();
    if (s_act && nd_j == 0) { st[si] = y_now; st[CT + si] = psi_now; }
    if (tid == 0) {
        st[2 * CT] = P_now; st[2 * CT + 1] = L_now;
        ct.m_next[chain] = m;
        ct.cnt[chain] = cnt;
        ct.flags[chain] = flags;
        ct.m_ret[chain] = m_ret;
        ct.state[chain] = state;
        if (W_IN_SMEM && use_tma)
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(dg_smem_u32(mbar)) : "memory");
    }
    __syncthreads();
}

}  // namespace dg
