// dicegp_chain.cuh -- the per-chain MH loop (one thread block per chain), kernel v1.
//
// Warp-specialised: the first S = ceil(C*Tc/32) warps are STATE warps, one thread per
// latent value y[c,t]; they own the serial part of a step (proposal, softmax / length
// reweighting, GP prior, MH decision, trace row).  The remaining warps are LIKELIHOOD
// warps; each owns fixed read slices of fixed time points, keeps that time point's
// isoform weights f in registers and accumulates prod(x_n) as (mantissa, exponent).
// Per step the two groups meet at three block barriers (B1: f is ready, B2: partial
// likelihoods are ready, B3: decision is known); the rest of the serial part overlaps
// with the likelihood pass, and the device-generated random stream for step m+2 is
// produced by the likelihood warps while the state warps write the trace row of step m.
#pragma once
#include "dicegp_kernels.cuh"

#ifdef DG_PROFILE
#define DG_TICK(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); dg_prof[i] += t__ - dg_t0; dg_t0 = t__; } } while (0)
#else
#define DG_TICK(i) do { } while (0)
#endif

namespace dg {

// named barrier 2: state warps only (barrier 0 = __syncthreads)
__device__ __forceinline__ void bar_state(int nstate) {
    asm volatile("bar.sync 2, %0;" ::"r"(nstate) : "memory");
}

// One likelihood segment: pairs p = p0, p0+stride, ... of one time point.
//   x_n = sum_k W[n,k] f[k]                                    model_GP.py:338
// W block: isoform-major, row k = npairs double2 (two reads per 16-byte load; consecutive
// lanes take consecutive pairs: conflict-free LDS.128 / coalesced LDG.128).
// prod(x) is carried as two mantissa products (one per read of the pair) and one
// exponent sum; anything that is not a positive normal double raises `bad`.
template <int KC, bool LOGPATH>
__device__ __forceinline__ void lik_segment(const double2 *__restrict__ Wt, int npairs, int n, int C,
                                            const double *fcol, int Tc, int p0, int stride, double &m0,
                                            double &m1, int &e, unsigned &bad, double &logsum) {
    constexpr int KR = KC > 0 ? KC : 1;
    double f[KR];
    if (KC > 0) {
#pragma unroll
        for (int k = 0; k < KR; ++k) f[k] = fcol[k * Tc];
    }
    int it = 0;
#pragma unroll 2
    for (int p = p0; p < npairs; p += stride) {
        double x0 = 0.0, x1 = 0.0;
        if (KC > 0) {
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const double2 w = Wt[(size_t)k * npairs + p];
                x0 = fma(w.x, f[k], x0);
                x1 = fma(w.y, f[k], x1);
            }
        } else {
            for (int k = 0; k < C; ++k) {
                const double2 w = Wt[(size_t)k * npairs + p];
                const double fk = fcol[k * Tc];
                x0 = fma(w.x, fk, x0);
                x1 = fma(w.y, fk, x1);
            }
        }
        const bool v1 = (2 * p + 1) < n;              // the pad slot of an odd block
        if (LOGPATH) {
            logsum += log(x0);
            if (v1) logsum += log(x1);
        } else {
            const double y1 = v1 ? x1 : 1.0;
            const int h0 = __double2hiint(x0), h1 = __double2hiint(y1);
            bad |= (((unsigned)(h0 - 0x00100000) >= 0x7fe00000u) || ((unsigned)(h1 - 0x00100000) >= 0x7fe00000u)) ? 1u : 0u;
            e += (h0 >> 20) + (h1 >> 20) - 2046;
            m0 *= __hiloint2double((h0 & 0x000fffff) | 0x3ff00000, __double2loint(x0));
            m1 *= __hiloint2double((h1 & 0x000fffff) | 0x3ff00000, __double2loint(y1));
            if (((++it) & 511) == 0) {                 // keep the running products in range
                const int g0 = __double2hiint(m0), g1 = __double2hiint(m1);
                e += (g0 >> 20) + (g1 >> 20) - 2046;
                m0 = __hiloint2double((g0 & 0x000fffff) | 0x3ff00000, __double2loint(m0));
                m1 = __hiloint2double((g1 & 0x000fffff) | 0x3ff00000, __double2loint(m1));
            }
        }
    }
}

// shapes and shared arrays of the chain a block is working on
struct ChainCtx {
    int C, Tc, CT, CTm;               // CTm = (C-1)*Tc free latent values
    int tid, nthreads, lane, warp, nwarps;
    int nstate, S, NWL;               // state threads (multiple of 32), state warps, likelihood warps
    double *ytry, *fsi, *ex, *gs, *pr, *qq, *zb;
    double *kinv, *pfac, *ijs, *jcon, *sqs;
    double *red, *bc, *gw;
    int *redi, *rel, *ncnt, *seg;
    unsigned *redb;
    const double *Wc;                 // shared or global
    double prior_const;
};

// Likelihood pass of one likelihood warp over its segments; leaves (mantissa, exponent,
// bad) of the warp in red/redi/redb[wl].
template <int KC>
__device__ __forceinline__ void lik_warp_pass(const ChainCtx &s, int wl, int nseg) {
    double m0 = 1.0, m1 = 1.0, dummy = 0.0;
    int e = 0;
    unsigned bad = 0u;
    for (int sg = wl; sg < nseg; sg += s.NWL) {
        const int tt = s.seg[3 * sg], first = s.seg[3 * sg + 1], stride = s.seg[3 * sg + 2];
        const int r0 = s.rel[tt], npairs = (s.rel[tt + 1] - r0) >> 1;
        const double2 *Wt = reinterpret_cast<const double2 *>(s.Wc + (size_t)s.C * r0);
        lik_segment<KC, false>(Wt, npairs, s.ncnt[tt], s.C, s.fsi + tt, s.Tc, first + s.lane, stride, m0, m1, e, bad, dummy);
        // a warp that walks many segments keeps its products in range
        const int g0 = __double2hiint(m0), g1 = __double2hiint(m1);
        e += (g0 >> 20) + (g1 >> 20) - 2046;
        m0 = __hiloint2double((g0 & 0x000fffff) | 0x3ff00000, __double2loint(m0));
        m1 = __hiloint2double((g1 & 0x000fffff) | 0x3ff00000, __double2loint(m1));
    }
    double m = m0 * m1;                                     // < 4
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        m *= __shfl_xor_sync(0xffffffffu, m, off);          // < 4^32 = 2^64
        e += __shfl_xor_sync(0xffffffffu, e, off);
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if (s.lane == 0) {
        const int g = __double2hiint(m);
        s.red[wl] = __hiloint2double((g & 0x000fffff) | 0x3ff00000, __double2loint(m));
        s.redi[wl] = e + (g >> 20) - 1023;
        s.redb[wl] = bad;
    }
}

// Slow path (some x_n is 0 / negative / non-finite): the same sum with real logarithms so
// that -inf and NaN propagate like np.log(...).sum() does (model_GP.py:344).
template <int KC>
__device__ __noinline__ void lik_warp_pass_log(const double *Wc, const int *seg, const int *rel, const int *ncnt,
                                               const double *fsi, double *red, int C, int Tc, int NWL, int lane,
                                               int wl, int nseg) {
    double m0 = 1.0, m1 = 1.0, ls = 0.0;
    int e = 0;
    unsigned bad = 0u;
    for (int sg = wl; sg < nseg; sg += NWL) {
        const int tt = seg[3 * sg], first = seg[3 * sg + 1], stride = seg[3 * sg + 2];
        const int r0 = rel[tt], npairs = (rel[tt + 1] - r0) >> 1;
        const double2 *Wt = reinterpret_cast<const double2 *>(Wc + (size_t)C * r0);
        lik_segment<KC, true>(Wt, npairs, ncnt[tt], C, fsi + tt, Tc, first + lane, stride, m0, m1, e, bad, ls);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, off);
    if (lane == 0) red[wl] = ls;
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

// ---------------------------------------------------------------------------------------
// one chain, up to n_steps steps.  `slot` indexes the host stream blocks.
// ---------------------------------------------------------------------------------------
template <int KC, bool W_IN_SMEM>
__device__ void run_chain(const GeneTab &gt, const ChainTab &ct, const StreamArgs &sa, int chain,
                          int slot, int n_steps, double *smem, int use_tma) {
    ChainCtx s;
    s.tid = threadIdx.x; s.nthreads = blockDim.x;
    s.lane = s.tid & 31; s.warp = s.tid >> 5; s.nwarps = s.nthreads >> 5;
    if (ct.state[chain] != 0) return;                    // already finished (block-uniform)

    const int g = ct.gene[chain], t0 = ct.t0[chain];
    const int C = gt.C[g], Tc = ct.Tc[chain], T = gt.T;
    s.C = C; s.Tc = Tc; s.CT = C * Tc; s.CTm = (C - 1) * Tc;
    s.S = (s.CT + 31) >> 5; s.nstate = s.S << 5; s.NWL = s.nwarps - s.S;
    const SmLayout L = dg_layout(C, Tc, s.nwarps);
    s.ytry = smem + L.ytry; s.fsi = smem + L.fsi; s.ex = smem + L.ex; s.gs = smem + L.gs;
    s.pr = smem + L.pr; s.qq = smem + L.qq; s.zb = smem + L.zb; s.kinv = smem + L.kinv;
    s.pfac = smem + L.pfac; s.ijs = smem + L.ijs; s.jcon = smem + L.jcon;
    s.sqs = smem + L.sqs; s.red = smem + L.red; s.redi = reinterpret_cast<int *>(smem + L.redi);
    s.redb = reinterpret_cast<unsigned *>(smem + L.redb); s.bc = smem + L.bc; s.gw = smem + L.gw;
    s.rel = reinterpret_cast<int *>(smem + L.rel); s.ncnt = s.rel + (Tc + 1);
    s.seg = reinterpret_cast<int *>(smem + L.seg);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + L.mbar);
    double *Wsm = smem + L.w;
    s.prior_const = ct.prior_const[chain];
    const bool is_state = s.tid < s.nstate;
    const int wl = s.warp - s.S;                          // likelihood-warp index (>= 0 for those)

    const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t0;
    const int r_first = roff[0];
    const int r_total = roff[Tc] - r_first;
    const double *Wg = gt.W + gt.woff[g] + (size_t)C * r_first;
    const size_t wbytes = (size_t)C * r_total * sizeof(double);

    // ---- stage W (TMA bulk copy, asynchronous) and the small per-chain constants --------
    if (W_IN_SMEM) {
        if (use_tma) {
            if (s.tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dg_smem_u32(mbar)));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (s.tid == 0 && wbytes > 0) {
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
            for (size_t i = s.tid; i < n2; i += s.nthreads) dst[i] = src[i];
        }
        s.Wc = Wsm;
    } else {
        s.Wc = Wg;
    }
    for (int i = s.tid; i <= Tc; i += s.nthreads) s.rel[i] = roff[i] - r_first;
    for (int i = s.tid; i < Tc; i += s.nthreads) s.ncnt[i] = gt.ncnt[(size_t)g * T + t0 + i];
    {
        const double *pool = ct.pool;
        const double *kinv = pool + ct.kinv_off[chain];
        for (int i = s.tid; i < Tc * Tc; i += s.nthreads) s.kinv[i] = kinv[i];
        if (ct.pf_off[chain] >= 0) {
            const double *pf = pool + ct.pf_off[chain];
            for (int i = s.tid; i < Tc * Tc; i += s.nthreads) s.pfac[i] = pf[i];
        }
        const double *js = pool + ct.js_off[chain], *jc = pool + ct.jc_off[chain];
        for (int i = s.tid; i < C - 1; i += s.nthreads) {
            s.ijs[i] = 1.0 / js[i];
            s.sqs[i] = sqrt(js[i]);
            s.jcon[i] = jc[i];
        }
    }
    const int M = ct.M[chain], initial = ct.initial[chain], gap = ct.gap[chain];
    int m = ct.m_next[chain];
    double *st = ct.st + ct.st_off[chain];
    const int rowlen = 2 * s.CT + 2;
    int64_t *ptab = ct.ptab + ct.ptab_off[chain];
    const int pshift = ct.pshift[chain];
    const int pmask = (1 << pshift) - 1;
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
    const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;
    const bool dev_stream = sa.mode != 0;

    // per state thread: its latent value (c, t); current state lives in registers
    const int si = s.tid;                                 // state index = c*Tc + t
    const bool s_act = is_state && si < s.CT;             // owns a value
    const bool s_free = is_state && si < s.CTm;           // value is sampled (c < C-1)
    const int sc = s_act ? si / Tc : 0;
    const int stt = s_act ? si - sc * Tc : 0;
    double y_now = 0.0, psi_now = 0.0;
    double P_now = 0.0, L_now = 0.0;                      // thread 0
    int cnt = 0, flags = 0;
    if (s.tid == 0) { cnt = ct.cnt[chain]; flags = ct.flags[chain]; }
    const double len_own = s_act ? gt.len[gt.lenoff[g] + (size_t)(t0 + stt) * C + sc] : 0.0;

    __syncthreads();                                      // rel / ncnt visible
    if (s.tid == 0) s.bc[6] = (double)build_segments(s.seg, s.rel, Tc, s.NWL);
    if (m > 0 && s_act) { y_now = st[si]; psi_now = st[s.CT + si]; }
    if (m > 0 && s.tid == 0) { P_now = st[2 * s.CT]; L_now = st[2 * s.CT + 1]; }
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
    const int nseg = (int)s.bc[6];

    const int m_start = m;
    const int m_end = (m_start + n_steps < M) ? (m_start + n_steps) : M;
    int state = 0, m_ret = 0;
    double *page = nullptr;                               // state threads: page of the row being written

    // ---- stream set-up ---------------------------------------------------------------------
    const double *dblk = nullptr, *ublk = nullptr;
    double d_next = 0.0, lu_next = 0.0;
    if (!dev_stream) {
        dblk = sa.delta + sa.delta_off[slot];
        ublk = sa.u + sa.u_off[slot];
        if (m < m_end) {
            if (s_free) d_next = dblk[si];
            if (s.tid == 0) lu_next = log(ublk[0]);
        }
    } else {
        // z and log(u) of the first two steps (later ones are produced two steps ahead by
        // the likelihood warps); slot CT-1 of a z buffer holds log(u)
        for (int k = 0; k < 2; ++k) {
            const int mm = m + k;
            for (int i = s.tid; i < s.CTm; i += s.nthreads)
                s.zb[(mm & 1) * s.CT + i] = dg_normal(sa.seed, sid, (uint32_t)mm, (uint32_t)i);
            if (s.tid == s.nthreads - 1) s.zb[(mm & 1) * s.CT + s.CT - 1] = log(dg_uniform(sa.seed, sid, (uint32_t)mm));
        }
        __syncthreads();
    }
#ifdef DG_PROFILE
    long long dg_prof[10];
    long long dg_t0 = clock64();
    for (int i = 0; i < 10; ++i) dg_prof[i] = 0;
#endif

    // it == -1: evaluate the start value (model_GP.py:275-292) with the same machinery as a
    // step (always "accepted", no trace row), then steps m .. m_end-1
    for (int it = (m == 0 ? -1 : m); it < m_end; ++it) {
        const bool init = it < 0;
        const int mcur = init ? 0 : it;
        double y_try = 0.0, e_try = 1.0, g_own = 0.0, psi_try = 0.0, lu_cur = 0.0;
        // ================= S phase (state warps): proposal -> f =================
        if (is_state) {
            if (init) {
                const double *y0 = ct.y0_off[chain] >= 0 ? ct.pool + ct.y0_off[chain] : nullptr;
                if (s_act) y_try = y0 ? y0[si] : 0.0;                       // Y_now = Ymean (:275)
            } else if (s_free) {
                double d;
                if (!dev_stream) {
                    d = d_next;
                } else {
                    const double *z = s.zb + (mcur & 1) * s.CT + sc * Tc;
                    double a = 0.0;
                    for (int j = 0; j < Tc; ++j) a = fma(z[j], s.pfac[j * Tc + stt], a);
                    d = s.sqs[sc] * a;
                }
                y_try = d + y_now;                                           // (:329) numpy adds the mean last
            }                                                                // else Y_try[C-1,:] stays 0 (:295)
            if (s.tid == 0 && !init) lu_cur = dev_stream ? s.zb[(mcur & 1) * s.CT + s.CT - 1] : lu_next;
            if (s_act) {
                e_try = exp(y_try);
                g_own = len_own * e_try;
                s.ytry[si] = y_try;
                s.ex[si] = e_try;
                s.gs[si] = g_own;
            }
            bar_state(s.nstate);
            if (s_act) {
                // fsi = len*psi / sum(len*psi) with psi = e/sum(e): the common 1/sum(e) cancels
                // (model_GP.py:335-337; one division on the critical path instead of two)
                const double sg = dg_np_sum(s.gs + stt, C, Tc);
                s.fsi[si] = g_own / sg;
            }
        }
        __syncthreads();                                                     // B1: f ready
        DG_TICK(0);
        // ================= likelihood warps || rest of the serial part =================
        if (!is_state) {
            lik_warp_pass<KC>(s, wl, nseg);
        } else {
            if (!dev_stream && !init && mcur + 1 < m_end) {                  // host stream: prefetch
                if (s_free) d_next = dblk[(size_t)(mcur + 1 - m_start) * s.CTm + si];
                if (s.tid == 0) lu_next = log(ublk[mcur + 1 - m_start]);
            }
            if (s_act) {
                const double se = dg_np_sum(s.ex + stt, C, Tc);
                psi_try = e_try / se;                                        // (:335)
            }
            // GP prior (:332) and proposal density (:330-331): x K^-1 x as (x K^-1)[t] * x[t]
            if (s_free) {
                const double *yc = s.ytry + sc * Tc;
                double r = 0.0;
                for (int j = 0; j < Tc; ++j) r = fma(yc[j], s.kinv[j * Tc + stt], r);
                s.pr[si] = r * y_try;
                s.qq[si] = y_now - y_try;                                    // d = Y_now - Y_try
            }
            bar_state(s.nstate);
            double qpart = 0.0;
            if (s_free && !init) {
                const double *dc = s.qq + sc * Tc;
                double dq = 0.0;
                for (int j = 0; j < Tc; ++j) dq = fma(dc[j], s.kinv[j * Tc + stt], dq);
                qpart = dq * (y_now - y_try);
            }
            bar_state(s.nstate);                                             // everyone has read qq
            if (s_free) s.qq[si] = qpart;
            bar_state(s.nstate);
            if (s.tid < C - 1) {
                const int c = s.tid;
                double a = 0.0, b = 0.0;
                for (int t = 0; t < Tc; ++t) { a += s.pr[c * Tc + t]; b += s.qq[c * Tc + t]; }
                s.pr[c * Tc] = s.prior_const - a / 2.0;
                s.qq[c * Tc] = s.jcon[c] - (b * s.ijs[c]) / 2.0;
            }
            bar_state(s.nstate);
            if (s.tid == 0) {
                double prior = 0.0, Q = 0.0;                                 // c ascending (:312-332)
                for (int c = 0; c < C - 1; ++c) {
                    prior += s.pr[c * Tc];
                    if (!init) Q += s.qq[c * Tc];
                }
                s.bc[0] = prior;
                s.bc[1] = Q;
            }
        }
        DG_TICK(1);
        __syncthreads();                                                     // B2: partial likelihoods ready
        DG_TICK(2);
        // ================= decision (warp 0; thread 0 decides) =================
        bool redo = false;
        do {
            if (s.warp == 0) {
                double lik;
                unsigned bad = 0u;
                if (!redo) {
                    double mm = s.lane < s.NWL ? s.red[s.lane] : 1.0;
                    int ee = s.lane < s.NWL ? s.redi[s.lane] : 0;
                    bad = s.lane < s.NWL ? s.redb[s.lane] : 0u;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        mm *= __shfl_xor_sync(0xffffffffu, mm, off);         // < 2^32
                        ee += __shfl_xor_sync(0xffffffffu, ee, off);
                        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
                    }
                    lik = fma((double)ee, 0.693147180559945309417232121458, log(mm));
                } else {
                    lik = 0.0;
                    for (int w = 0; w < s.NWL; ++w) lik += s.red[w];
                }
                if (s.lane == 0) {
                    s.bc[3] = bad ? 1.0 : 0.0;
                    if (!bad) {
                        const double Q = s.bc[1];
                        const double P_try = s.bc[0] + lik;
                        bool acc;
                        if (init) {
                            acc = true;                                      // P_now / L_now of the start value
                        } else {
                            // alpha = exp(min(a, 0)), accept iff u < alpha (:348-351): tested in the
                            // log domain, log(u) having been taken off the critical path
                            const double a = ((P_try + Q) - P_now) - Q;
                            acc = (a >= 0.0) || (lu_cur < a);                // NaN a: both false
                            if (acc) ++cnt;
                        }
                        if (!isfinite(P_try)) flags |= 1;
                        if (acc) { P_now = P_try; L_now = lik; }
                        // trace page of row m (allocated on first touch; the first page always exists)
                        long long pg = 0;
                        if (!init && (page == nullptr || (mcur & pmask) == 0)) {
                            pg = __ldcg(ptab + (mcur >> pshift));
                            if (pg < 0) {
                                const unsigned long long need = (unsigned long long)rowlen << pshift;
                                const unsigned long long off = atomicAdd(ct.pool_head, need);
                                if (off + need <= (unsigned long long)ct.pool_cap) {
                                    pg = (long long)off;
                                    ptab[mcur >> pshift] = pg;
                                }
                            }
                        }
                        s.bc[4] = acc ? 1.0 : 0.0;
                        s.bc[5] = __longlong_as_double(pg);
                    }
                }
            }
            __syncthreads();                                                 // B3: decision (or redo request)
            redo = s.bc[3] != 0.0;
            if (redo) {                                                      // rare: some x_n was not a positive normal
                if (!is_state) lik_warp_pass_log<KC>(s.Wc, s.seg, s.rel, s.ncnt, s.fsi, s.red, C, Tc, s.NWL, s.lane, wl, nseg);
                __syncthreads();
            }
        } while (redo);
        DG_TICK(3);
        // ================= update + trace row (state warps) || stream production (others) ======
        const long long pg_new = __double_as_longlong(s.bc[5]);
        if (is_state) {
            if (s.bc[4] != 0.0 && s_act) { y_now = y_try; psi_now = psi_try; }
            if (!init) {
                if (page == nullptr || (mcur & pmask) == 0) page = pg_new >= 0 ? ct.trace + pg_new : nullptr;
                if (page != nullptr) {
                    double *row = page + (size_t)(mcur & pmask) * rowlen;
                    if (s_act) {
                        row[si] = psi_now;                                   // Psi_all[m]  (:365)
                        row[s.CT + si] = y_now;                              // Y_all[m]    (:364)
                    }
                    if (s.tid == 0) {
                        row[2 * s.CT] = P_now;                               // Pro_all[m]  (:362)
                        row[2 * s.CT + 1] = L_now;                           // Lik_all[m]  (:363)
                    }
                }
            }
        } else if (dev_stream && !init && mcur + 2 < m_end) {
            const int mm = mcur + 2;                                         // buffer released by step m
            for (int i = s.tid - s.nstate; i < s.CTm; i += s.nthreads - s.nstate)
                s.zb[(mm & 1) * s.CT + i] = dg_normal(sa.seed, sid, (uint32_t)mm, (uint32_t)i);
            if (s.tid == s.nthreads - 1) s.zb[(mm & 1) * s.CT + s.CT - 1] = log(dg_uniform(sa.seed, sid, (uint32_t)mm));
        }
        if (init) continue;
        if (pg_new < 0) {                                                    // trace pool exhausted (uniform)
            state = DICEGP_CHAIN_NOMEM_;
            m_ret = mcur - 1;
            m = mcur;
            break;
        }
        m = mcur + 1;
        // convergence diagnostics (:369-391); rows [0, m) exclude the one just written
        if (mcur >= initial && (mcur % gap) == 0) {
            __syncthreads();
            DG_TICK(4);
            const int conv = geweke_converged(s.gw, s.CTm, s.nwarps, s.warp, s.lane, tv, mcur);
            DG_TICK(5);
            if (conv) {
                state = DICEGP_CHAIN_CONVERGED_;
                m_ret = mcur;
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
    if (s_act) { st[si] = y_now; st[s.CT + si] = psi_now; }
    if (s.tid == 0) {
        st[2 * s.CT] = P_now; st[2 * s.CT + 1] = L_now;
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
