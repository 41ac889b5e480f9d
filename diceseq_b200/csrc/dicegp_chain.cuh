// dicegp_chain.cuh -- the per-chain MH loop (one thread block per chain).
#pragma once
#include "dicegp_kernels.cuh"

#ifdef DG_PROFILE
#define DG_TICK(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); dg_prof[i] += t__ - dg_t0; dg_t0 = t__; } } while (0)
#define DG_PROF_ARGS , long long *dg_prof, long long &dg_t0
#define DG_PROF_PASS , dg_prof, dg_t0
#else
#define DG_TICK(i) do { } while (0)
#define DG_PROF_ARGS
#define DG_PROF_PASS
#endif

namespace dg {

// ---------------------------------------------------------------------------------------
// read likelihood: sum_t sum_n log(sum_k W_t[n,k] f_t[k])        model_GP.py:338,344-345
// W block of time tt: isoform-major, row k = npad doubles, npad even.  Each thread takes
// read PAIRS (one 16-byte load per isoform), consecutive threads consecutive pairs:
// conflict-free LDS.128 / fully coalesced LDG.128.
// ---------------------------------------------------------------------------------------
template <int KC, bool LOGPATH>
__device__ __forceinline__ void lik_partial(const double *__restrict__ Wc, int C, int Tc,
                                            const int *rel, const int *ncnt,
                                            const double *fsi, int tid, int nthreads,
                                            MantExp &acc, double &logsum) {
    constexpr int KR = KC > 0 ? KC : 1;
    int it = 0;
    for (int tt = 0; tt < Tc; ++tt) {
        const int r0 = rel[tt];
        const int npairs = (rel[tt + 1] - r0) >> 1;
        const int n = ncnt[tt];
        const double2 *__restrict__ Wt = reinterpret_cast<const double2 *>(Wc + (size_t)C * r0);
        double f[KR];
        if (KC > 0) {
#pragma unroll
            for (int k = 0; k < KR; ++k) f[k] = fsi[k * Tc + tt];
        }
        for (int p = tid; p < npairs; p += nthreads) {
            double x0 = 0.0, x1 = 0.0;
            if (KC > 0) {
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    double2 w = Wt[(size_t)k * npairs + p];
                    x0 = fma(w.x, f[k], x0);
                    x1 = fma(w.y, f[k], x1);
                }
            } else {
                for (int k = 0; k < C; ++k) {
                    double2 w = Wt[(size_t)k * npairs + p];
                    double fk = fsi[k * Tc + tt];
                    x0 = fma(w.x, fk, x0);
                    x1 = fma(w.y, fk, x1);
                }
            }
            const bool v1 = (2 * p + 1) < n;      // the pad slot of an odd block
            if (LOGPATH) {
                logsum += log(x0);
                if (v1) logsum += log(x1);
            } else {
                dg_acc(acc, x0);
                dg_acc(acc, v1 ? x1 : 1.0);
                if (((++it) & 127) == 0) dg_renorm(acc);
            }
        }
    }
}

template <bool LOGPATH>
__device__ __forceinline__ void lik_dispatch(const double *Wc, int C, int Tc, const int *rel,
                                             const int *ncnt, const double *fsi, int tid, int nthreads,
                                             MantExp &acc, double &logsum) {
    switch (C) {
        case 1: lik_partial<1, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 2: lik_partial<2, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 3: lik_partial<3, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 4: lik_partial<4, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 5: lik_partial<5, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 6: lik_partial<6, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 7: lik_partial<7, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        case 8: lik_partial<8, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
        default: lik_partial<0, LOGPATH>(Wc, C, Tc, rel, ncnt, fsi, tid, nthreads, acc, logsum); break;
    }
}

// ---------------------------------------------------------------------------------------
// block-wide state of one chain
// ---------------------------------------------------------------------------------------
struct ChainCtx {
    // shapes
    int C, Tc, CT, CTm;          // CTm = (C-1)*Tc free latent values
    int tid, nthreads, lane, warp, nwarps;
    // shared arrays
    double *ynow, *ytry, *psinow, *psitry, *fsi, *ex, *pr, *qq, *zb;
    double *kinv, *pfac, *lens, *ijs, *jcon, *sqs;
    double *red, *bc, *gw;
    int *redi, *rel, *ncnt;
    unsigned *redb;
    const double *Wc;            // shared or global
    double prior_const;
};

// Full evaluation of the candidate in s.ytry: softmax / length reweighting
// (model_GP.py:335-337), GP prior (:332, normal_pdf :89-91), proposal density (:330-331)
// and read likelihood (:338-345).  Results land in bc[0]=prior, bc[1]=Q, bc[2]=lik.
// Every thread of the block must call it.
__device__ __forceinline__ void evaluate(ChainCtx &s, bool with_q DG_PROF_ARGS) {
    const int C = s.C, Tc = s.Tc, CT = s.CT, tid = s.tid;
    // phase 2: per (c,t): quadratic-form partials and exp
    if (tid < CT) {
        const int c = tid / Tc, t = tid - c * Tc;
        const double *yc = s.ytry + c * Tc;
        s.ex[tid] = exp(yc[t]);
        if (c < C - 1) {
            double r = 0.0;
            for (int j = 0; j < Tc; ++j) r = fma(yc[j], s.kinv[j * Tc + t], r);
            s.pr[tid] = r * yc[t];
            if (with_q) {
                const double *yn = s.ynow + c * Tc;
                double dq = 0.0;
                for (int j = 0; j < Tc; ++j) dq = fma(yn[j] - yc[j], s.kinv[j * Tc + t], dq);
                s.qq[tid] = dq * (yn[t] - yc[t]);
            }
        }
    }
    __syncthreads();
    DG_TICK(1);
    // phase 3: per time point softmax + reweight; one thread sums prior and Q
    if (tid < Tc) {
        const int t = tid;
        const double se = dg_np_sum(s.ex + t, C, Tc);
        for (int k = 0; k < C; ++k) {
            const double p = s.ex[k * Tc + t] / se;
            s.psitry[k * Tc + t] = p;
            s.fsi[k * Tc + t] = s.lens[t * C + k] * p;          // g_k, normalised below
        }
        const double sg = dg_np_sum(s.fsi + t, C, Tc);
        for (int k = 0; k < C; ++k) s.fsi[k * Tc + t] = s.fsi[k * Tc + t] / sg;
    } else if (tid == 32) {
        // warp 1 (idle in phase 3 when Tc <= 32): prior and proposal density, c ascending
        double prior = 0.0, Q = 0.0;
        for (int c = 0; c < C - 1; ++c) {
            double a = 0.0;
            for (int t = 0; t < Tc; ++t) a += s.pr[c * Tc + t];
            prior += s.prior_const - a / 2.0;
            if (with_q) {
                double b = 0.0;
                for (int t = 0; t < Tc; ++t) b += s.qq[c * Tc + t];
                Q += s.jcon[c] - (b * s.ijs[c]) / 2.0;
            }
        }
        s.bc[0] = prior;
        s.bc[1] = Q;
    }
    __syncthreads();
    DG_TICK(2);
    // phase 4: likelihood partials
    MantExp acc;
    acc.m = 1.0; acc.e = 0; acc.bad = 0u;
    double dummy = 0.0;
    lik_dispatch<false>(s.Wc, C, Tc, s.rel, s.ncnt, s.fsi, tid, s.nthreads, acc, dummy);
    dg_renorm(acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        acc.m *= __shfl_xor_sync(0xffffffffu, acc.m, off);
        acc.e += __shfl_xor_sync(0xffffffffu, acc.e, off);
        acc.bad |= __shfl_xor_sync(0xffffffffu, acc.bad, off);
    }
    if (s.lane == 0) {
        dg_renorm(acc);
        s.red[s.warp] = acc.m;
        s.redi[s.warp] = acc.e;
        s.redb[s.warp] = acc.bad;
    }
    DG_TICK(3);
    __syncthreads();
    DG_TICK(4);
    if (tid == 0) {
        double m = 1.0;
        long long e = 0;
        unsigned bad = 0u;
        for (int w = 0; w < s.nwarps; ++w) { m *= s.red[w]; e += s.redi[w]; bad |= s.redb[w]; }
        s.bc[2] = fma((double)e, 0.693147180559945309417232121458, log(m));
        s.bc[3] = bad ? 1.0 : 0.0;
    }
    __syncthreads();
    DG_TICK(5);
    if (s.bc[3] != 0.0) {
        // slow path: some read has a non-positive / non-finite likelihood term
        double ls = 0.0;
        lik_dispatch<true>(s.Wc, C, Tc, s.rel, s.ncnt, s.fsi, tid, s.nthreads, acc, ls);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, off);
        __syncthreads();                       // everyone has read bc[3]
        if (s.lane == 0) s.red[s.warp] = ls;
        __syncthreads();
        if (tid == 0) {
            double a = 0.0;
            for (int w = 0; w < s.nwarps; ++w) a += s.red[w];
            s.bc[2] = a;
            s.bc[3] = 0.0;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Geweke convergence rule over trace rows [0, m)          model_GP.py:156-185, :369-391
// Series = Psi_all[:m, c, t] for c < C-1 (columns 0..CTm-1 of a trace row).
// Two-pass mean / population variance like np.mean / np.var.  Lanes walk columns
// (coalesced), warps walk rows.  Returns 1 to every thread when all |Z| <= 2.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int geweke_converged(ChainCtx &s, const TraceView &tv, int m) {
    const int nA = (int)(0.1 * (double)m);
    const int iB = (int)(0.5 * (double)m);
    const int nB = m - iB;
    int ok = 1;
    for (int c0 = 0; c0 < s.CTm; c0 += DG_GW_COLS) {
        const int col = c0 + s.lane;
        const bool act = col < s.CTm;
        double *gA = s.gw, *gB = s.gw + s.nwarps * DG_GW_COLS;
        // pass 1: sums
        double sa = 0.0, sb = 0.0;
        if (act) {
            for (int r = s.warp; r < nA; r += s.nwarps) sa += __ldcg(tv.row(r) + col);
            for (int r = iB + s.warp; r < m; r += s.nwarps) sb += __ldcg(tv.row(r) + col);
        }
        gA[s.warp * DG_GW_COLS + s.lane] = sa;
        gB[s.warp * DG_GW_COLS + s.lane] = sb;
        __syncthreads();
        double meanA = 0.0, meanB = 0.0;
        for (int w = 0; w < s.nwarps; ++w) {
            meanA += gA[w * DG_GW_COLS + s.lane];
            meanB += gB[w * DG_GW_COLS + s.lane];
        }
        meanA /= (double)nA;
        meanB /= (double)nB;
        __syncthreads();
        // pass 2: centred sums of squares
        double qa = 0.0, qb = 0.0;
        if (act) {
            for (int r = s.warp; r < nA; r += s.nwarps) {
                double d = __ldcg(tv.row(r) + col) - meanA;
                qa = fma(d, d, qa);
            }
            for (int r = iB + s.warp; r < m; r += s.nwarps) {
                double d = __ldcg(tv.row(r) + col) - meanB;
                qb = fma(d, d, qb);
            }
        }
        gA[s.warp * DG_GW_COLS + s.lane] = qa;
        gB[s.warp * DG_GW_COLS + s.lane] = qb;
        __syncthreads();
        if (s.warp == 0 && act) {
            double va = 0.0, vb = 0.0;
            for (int w = 0; w < s.nwarps; ++w) {
                va += gA[w * DG_GW_COLS + s.lane];
                vb += gB[w * DG_GW_COLS + s.lane];
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

// ---------------------------------------------------------------------------------------
// one chain, up to n_steps steps.  `slot` indexes the host stream blocks.
// ---------------------------------------------------------------------------------------
template <bool W_IN_SMEM>
__device__ void run_chain(const GeneTab &gt, const ChainTab &ct, const StreamArgs &sa, int chain,
                          int slot, int n_steps, double *smem, int use_tma) {
    ChainCtx s;
    s.tid = threadIdx.x; s.nthreads = blockDim.x;
    s.lane = s.tid & 31; s.warp = s.tid >> 5; s.nwarps = (s.nthreads + 31) >> 5;
    if (ct.state[chain] != 0) return;                    // already finished (block-uniform)

    const int g = ct.gene[chain], t0 = ct.t0[chain];
    const int C = gt.C[g], Tc = ct.Tc[chain], T = gt.T;
    s.C = C; s.Tc = Tc; s.CT = C * Tc; s.CTm = (C - 1) * Tc;
    const SmLayout L = dg_layout(C, Tc, s.nwarps);
    s.ynow = smem + L.ynow; s.ytry = smem + L.ytry; s.psinow = smem + L.psinow;
    s.psitry = smem + L.psitry; s.fsi = smem + L.fsi; s.ex = smem + L.ex; s.pr = smem + L.pr;
    s.qq = smem + L.qq; s.zb = smem + L.zb; s.kinv = smem + L.kinv; s.pfac = smem + L.pfac;
    s.lens = smem + L.lens; s.ijs = smem + L.ijs; s.jcon = smem + L.jcon; s.sqs = smem + L.sqs;
    s.red = smem + L.red; s.redi = reinterpret_cast<int *>(smem + L.redi);
    s.redb = reinterpret_cast<unsigned *>(smem + L.redb); s.bc = smem + L.bc; s.gw = smem + L.gw;
    s.rel = reinterpret_cast<int *>(smem + L.rel); s.ncnt = s.rel + (Tc + 1);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem + L.mbar);
    double *Wsm = smem + L.w;
    s.prior_const = ct.prior_const[chain];

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
                    uint32_t sz = (uint32_t)((wbytes - done) < kChunk ? (wbytes - done) : kChunk);
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
        const double *lens = gt.len + gt.lenoff[g] + (size_t)t0 * C;
        for (int i = s.tid; i < Tc * C; i += s.nthreads) s.lens[i] = lens[i];
    }
#ifdef DG_PROFILE
    long long dg_prof[10];
    long long dg_t0 = 0;
    long long *dg_prof_out = reinterpret_cast<long long *>(const_cast<double *>(sa.prof));
#endif
    const int M = ct.M[chain], initial = ct.initial[chain], gap = ct.gap[chain];
    int m = ct.m_next[chain];
    int cnt = ct.cnt[chain];
    int flags = ct.flags[chain];
    double *st = ct.st + ct.st_off[chain];
    double P_now = 0.0, L_now = 0.0;                      // live in thread 0
    const int rowlen = 2 * s.CT + 2;
    int64_t *ptab = ct.ptab + ct.ptab_off[chain];
    const int pshift = ct.pshift[chain];
    const int pmask = (1 << pshift) - 1;
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
    double *page = nullptr;                               // base of the page holding row m
    const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;

    if (m == 0) {
        const double *y0 = ct.y0_off[chain] >= 0 ? ct.pool + ct.y0_off[chain] : nullptr;
        if (s.tid < s.CT) s.ytry[s.tid] = y0 ? y0[s.tid] : 0.0;          // Y_now = Ymean (:275)
    } else {
        if (s.tid < s.CT) { s.ynow[s.tid] = st[s.tid]; s.psinow[s.tid] = st[s.CT + s.tid]; }
        if (s.tid == 0) { P_now = st[2 * s.CT]; L_now = st[2 * s.CT + 1]; }
    }
    if (W_IN_SMEM && use_tma && wbytes > 0) {
        // wait for the bulk copy (phase 0)
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(dg_smem_u32(mbar)), "r"(0u) : "memory");
        }
    }
    __syncthreads();

    if (m == 0) {
        // initial state: P_now / L_now of the start value (model_GP.py:278-292)
        evaluate(s, false DG_PROF_PASS);
        if (s.tid < s.CT) {
            s.ynow[s.tid] = s.ytry[s.tid];
            s.psinow[s.tid] = s.psitry[s.tid];
        }
        if (s.tid == 0) {
            L_now = s.bc[2];
            P_now = s.bc[0] + s.bc[2];
            if (!isfinite(P_now)) flags |= 1;
        }
        __syncthreads();
        if (s.tid >= s.CTm && s.tid < s.CT) s.ytry[s.tid] = 0.0;          // Y_try[C-1,:] stays 0 (:295)
        __syncthreads();
    }

    int state = 0, m_ret = 0;
    // ---- stream set-up ---------------------------------------------------------------------
    const int m_start = m;
    const int m_end = (m_start + n_steps < M) ? (m_start + n_steps) : M;
    const double *dblk = nullptr, *ublk = nullptr;
    double d_next = 0.0, u_next = 0.0;
    if (sa.mode == 0) {
        dblk = sa.delta + sa.delta_off[slot];
        ublk = sa.u + sa.u_off[slot];
        if (m < m_end) {
            if (s.tid < s.CTm) d_next = dblk[s.tid];
            if (s.tid == 0) u_next = ublk[0];
        }
    } else if (m < m_end) {
        if (s.tid < s.CTm) s.zb[(m & 1) * s.CT + s.tid] = dg_normal(sa.seed, sid, (uint32_t)m, (uint32_t)s.tid);
        if (s.tid == 0) u_next = dg_uniform(sa.seed, sid, (uint32_t)m);
        __syncthreads();
    }

#ifdef DG_PROFILE
    dg_t0 = clock64();
    for (int i = 0; i < 10; ++i) dg_prof[i] = 0;
#endif
    for (; m < m_end; ++m) {
        // trace page of row m (allocated on first touch; the first page always exists)
        if (page == nullptr || (m & pmask) == 0) {
            if (s.tid == 0) {
                long long e = __ldcg(ptab + (m >> pshift));
                if (e < 0) {
                    const unsigned long long need = (unsigned long long)rowlen << pshift;
                    const unsigned long long off = atomicAdd(ct.pool_head, need);
                    if (off + need <= (unsigned long long)ct.pool_cap) {
                        e = (long long)off;
                        ptab[m >> pshift] = e;
                    }
                }
                s.bc[5] = __longlong_as_double(e);
            }
            __syncthreads();
            const long long e = __double_as_longlong(s.bc[5]);
            if (e < 0) {                                  // trace pool exhausted
                state = DICEGP_CHAIN_NOMEM_;
                m_ret = m - 1;
                break;
            }
            page = ct.trace + e;
        }
        // phase 1: proposal Y_try = delta + Y_now (model_GP.py:329; numpy adds the mean last)
        double u_cur = u_next;
        if (s.tid < s.CTm) {
            double d;
            if (sa.mode == 0) {
                d = d_next;
            } else {
                const int c = s.tid / Tc, t = s.tid - c * Tc;
                const double *z = s.zb + (m & 1) * s.CT + c * Tc;
                double a = 0.0;
                for (int j = 0; j < Tc; ++j) a = fma(z[j], s.pfac[j * Tc + t], a);
                d = s.sqs[c] * a;
            }
            s.ytry[s.tid] = d + s.ynow[s.tid];
        }
        // prefetch the next step's stream entries (off the critical path)
        if (m + 1 < m_end) {
            if (sa.mode == 0) {
                if (s.tid < s.CTm) d_next = dblk[(size_t)(m + 1 - m_start) * s.CTm + s.tid];
                if (s.tid == 0) u_next = ublk[m + 1 - m_start];
            } else {
                if (s.tid < s.CTm)
                    s.zb[((m + 1) & 1) * s.CT + s.tid] = dg_normal(sa.seed, sid, (uint32_t)(m + 1), (uint32_t)s.tid);
                if (s.tid == 0) u_next = dg_uniform(sa.seed, sid, (uint32_t)(m + 1));
            }
        }
        __syncthreads();
        DG_TICK(0);
        evaluate(s, true DG_PROF_PASS);
        // MH decision (model_GP.py:348-360), thread 0
        if (s.tid == 0) {
            const double lik = s.bc[2], Q = s.bc[1];
            const double P_try = s.bc[0] + lik;
            const double a = ((P_try + Q) - P_now) - Q;
            const double amin = (0.0 < a) ? 0.0 : a;          // Python min(a, 0): NaN stays NaN
            const double alpha = exp(amin);
            const bool acc = u_cur < alpha;
            if (!isfinite(P_try)) flags |= 1;
            if (acc) { P_now = P_try; L_now = lik; ++cnt; }
            s.bc[4] = acc ? 1.0 : 0.0;
            double *row = page + (size_t)(m & pmask) * rowlen;
            row[2 * s.CT] = P_now;                             // Pro_all[m]  (:362)
            row[2 * s.CT + 1] = L_now;                         // Lik_all[m]  (:363)
        }
        __syncthreads();
        DG_TICK(6);
        if (s.tid < s.CT) {
            if (s.bc[4] != 0.0) {
                s.ynow[s.tid] = s.ytry[s.tid];
                s.psinow[s.tid] = s.psitry[s.tid];
            }
            double *row = page + (size_t)(m & pmask) * rowlen;
            row[s.tid] = s.psinow[s.tid];                      // Psi_all[m]  (:365)
            row[s.CT + s.tid] = s.ynow[s.tid];                 // Y_all[m]    (:364)
        }
        // convergence diagnostics (:369-391); the rows [0, m) exclude the one just written
        if (m >= initial && (m % gap) == 0) {
            __syncthreads();
            DG_TICK(7);
            if (geweke_converged(s, tv, m)) {
                state = DICEGP_CHAIN_CONVERGED_;
                m_ret = m;
                ++m;
                break;
            }
            DG_TICK(8);
        }
        __syncthreads();
        DG_TICK(7);
    }
#ifdef DG_PROFILE
    if (threadIdx.x == 0 && dg_prof_out) {
        for (int i = 0; i < 10; ++i) atomicAdd((unsigned long long *)dg_prof_out + i, (unsigned long long)dg_prof[i]);
        atomicAdd((unsigned long long *)dg_prof_out + 10, (unsigned long long)(m - m_start));
    }
#endif
    if (state == 0 && m >= M) { state = DICEGP_CHAIN_EXHAUSTED_; m_ret = M - 1; }

    // ---- save state ------------------------------------------------------------------------
    __syncthreads();
    if (s.tid < s.CT) { st[s.tid] = s.ynow[s.tid]; st[s.CT + s.tid] = s.psinow[s.tid]; }
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
