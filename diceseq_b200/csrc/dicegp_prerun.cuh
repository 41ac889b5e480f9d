// dicegp_prerun.cuh -- single-time-point chains, one WARP per chain: the pre-runs of get_psi
// (diceseq/utils/run_utils.py:107-114: T chains per gene, M = 100, no convergence check) and
// the chains of static mode (one time point: `diceseq` on a single sample).
//
// The general chain kernel gives such a chain a whole thread block whose warps take turns
// (state warp -> likelihood warps -> decision), and the registers of those mostly idle warps
// cap an SM at 2-8 chains: 80k pre-run chains cost 59 ms of a 730 ms batch.  Here ONE WARP is
// one chain and one block: lane c owns isoform c, the 1000 reads of the time point are split
// over the 32 lanes, there is no block barrier at all, and an SM holds as many chains as their
// W fits shared memory (3 with 8 isoforms ... 14 with 2).  Same arithmetic per step as
// dg::run_chain with Tc = 1 (model_GP.py:312-391): same Philox stream, same softmax / prior /
// accept expressions, likelihood as a (mantissa product, exponent sum) over read pairs with the
// real-log path for anything that is not a positive normal number, the same paged trace, the
// same incremental Geweke rule and the same saved state -- a chain can be advanced by either
// kernel in turn.
#pragma once
#include "dicegp_kernels.cuh"

namespace dg {

#define DG_PR_BATCH 32          // stream steps produced at a time (all lanes busy)

// shared memory of one pre-run chain, in doubles, after its W: normals/increments and Q terms of a
// batch (2 x 32 x ZP), log u and Q per step (2 x 32), per-isoform scratch (7 x CP), Geweke scratch (64)
__host__ __device__ inline int dg_prerun_zp(int C) { return 2 * (C >> 1) > 2 ? 2 * (C >> 1) : 2; }
__host__ __device__ inline int dg_prerun_scratch(int C) {
    return 2 * DG_PR_BATCH * dg_prerun_zp(C) + 2 * DG_PR_BATCH + 7 * ((C + 1) & ~1) + 2 * DG_GW_COLS;
}

struct PrerunEval { double lik, prior; };

// likelihood of the state whose f sits in fv[0..C): sum_n log(sum_k W[n,k] f[k])  (model_GP.py:338-345).
// KC = compile-time isoform count (0: run-time C).  Eight read pairs per lane and iteration: their
// dot products and the four running products are independent dependency chains (a chain's single
// warp has nothing else to hide the shared-memory and the 23-cycle fp64 latency with; registers
// are plentiful at 3-14 warps per SM).
template <int KC>
__device__ __forceinline__ double dg_prerun_lik(const double2 *W2, const double *fv, int C_rt, int npairs, int n, int lane) {
    const int C = KC > 0 ? KC : C_rt;
    constexpr int FR = KC > 0 ? KC : 1;
    constexpr int U = 8;                                  // read pairs per lane and iteration
    double fr[FR];
    if (KC > 0) {
#pragma unroll
        for (int k = 0; k < FR; ++k) fr[k] = fv[k];
    }
    double mm[4] = {1.0, 1.0, 1.0, 1.0};
    int ee = 0, it = 0;
    unsigned bad = 0u;
    const int last = npairs - 1;
    for (int p0 = lane; p0 < npairs; p0 += 32 * U, ++it) {
        double x0[U], x1[U];
        int pc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { pc[u] = min(p0 + 32 * u, last); x0[u] = 0.0; x1[u] = 0.0; }
        if (KC > 0) {
#pragma unroll
            for (int k = 0; k < FR; ++k) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double2 w = W2[(size_t)k * npairs + pc[u]];
                    x0[u] = fma(w.x, fr[k], x0[u]);
                    x1[u] = fma(w.y, fr[k], x1[u]);
                }
            }
        } else {
            for (int k = 0; k < C; ++k) {
                const double f = fv[k];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double2 w = W2[(size_t)k * npairs + pc[u]];
                    x0[u] = fma(w.x, f, x0[u]);
                    x1[u] = fma(w.y, f, x1[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + 32 * u;
            const double a0 = p < npairs ? x0[u] : 1.0;
            const double a1 = (p < npairs && 2 * p + 1 < n) ? x1[u] : 1.0;   // the pad slot of an odd read count
            const int sg = __double2hiint(a0) | __double2hiint(a1);          // two negative x would hide each other
            const double pr2 = a0 * a1;
            const int h = __double2hiint(pr2);
            bad |= ((unsigned)(h - 0x00100000) >= 0x7fe00000u) ? 1u : 0u;
            bad |= (unsigned)sg >> 31;
            ee += (h >> 20) - 1023;
            mm[u & 3] *= __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(pr2));
        }
        if ((it & 31) == 31) {                                               // keep the running products in range
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int g = __double2hiint(mm[u]);
                ee += (g >> 20) - 1023;
                mm[u] = __hiloint2double((g & 0x000fffff) | 0x3ff00000, __double2loint(mm[u]));
            }
        }
    }
    double mt;
    {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int g = __double2hiint(mm[u]);
            ee += (g >> 20) - 1023;
            mm[u] = __hiloint2double((g & 0x000fffff) | 0x3ff00000, __double2loint(mm[u]));
        }
        mt = (mm[0] * mm[1]) * (mm[2] * mm[3]);                               // < 16
        const int g = __double2hiint(mt);
        ee += (g >> 20) - 1023;
        mt = __hiloint2double((g & 0x000fffff) | 0x3ff00000, __double2loint(mt));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mt *= __shfl_xor_sync(0xffffffffu, mt, off);                          // < 2^32
        ee += __shfl_xor_sync(0xffffffffu, ee, off);
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if (!bad) {
        const int g = __double2hiint(mt);
        const double mv = __hiloint2double((g & 0x000fffff) | 0x3ff00000, __double2loint(mt));
        const int ev = ee + (g >> 20) - 1023;
        return fma((double)ev, 0.693147180559945309417232121458, log(mv));
    }
    // some x was zero, negative, denormal or not finite: real logarithms, so that -inf / NaN
    // come out like np.log(...).sum()
    double ls = 0.0;
    for (int p = lane; p < npairs; p += 32) {
        double x0 = 0.0, x1 = 0.0;
        for (int k = 0; k < C; ++k) {
            const double2 w = W2[(size_t)k * npairs + p];
            const double f = fv[k];
            x0 = fma(w.x, f, x0);
            x1 = fma(w.y, f, x1);
        }
        ls += log(x0);
        if (2 * p + 1 < n) ls += log(x1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, off);
    return ls;
}

}  // namespace dg

// One warp (= one block) per chain; persistent blocks pull chains from `queue` and advance them by
// up to n_steps.  w_doubles: doubles reserved for W at the start of the dynamic shared memory.
template <int KC>
__global__ void __launch_bounds__(32)
dicegp_prerun_kernel(GeneTab gt, ChainTab ct, uint64_t seed, const int32_t *__restrict__ ids, int n_ids, int *queue,
                     int w_doubles, int n_steps) {
    extern __shared__ __align__(16) double dg_pr_smem[];
    const int lane = threadIdx.x;
    constexpr unsigned FULL = 0xffffffffu;
    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(queue, 1);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot >= n_ids) break;
        const int chain = ids[slot];
        if (ct.state[chain] != 0) continue;                  // finished earlier
        const int g = ct.gene[chain], t = ct.t0[chain], T = gt.T, C = KC > 0 ? KC : gt.C[g], M = ct.M[chain];
        const int initial = ct.initial[chain], gap = ct.gap[chain];
        const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t;
        const int npairs = (roff[1] - roff[0]) >> 1, n = gt.ncnt[(size_t)g * T + t];
        const double2 *Wg = reinterpret_cast<const double2 *>(gt.W + gt.woff[g] + (size_t)C * roff[0]);
        const int ZP = dg::dg_prerun_zp(C), CP = (C + 1) & ~1;
        double2 *W2 = reinterpret_cast<double2 *>(dg_pr_smem);
        double *zr = dg_pr_smem + w_doubles;                 // [32][ZP] normals, then increments
        double *qt = zr + DG_PR_BATCH * ZP;                  // [32][ZP] proposal log-density terms
        double *lur = qt + DG_PR_BATCH * ZP;                 // [32] log u
        double *Qr = lur + DG_PR_BATCH;                      // [32] Q
        double *ev = Qr + DG_PR_BATCH, *gv = ev + CP, *fv = gv + CP, *pv = fv + CP;
        double *sqs = pv + CP, *ijs = sqs + CP, *jcon = ijs + CP, *gw = jcon + CP;

        __syncwarp();                                        // the previous chain is done with the shared memory
        for (int i = lane; i < C * npairs; i += 32) W2[i] = Wg[i];
        const double *pool = ct.pool;
        const double kinv = pool[ct.kinv_off[chain]], pfac = pool[ct.pf_off[chain]];
        const double prior_const = ct.prior_const[chain];
        if (lane < C - 1) {
            const double js = pool[ct.js_off[chain] + lane];
            ijs[lane] = 1.0 / js;
            sqs[lane] = sqrt(js);
            jcon[lane] = pool[ct.jc_off[chain] + lane];
        }
        const bool own = lane < C, free_ = lane < C - 1;
        const double len_own = own ? gt.len[gt.lenoff[g] + (size_t)t * C + lane] : 0.0;
        const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;
        const int rowlen = 2 * C + 2;
        int64_t *ptab = ct.ptab + ct.ptab_off[chain];
        const int pshift = ct.pshift[chain], pmask = (1 << pshift) - 1;
        TraceView tv;
        tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
        double *st = ct.st + ct.st_off[chain];
        double *gst = st + rowlen;                                        // running Geweke sums
        int m = ct.m_next[chain];
        const int m_first = m;
        const int m_end = (m + n_steps < M) ? m + n_steps : M;
        double y_now = 0.0, psi_now = 0.0, P_now = 0.0, L_now = 0.0;
        int cnt = 0, flags = 0, state = 0, m_ret = 0;
        if (m == 0) {
            if (own && ct.y0_off[chain] >= 0) y_now = pool[ct.y0_off[chain] + lane];   // Y_now = Ymean (:275)
            if (lane < 4) gst[lane] = 0.0;
        } else {                                                          // resume (saved by either kernel)
            if (own) { y_now = st[lane]; psi_now = st[C + lane]; }
            P_now = st[2 * C]; L_now = st[2 * C + 1];
            cnt = ct.cnt[chain]; flags = ct.flags[chain];
        }
        long long page_cur = -1;
        __syncwarp();

        // candidate y (lane c) -> psi (lane c), prior, likelihood            (model_GP.py:332-345)
        auto evaluate = [&](double y, double &psi) -> dg::PrerunEval {
            const double e = exp(y), gl = len_own * e;
            if (own) { ev[lane] = e; gv[lane] = gl; }
            if (free_) pv[lane] = prior_const - ((y * kinv) * y) / 2.0;        // GP prior of one isoform (:332)
            __syncwarp();
            if (own) {
                fv[lane] = gl / dg_np_sum(gv, C, 1);                          // fsi (:336-337), common 1/sum(e) cancelled
                psi = e / dg_np_sum(ev, C, 1);                                // psi (:335)
            }
            double prior = 0.0;
            for (int c = 0; c < C - 1; ++c) prior += pv[c];                   // c ascending (:312-332)
            __syncwarp();
            dg::PrerunEval r;
            r.prior = prior;
            r.lik = dg::dg_prerun_lik<KC>(W2, fv, C, npairs, n, lane);
            __syncwarp();                                                     // ev / gv / fv / pv free again
            return r;
        };

        if (m == 0) {   // start value (:275-292)
            double psi = 0.0;
            const dg::PrerunEval r = evaluate(y_now, psi);
            psi_now = psi;
            L_now = r.lik;
            P_now = r.prior + r.lik;
            if (!isfinite(P_now)) flags |= 1;
        }
        for (int m0 = m; m0 < m_end && state == 0; m0 += DG_PR_BATCH) {
            const int nb = min(DG_PR_BATCH, m_end - m0), npair = C >> 1;
            // ---- the stream of nb steps: normals -> increments, log u, proposal log-density Q ----
            for (int q = lane; q < nb * npair; q += 32) {
                const int s = q / npair, i = q - s * npair;
                double z0, z1;
                dg_normal_pair(seed, sid, (uint32_t)(m0 + s), (uint32_t)i, z0, z1);
                zr[s * ZP + 2 * i] = z0;
                zr[s * ZP + 2 * i + 1] = z1;
            }
            if (lane < nb) lur[lane] = log(dg_uniform(seed, sid, (uint32_t)(m0 + lane)));
            __syncwarp();
            for (int q = lane; q < nb * (C - 1); q += 32) {
                const int s = q / (C - 1), c = q - s * (C - 1);
                const double d = sqs[c] * (zr[s * ZP + c] * pfac);            // delta = sqrt(jump_scale) * (z @ A_K)
                zr[s * ZP + c] = d;
                const double acc = (d * kinv) * d;                            // d K^-1 d
                qt[s * ZP + c] = jcon[c] - (acc * ijs[c]) / 2.0;              // (:330-331)
            }
            __syncwarp();
            if (lane < nb) {
                double Q = 0.0;
                for (int c = 0; c < C - 1; ++c) Q += qt[lane * ZP + c];
                Qr[lane] = Q;
            }
            __syncwarp();
            // ---- the steps ----
            for (int s = 0; s < nb; ++s) {
                // trace page of row m (first touch allocates from the pool)
                if (page_cur < 0 || (m & pmask) == 0) {
                    long long pg = 0;
                    if (lane == 0) pg = dg::trace_page(ptab, m, pshift, rowlen, ct.pool_head, ct.pool_cap);
                    page_cur = __shfl_sync(FULL, pg, 0);
                    if (page_cur < 0) { state = dg::DICEGP_CHAIN_NOMEM_; m_ret = m - 1; break; }
                }
                const double y_try = free_ ? zr[s * ZP + lane] + y_now : 0.0; // (:329; Y_try[C-1] stays 0)
                double psi_try = 0.0;
                const dg::PrerunEval r = evaluate(y_try, psi_try);
                const double P_try = r.prior + r.lik;
                const double lu = lur[s], Q = Qr[s];
                const double a = ((P_try + Q) - P_now) - Q;                   // (:348)
                const bool acc = (a >= 0.0) || (lu < a);                      // u < exp(min(a, 0)); NaN a: both false
                if (!isfinite(P_try)) flags |= 1;
                if (acc) { ++cnt; P_now = P_try; L_now = r.lik; y_now = y_try; psi_now = psi_try; }
                double *row = ct.trace + page_cur + (size_t)(m & pmask) * rowlen;   // (:362-366)
                if (own) { row[lane] = psi_now; row[C + lane] = y_now; }
                if (lane == 0) { row[2 * C] = P_now; row[2 * C + 1] = L_now; }
                const int mlast = m;
                ++m;
                // convergence diagnostics (:369-391) over rows [0, mlast)
                if (mlast >= initial && (mlast % gap) == 0) {
                    __syncwarp();                                             // the rows are visible to the warp
                    if (dg::geweke_converged(gw, gst, C - 1, 1, 0, lane, tv, mlast)) {
                        state = dg::DICEGP_CHAIN_CONVERGED_;
                        m_ret = mlast;
                        break;
                    }
                }
            }
        }
        if (state == 0 && m >= M) { state = dg::DICEGP_CHAIN_EXHAUSTED_; m_ret = M - 1; }
        // ---- save state, as dg::run_chain does ----
        __syncwarp();
        if (own) { st[lane] = y_now; st[C + lane] = psi_now; }
        if (lane == 0) {
            st[2 * C] = P_now; st[2 * C + 1] = L_now;
            ct.m_next[chain] = m;
            ct.cnt[chain] = cnt;
            ct.flags[chain] = flags;
            ct.m_ret[chain] = m_ret;
            ct.state[chain] = state;
            if (ct.ctr) {
                const unsigned long long ev = (unsigned long long)(m - m_first) + (m_first == 0 ? 1ull : 0ull);
                atomicAdd(ct.ctr + 0, (unsigned long long)C * npairs * 16ull);
                atomicAdd(ct.ctr + 1, ev);
                atomicAdd(ct.ctr + 2, ev * 2ull * (unsigned long long)npairs);
                atomicAdd(ct.ctr + 3, (unsigned long long)(m - m_first));
                atomicAdd(ct.ctr + 4, ev * 2ull * (unsigned long long)npairs * (unsigned long long)C);
            }
        }
    }
}
