// dicegp_tswarp.cuh -- joint (time-series) chains of NARROW genes, one WARP per chain.
//
// A chain with C x Tc <= 32 latent values (2-4 isoforms at 8 time points) gives the general kernel nothing to
// parallelise but the reads: its block of 3-4 warps spends a round mostly waiting -- the state warp for
// the likelihood warps and back (three block barriers), everybody for the dependent fp64 chains of exp /
// divide / log (23 cycles per operation) -- and an SM holds only 4 such blocks: 4.5-5.4 k SM-cycles per MH
// step where the fp64 pipe needs 0.5 k and the L2 0.8 k.  Here ONE WARP is one chain and one block (like the
// single-time-point chains of dicegp_prerun.cuh): lane (c, t) owns a latent value, the serial part of a round
// runs in that warp without any block barrier, an SM holds 10-16 chains whose serial parts and likelihood
// passes interleave on the schedulers, and W (128-256 KB per chain, too much for shared memory at that
// occupancy) streams from L2 through a per-warp ring of cp.async stages: every lane copies exactly the
// 16-byte words it will consume (coalesced 512-byte requests, no barrier, several groups in flight).
// (TMA bulk copies were measured first: issuing one costs the issuing warp ~1000 cycles, which a block of
// one warp cannot hide -- 125 copies of 2 KB per round made the kernel 2 x slower than the general one.)
//
// Round = 4 candidates as an exact rejection chain (candidate d proposes step m+d from the unchanged
// state; the round commits up to the first acceptance, dicegp_wide.cuh explains why that is exact):
//   S     lane (c, t): y_j = delta[m+j] + y_now, e = exp(y), softmax / length reweighting over c by
//         shuffles (numpy's summation order), f[t][k][j] to the warp's scratch; GP prior by shuffles
//   LIK   groups of 64 words (128 reads) of one time point through the cp.async ring; lane l takes words l and
//         l + 32: a 4 reads x 4 candidates register tile, f of the time
//         point in registers; prod x_n as (mantissa product, exponent sum), real-log redo for anything
//         that is not a positive normal number
//   DECIDE lane j: log, P_try = prior + lik, test log u < ((P_try + Q) - P_now) - Q; ballot -> first accept
//   UPDATE state, trace rows, Geweke rule when due (the incremental sums of dicegp_chain.cuh, one warp)
// Same stream, saved state, paged trace and Geweke sums as dg::run_chain: a chain can be advanced by either
// kernel in turn (the straggler pass resumes these chains with the wide-round kernel).
#pragma once
#include "dicegp_wide.cuh"

#define DG_TSW_J 4                  // candidates per round
#define DG_TSW_BATCH 16             // stream steps produced at a time
#define DG_TSW_RING 32              // stream ring rows (two batches)
#define DG_TSW_GW 64                // words (read pairs) of one time point per likelihood group / ring slot
#define DG_TSW_SLOTS 3              // ring slots: two groups in flight while one is consumed

namespace dg {

struct TswLayout { int dring, zbuf, kinv, pfac, ijs, jcon, sqs, len, ff, gw, ints, ring, total; };

// shared memory of one chain (doubles): stream ring + batch scratch, constants, f table, Geweke scratch, then
// the W ring of n_slots groups
__host__ __device__ inline TswLayout dg_tsw_layout(int C, int Tc, int n_slots = DG_TSW_SLOTS) {
    TswLayout L;
    const int CT = C * Tc, RL = CT + 2, ctm = CT - Tc;
    int o = 0;
    L.dring = o; o += DG_TSW_RING * RL;
    L.zbuf = o; o += DG_TSW_BATCH * RL;
    L.kinv = o; o += Tc * Tc;
    L.pfac = o; o += Tc * Tc;
    L.ijs = o; o += C; L.jcon = o; o += C; L.sqs = o; o += C;
    L.len = o; o += CT;
    o = (o + 1) & ~1;
    L.ff = o; o += CT * DG_TSW_J;
    L.gw = o; o += 2 * (ctm < DG_GW_COLS ? (ctm > 0 ? ctm : 1) : DG_GW_COLS);
    L.ints = o; o += (3 * (Tc + 1) + 1) / 2 + 1;
    o = (o + 15) & ~15;
    L.ring = o; o += n_slots * C * DG_TSW_GW * 2;               // [slot][k][64 words]
    L.total = o;
    return L;
}

}  // namespace dg

// One warp (= one block) per chain; persistent blocks pull chains from `queue` (device-stream runs) or block b
// advances chain ids[b] (host-stream chunks), by up to n_steps steps.
#ifndef DG_TSW_MINB
#define DG_TSW_MINB 12               // chains (one-warp blocks) per SM the register budget is set for
#endif
template <int KC>
__global__ void __launch_bounds__(32, DG_TSW_MINB)
dicegp_tswarp_kernel(GeneTab gt, ChainTab ct, StreamArgs sa, const int32_t *__restrict__ ids, int n_ids, int n_steps,
                     int *queue) {
    using namespace dg;
    extern __shared__ __align__(128) double dg_tsw_smem[];
    constexpr int J = DG_TSW_J, C = KC, RING = DG_TSW_RING, GW = DG_TSW_GW, n_slots = DG_TSW_SLOTS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    int slot_id = blockIdx.x;
    for (;;) {
        if (queue != nullptr) {
            int s_ = 0;
            if (lane == 0) s_ = atomicAdd(queue, 1);
            slot_id = __shfl_sync(FULL, s_, 0);
        }
        if (slot_id >= n_ids) break;
        const int chain = ids[slot_id];
        if (ct.state[chain] != 0) { if (queue == nullptr) break; continue; }
        const int g = ct.gene[chain], t0 = ct.t0[chain], Tc = ct.Tc[chain], T = gt.T;
        const int CT = C * Tc, CTm = CT - Tc, RL = CT + 2;
        const TswLayout L = dg_tsw_layout(C, Tc, n_slots);
        double *smem = dg_tsw_smem;
        double *dring = smem + L.dring, *zbuf = smem + L.zbuf, *kinv = smem + L.kinv, *pfac = smem + L.pfac;
        double *ijs = smem + L.ijs, *jcon = smem + L.jcon, *sqs = smem + L.sqs, *lens = smem + L.len, *ff = smem + L.ff;
        double *gw = smem + L.gw;
        int *rel = reinterpret_cast<int *>(smem + L.ints), *ncnt = rel + (Tc + 1), *cum = ncnt + Tc;
        double2 *ring2 = reinterpret_cast<double2 *>(smem + L.ring);
        constexpr int slot_w2 = C * GW;                       // double2 per ring slot

        __syncwarp();                                         // the previous chain is done with the shared memory
        const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t0;
        for (int i = lane; i <= Tc; i += 32) rel[i] = roff[i] - roff[0];
        for (int i = lane; i < Tc; i += 32) ncnt[i] = gt.ncnt[(size_t)g * T + t0 + i];
        {
            const double *pool = ct.pool;
            const double *kv = pool + ct.kinv_off[chain];
            for (int i = lane; i < Tc * Tc; i += 32) kinv[i] = kv[i];
            if (ct.pf_off[chain] >= 0) {
                const double *pf = pool + ct.pf_off[chain];
                for (int i = lane; i < Tc * Tc; i += 32) pfac[i] = pf[i];
            }
            const double *js = pool + ct.js_off[chain], *jc = pool + ct.jc_off[chain];
            if (lane < C - 1) { ijs[lane] = 1.0 / js[lane]; sqs[lane] = sqrt(js[lane]); jcon[lane] = jc[lane]; }
            const double *lg = gt.len + gt.lenoff[g] + (size_t)t0 * C;
            for (int i = lane; i < CT; i += 32) lens[i] = lg[i];                 // len[t][c]
        }
        const double prior_const = ct.prior_const[chain];
        const int M = ct.M[chain], initial = ct.initial[chain], gap = ct.gap[chain];
        int m = ct.m_next[chain];
        double *st = ct.st + ct.st_off[chain];
        const int rowlen = 2 * CT + 2;
        double *gst = st + rowlen;
        if (m == 0 && lane < 4) gst[lane] = 0.0;
        int64_t *ptab = ct.ptab + ct.ptab_off[chain];
        const int pshift = ct.pshift[chain], pmask = (1 << pshift) - 1;
        TraceView tv;
        tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
        const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;
        const bool dev_stream = sa.mode != 0;
        const double *dblk = dev_stream ? nullptr : sa.delta + sa.delta_off[slot_id];
        const double *ublk = dev_stream ? nullptr : sa.u + sa.u_off[slot_id];

        // lane (c, t) owns one latent value
        const bool own = lane < CT;
        const int li = own ? lane : 0;
        const int lc = li / Tc, lt = li - lc * Tc;
        const bool lfree = own && lc < C - 1;
        double y_now = 0.0, psi_now = 0.0, P_now = 0.0, L_now = 0.0;
        int cnt = 0, flags = 0, state_code = 0, m_ret = 0;
        if (m == 0) {
            if (own && ct.y0_off[chain] >= 0) y_now = ct.pool[ct.y0_off[chain] + li];       // Y_now = Ymean (:275)
        } else {
            if (own) { y_now = st[li]; psi_now = st[CT + li]; }
            P_now = st[2 * CT]; L_now = st[2 * CT + 1];
            cnt = ct.cnt[chain]; flags = ct.flags[chain];
        }
        __syncwarp();
        const double len_own = own ? lens[lt * C + lc] : 0.0;

        // ---- the W ring: groups of GW words of one time point, round-periodic; cum[t] = first group of time t ----
        __syncwarp();
        if (lane == 0) {
            int a_ = 0;
            for (int t = 0; t < Tc; ++t) { cum[t] = a_; a_ += (((rel[t + 1] - rel[t]) >> 1) + GW - 1) / GW; }
            cum[Tc] = a_;
        }
        __syncwarp();
        const int NG = cum[Tc];                               // groups per round
        const double2 *Wg2 = reinterpret_cast<const double2 *>(gt.W + gt.woff[g] + (size_t)C * roff[0]);
        // issue cursor (uniform over the warp): group i_g of the i_ngrp groups of time i_t; every lane copies its own
        // two words of each row
        int i_t = -1, i_g = 0, i_ngrp = 0, i_np = 0, i_slot = 0;
        const double2 *i_src = Wg2;
        auto issue_next = [&]() {
            if (NG > 0) {
                if (i_g == i_ngrp) {                          // next time point with reads
                    do { i_t = (i_t + 1 == Tc) ? 0 : i_t + 1; i_ngrp = cum[i_t + 1] - cum[i_t]; } while (i_ngrp == 0);
                    i_g = 0;
                    i_np = (rel[i_t + 1] - rel[i_t]) >> 1;
                    i_src = Wg2 + (size_t)C * (rel[i_t] >> 1) + lane;
                }
                double2 *dst = ring2 + (size_t)i_slot * slot_w2 + lane;
                const int w = i_g * GW + lane;
                const double2 *src = i_src + i_g * GW;
                if (w + 32 < i_np) {                          // both words of this lane exist (all but the last group)
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        cp_async16(dst + k * GW, src + (size_t)k * i_np);
                        cp_async16(dst + k * GW + 32, src + (size_t)k * i_np + 32);
                    }
                } else if (w < i_np) {
#pragma unroll
                    for (int k = 0; k < C; ++k) cp_async16(dst + k * GW, src + (size_t)k * i_np);
                }
                ++i_g;
                i_slot = (i_slot + 1 == n_slots) ? 0 : i_slot + 1;
            }
            cp_async_commit();
        };
        for (int s2 = 0; s2 < n_slots - 1; ++s2) issue_next();
        int c_slot = 0;                                       // slot of the next group to consume
        const int m_start = m;
        const int m_end = (m_start + n_steps < M) ? m_start + n_steps : M;
        int produced = m;                                     // stream ring holds steps [.., produced)
        long long page_cur = -1;
        unsigned long long rounds = 0, w_bytes = 0;
        bool init = (m == 0);
        int next_check = m > initial ? m : initial;
        next_check += (gap - next_check % gap) % gap;

        while (m < m_end || init) {
            // stream rows: a batch whenever the chain comes within J steps of the end of the ring's contents
            while (!init && produced < m_end && produced - m < J) {
                const int upto = min(produced + DG_TSW_BATCH, m_end);
                wide_produce<false>(dev_stream, produced, upto, lane, 32, zbuf, dring, RING, C, Tc, pfac, sqs, kinv, jcon, ijs,
                                    sa.seed, sid, dblk, ublk, m_start, false);
                produced = upto;
            }
            const int nd_max = init ? 0 : min(J, min(m_end - m, next_check - m + 1));
            ++rounds;
            // ================= S: candidates -> f, psi, prior =================
            double yj[J], psij[J], priorj[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const bool prop = !init && j < nd_max;
                double y;
                if (init) y = y_now;
                else if (lfree) y = (prop ? dring[((m + j) % RING) * RL + li] : 0.0) + y_now;   // Y_try = delta + Y_now (:329)
                else y = 0.0;                                                                  // Y_try[C-1,:] stays 0 (:295)
                yj[j] = y;
                const double e = exp(y);
                const double gl = __dmul_rn(len_own, e);
                // sums over the isoforms of this time point, c ascending (numpy: plain loop below 8 elements)
                double es = 0.0, gs = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < C; ++c2) {
                    es += __shfl_sync(FULL, e, c2 * Tc + lt);
                    gs += __shfl_sync(FULL, gl, c2 * Tc + lt);
                }
                psij[j] = e / es;                                                             // psi (:335)
                if (own) ff[(lt * C + lc) * J + j] = gl / gs;                                 // fsi (:336-337), 1/sum(e) cancelled
                // GP prior (:332): x K^-1 x as sum_t (x K^-1)[t] * x[t], then over the isoforms c ascending
                double r = 0.0;
                for (int t2 = 0; t2 < Tc; ++t2) r = fma(__shfl_sync(FULL, y, lc * Tc + t2), kinv[t2 * Tc + lt], r);
                const double pr = __dmul_rn(r, y);
                double a = 0.0;
                for (int t2 = 0; t2 < Tc; ++t2) a += __shfl_sync(FULL, pr, lc * Tc + t2);
                const double pc = prior_const - a / 2.0;                                      // isoform lc (valid in every lane of it)
                double prior = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < C - 1; ++c2) prior += __shfl_sync(FULL, pc, c2 * Tc);
                priorj[j] = prior;
            }
            __syncwarp();                                                                     // f visible to the warp
            // ================= LIK =================
            double lik[J];
            bool redo = false;
            do {
                double mm[J], ls[J];
                int ee[J];
                unsigned bad = 0u;
#pragma unroll
                for (int j = 0; j < J; ++j) { mm[j] = 1.0; ee[j] = 0; ls[j] = 0.0; }
                int it = 0;
                for (int t = 0; t < Tc; ++t) {
                    const int ngrp = cum[t + 1] - cum[t];
                    if (ngrp == 0) continue;
                    const int np = (rel[t + 1] - rel[t]) >> 1, nrd = ncnt[t];
                    double fr[C][J];
#pragma unroll
                    for (int k = 0; k < C; ++k) {
#pragma unroll
                        for (int j = 0; j < J; ++j) fr[k][j] = ff[(t * C + k) * J + j];
                    }
                    for (int gi = 0; gi < ngrp; ++gi, ++it) {
                        issue_next();                                                         // keep n_slots - 1 groups in flight
                        cp_async_wait<n_slots - 1>();
                        const double2 *sb = ring2 + (size_t)c_slot * slot_w2 + lane;
                        c_slot = (c_slot + 1 == n_slots) ? 0 : c_slot + 1;
                        const int wa = gi * GW + lane, wb = wa + 32;                          // words of this lane in the time point
                        const bool va = wa < np, vb = wb < np;
                        // a word past the end recomputes a word of lane 0's (copied whenever the group exists) and is dropped below
                        const double2 *pa = va ? sb : sb - lane, *pb = vb ? sb + 32 : sb - lane;
                        double x[4][J];
#pragma unroll
                        for (int r_ = 0; r_ < 4; ++r_) {
#pragma unroll
                            for (int j = 0; j < J; ++j) x[r_][j] = 0.0;
                        }
#pragma unroll
                        for (int k = 0; k < C; ++k) {
                            const double2 A = pa[k * GW], B = pb[k * GW];
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                x[0][j] = fma(A.x, fr[k][j], x[0][j]);
                                x[1][j] = fma(A.y, fr[k][j], x[1][j]);
                                x[2][j] = fma(B.x, fr[k][j], x[2][j]);
                                x[3][j] = fma(B.y, fr[k][j], x[3][j]);
                            }
                        }
                        const bool full = (gi + 1) * GW <= np && 2 * (gi + 1) * GW <= nrd;     // every read of the group exists (uniform)
                        const bool qa = va && (2 * wa + 1 < nrd), qb = vb && (2 * wb + 1 < nrd);   // second read of the word exists
                        if (redo) {
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                if (va) ls[j] += log(x[0][j]);
                                if (qa) ls[j] += log(x[1][j]);
                                if (vb) ls[j] += log(x[2][j]);
                                if (qb) ls[j] += log(x[3][j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                // scaling by powers of two is exact: mant(product) carries the bits of the product of the
                                // mantissas; a product outside the positive normal range raises `bad`
                                double a0 = x[0][j], a1 = x[1][j], b0 = x[2][j], b1 = x[3][j];
                                if (!full) { a0 = va ? a0 : 1.0; a1 = qa ? a1 : 1.0; b0 = vb ? b0 : 1.0; b1 = qb ? b1 : 1.0; }
                                const int sg = __double2hiint(a0) | __double2hiint(a1) | __double2hiint(b0) | __double2hiint(b1);
                                const double pr = (a0 * a1) * (b0 * b1);
                                const int hh = __double2hiint(pr);
                                bad |= ((unsigned)(hh - 0x00100000) >= 0x7fe00000u) ? 1u : 0u;
                                bad |= (unsigned)sg >> 31;
                                ee[j] += (hh >> 20) - 1023;
                                mm[j] *= dg_mant(pr, hh);
                            }
                            if ((it & 127) == 127) {
#pragma unroll
                                for (int j = 0; j < J; ++j) {
                                    const int gq = __double2hiint(mm[j]);
                                    ee[j] += (gq >> 20) - 1023;
                                    mm[j] = dg_mant(mm[j], gq);
                                }
                            }
                        }
                    }
                    w_bytes += (unsigned long long)np * 16ull * C;
                }
                if (!redo) {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int gq = __double2hiint(mm[j]);
                        ee[j] += (gq >> 20) - 1023;
                        mm[j] = dg_mant(mm[j], gq);
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            mm[j] *= __shfl_xor_sync(FULL, mm[j], off);                       // < 2^32
                            ee[j] += __shfl_xor_sync(FULL, ee[j], off);
                        }
                        bad |= __shfl_xor_sync(FULL, bad, off);
                    }
                    if (bad) { redo = true; continue; }                                       // rare: some x_n was not a positive normal
                    // lane j takes the logarithm of candidate j
                    double mine = mm[0];
                    int eme = ee[0];
#pragma unroll
                    for (int j = 1; j < J; ++j) if (lane == j) { mine = mm[j]; eme = ee[j]; }
                    const int gq = __double2hiint(mine);
                    const double lk = fma((double)(eme + (gq >> 20) - 1023), 0.693147180559945309417232121458, log(dg_mant(mine, gq)));
#pragma unroll
                    for (int j = 0; j < J; ++j) lik[j] = __shfl_sync(FULL, lk, j);
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        double v = ls[j];
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
                        lik[j] = v;
                    }
                }
                break;
            } while (true);
            // ================= decide =================
            int first = -1, nd = 0;
            double P_new = 0.0, L_new = 0.0;
            if (init) {
                L_new = lik[0];
                P_new = priorj[0] + lik[0];                                                   // (:284-292)
                if (!isfinite(P_new)) flags |= 1;
                first = 0;
            } else {
                // trace pages of the rows this round may write (first touch allocates)
                int nd_ok = nd_max;
                long long pg[J];
#pragma unroll
                for (int d = 0; d < J; ++d) {
                    pg[d] = page_cur;
                    if (d < nd_ok) {
                        const int step = m + d;
                        if (d == 0 || (step & pmask) == 0) {
                            long long p_ = 0;
                            if (lane == 0) p_ = trace_page(ptab, step, pshift, rowlen, ct.pool_head, ct.pool_cap);
                            page_cur = __shfl_sync(FULL, p_, 0);
                        }
                        pg[d] = page_cur;
                        if (page_cur < 0) nd_ok = d;
                    }
                }
                double P_try = 0.0, lk = 0.0;
#pragma unroll
                for (int j = 0; j < J; ++j) if (lane == j) { P_try = priorj[j] + lik[j]; lk = lik[j]; }
                bool acc = false;
                if (lane < nd_ok) {
                    const double *dl = dring + ((m + lane) % RING) * RL;
                    const double lu = dl[CT], Q = dl[CT + 1];
                    const double a = ((P_try + Q) - P_now) - Q;                               // (:348)
                    acc = (a >= 0.0) || (lu < a);                                             // u < exp(min(a, 0)); NaN a: both false
                }
                const unsigned am = __ballot_sync(FULL, acc);
                first = am ? (__ffs(am) - 1) : -1;
                nd = first >= 0 ? first + 1 : nd_ok;
                if (__ballot_sync(FULL, lane < nd && !isfinite(P_try))) flags |= 1;
                if (first >= 0) { P_new = __shfl_sync(FULL, P_try, first); L_new = __shfl_sync(FULL, lk, first); }
                // rows m .. m+nd-1: the unchanged state, then (an acceptance ends the round) the new one (:362-366)
                double y_acc = y_now, psi_acc = psi_now;
#pragma unroll
                for (int j = 0; j < J; ++j) if (j == first) { y_acc = yj[j]; psi_acc = psij[j]; }
#pragma unroll
                for (int d = 0; d < J; ++d) {
                    if (d < nd) {
                        const bool nw_ = d == first;
                        double *row = ct.trace + pg[d] + (size_t)((m + d) & pmask) * rowlen;
                        if (own) { row[li] = nw_ ? psi_acc : psi_now; row[CT + li] = nw_ ? y_acc : y_now; }
                        if (lane == 0) { row[2 * CT] = nw_ ? P_new : P_now; row[2 * CT + 1] = nw_ ? L_new : L_now; }
                    }
                }
                if (first >= 0) { y_now = y_acc; psi_now = psi_acc; P_now = P_new; L_now = L_new; ++cnt; }
                m += nd;
                if (nd_ok < nd_max) {                                                         // trace pool exhausted
                    state_code = DICEGP_CHAIN_NOMEM_;
                    m_ret = m - 1;
                    break;
                }
                const int mlast = m - 1;
                if (nd > 0 && mlast >= initial && (mlast % gap) == 0) {                       // convergence diagnostics (:369-391)
                    __syncwarp();
                    if (geweke_converged(gw, gst, CTm, 1, 0, lane, tv, mlast)) {
                        state_code = DICEGP_CHAIN_CONVERGED_;
                        m_ret = mlast;
                        break;
                    }
                    next_check += gap;
                }
                continue;
            }
            // start value evaluated: it becomes the state (:275-292)
            y_now = yj[0]; psi_now = psij[0]; P_now = P_new; L_now = L_new;
            init = false;
        }
        if (state_code == 0 && m >= M) { state_code = DICEGP_CHAIN_EXHAUSTED_; m_ret = M - 1; }

        // ---- drain the ring, save state ----
        cp_async_wait<0>();
        __syncwarp();
        if (own) { st[li] = y_now; st[CT + li] = psi_now; }
        if (lane == 0) {
            st[2 * CT] = P_now; st[2 * CT + 1] = L_now;
            ct.m_next[chain] = m;
            ct.cnt[chain] = cnt;
            ct.flags[chain] = flags;
            ct.m_ret[chain] = m_ret;
            ct.state[chain] = state_code;
            if (ct.ctr) {
                atomicAdd(ct.ctr + 0, w_bytes);
                atomicAdd(ct.ctr + 1, rounds);
                atomicAdd(ct.ctr + 2, rounds * (unsigned long long)J * (unsigned long long)rel[Tc]);
                atomicAdd(ct.ctr + 3, (unsigned long long)(m - m_start));
                atomicAdd(ct.ctr + 4, rounds * (unsigned long long)J * (unsigned long long)rel[Tc] * (unsigned long long)C);
            }
        }
        if (queue == nullptr) break;
    }
}
