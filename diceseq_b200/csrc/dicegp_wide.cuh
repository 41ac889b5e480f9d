// dicegp_wide.cuh -- the MH loop of a joint (time-series) chain, kernel v3: WIDE rejection-chain rounds
// without warp specialisation, W streamed through a block-shared TMA ring.
//
// What bounds a chain is the serial part of a round (proposal -> exp -> softmax -> f | likelihood |
// log -> accept test -> state update: ~3000 cycles of dependent fp64 latency) and, for wide genes, the
// bytes of W a round pulls from L2 / HBM.  Both are paid per ROUND, not per step, so a round here
// evaluates J = 8 or 16 candidates at once: candidate d is the proposal of step m+d made from the
// UNCHANGED state (model_GP.py:329 draws Y_try from Y_now; the increments and uniforms come from a
// stream that does not depend on the chain, :329,:351).  The round commits steps m .. m+d* where d* is
// the first accepted candidate (every earlier step was a rejection, so its proposal really was made from
// that state): (1 - (1-p)^J) / p steps per pass over W at acceptance rate p -- 5.9 at the time-series
// average p = 0.16, 12 for the slow chains at p = 0.04.  Arithmetic per committed step is that of
// sequential execution.
//
// One block = NCW consumer warps + 1 auxiliary warp; every consumer warp does every phase:
//   S      item (j, t): candidate y -> exp -> length-weighted softmax f[t][.][j];  item (j, c): GP prior
//   LIK    reads in TASKS of 16 words (32 reads) of one time point; a STAGE = NCW consecutive tasks, one
//          per consumer warp, brought in by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx,
//          one copy per isoform row and time point) into a ring of D stages.  The warp that releases a
//          stage last re-arms its barrier and issues the copies of stage q + D (no producer warp, no
//          empty barriers).  The stage sequence of a round is the same every round, so the first stages
//          of the next round are prefetched while this round decides; when the whole W fits the ring
//          (D >= stages per round) nothing is ever re-issued: W is resident.
//          Lane (s, h), s = lane & 7, h = lane >> 3: words s and s + 8 of the task (4 reads), candidates
//          4h .. 4h+3 (RJ = 4) -- a 4 x RJ register tile per lane; per isoform 2 conflict-free LDS.128 of
//          W (8 lanes x 16 B, broadcast over h) and RJ/2 LDS.128 of f feed 4 RJ DFMA: 0.25 shared-memory
//          wavefronts per DFMA, half of what the fp64 pipe can consume.
//          prod_n x_n per candidate as (mantissa product, exponent sum); anything that is not a positive
//          normal number sends the round to the real-log path (-inf / NaN like np.log(...).sum()).
//   AUX    meanwhile the auxiliary warp produces the stream of the NEXT rounds (Philox normals ->
//          increments, log u, proposal log-density Q) and reserves the trace pages of this round's rows.
//   DECIDE every warp redundantly: lane j combines the per-warp partial products of candidate j, takes one
//          log, adds the prior, tests  log u < ((P_try + Q) - P_now) - Q ; a ballot gives d*.
//   UPDATE psi of the accepted state, trace rows (coalesced, all threads), Geweke rule when due.
// Four block barriers per round.  Same stream, same saved state, same paged trace and incremental Geweke
// sums as dg::run_chain -- a chain can be advanced by either kernel in turn.
#pragma once
#include "dicegp_chain.cuh"

#define DG_WIDE_MAXSTAGES 64
#define DG_WIDE_MAXCT 128           // latent values (C * Tc) of a chain this kernel takes
#define DG_WIDE_THREADS 576         // launch bound: 17 warps (16 consumers + aux) or 2 blocks of 9 warps per SM

#ifdef DG_PROFILE
// warp 0 lane 0: 0 S, 1 wait f, 2 likelihood, 3 wait partials, 4 decide + new state, 5 wait state, 6 rows; aux lane 0: 7 busy
#define DG_WTICK(i) do { if (lane == 0 && warp == 0) { long long t__ = clock64(); wprof[i] += t__ - wt0; wt0 = t__; } } while (0)
#else
#define DG_WTICK(i) do { } while (0)
#endif

namespace dg {

struct WideLayout {
    int state, ycand, ex, pr, ff, priorc, redm, rede, redb, bc, zbuf, dring, kinv, pfac, ijs, jcon, sqs, len, ints, gw, bars;
    int ring;                   // doubles before the ring (128-byte aligned)
};

__host__ __device__ inline WideLayout dg_wide_layout(int C, int Tc, int ncw, int J) {
    WideLayout L;
    const int CT = C * Tc, RL = CT + 2, ctm = CT - Tc;
    int o = 0;
    L.state = o; o += 4 * CT;                          // two buffers of [psi | y]
    L.ycand = o; o += J * CT;
    L.ex = o; o += J * (CT | 1);                       // rows padded to an odd length (bank spread over candidates)
    L.pr = o; o += J * CT;
    o = (o + 1) & ~1;
    L.ff = o; o += CT * J;                             // f[t][k][j], rows of J doubles (16-byte aligned)
    L.priorc = o; o += J * C;
    L.redm = o; o += ncw * J;
    L.rede = o; o += (ncw * J + 1) / 2;                // ints
    L.redb = o; o += (ncw + 1) / 2;                    // ints
    L.bc = o; o += 96;
    L.zbuf = o; o += 2 * J * RL;
    L.dring = o; o += 2 * J * RL;
    L.kinv = o; o += Tc * Tc;
    L.pfac = o; o += Tc * Tc;
    L.ijs = o; o += C; L.jcon = o; o += C; L.sqs = o; o += C;
    L.len = o; o += CT;
    L.ints = o; o += (3 * (Tc + 1) + 1) / 2 + 1;       // rel[Tc+1], ncnt[Tc], cum[Tc+1]
    L.gw = o; o += 2 * (ncw + 1) * (ctm < DG_GW_COLS ? (ctm > 0 ? ctm : 1) : DG_GW_COLS);
    L.bars = o; o += DG_WIDE_MAXSTAGES + DG_WIDE_MAXSTAGES / 2;   // full barriers (8 B) + release counters (4 B)
    o = (o + 15) & ~15;
    L.ring = o;
    return L;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

// numpy's short pairwise sum (dg_np_sum) over values formed on the fly
template <typename F>
__device__ __forceinline__ double dg_np_sum_fn(F v, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += v(i);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = v(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += v(i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += v(i);
    return res;
}

// Stream rows of steps [s0, s1), s1 - s0 <= 2J, into the ring (row = step % RING): the arithmetic of
// dg::produce_stream (same Philox counters, Box-Muller pairing, z @ A_K accumulated j ascending with fma,
// Q from -delta).  Called by `nprod` threads (index pt) that synchronise through SYNC().
template <bool BLOCK>
__device__ __forceinline__ void wide_sync() { if (BLOCK) __syncthreads(); else __syncwarp(); }

template <bool BLOCK>
__device__ __noinline__ void wide_produce(bool dev_stream, int s0, int s1, int pt, int nprod, double *zbuf, double *dring,
                                          int RING, int C, int Tc, const double *pfac, const double *sqs, const double *kinv,
                                          const double *jcon, const double *ijs, uint64_t seed, int64_t sid,
                                          const double *dblk, const double *ublk, int m_start) {
    const int CT = C * Tc, CTm = CT - Tc, RL = CT + 2, ns = s1 - s0;
    if (ns <= 0) return;
    if (dev_stream) {
        const int npair = (CTm + 1) >> 1;
        for (int q = pt; q < ns * npair; q += nprod) {
            const int ds = q / npair, i = q - ds * npair;
            double z0, z1;
            dg_normal_pair(seed, sid, (uint32_t)(s0 + ds), (uint32_t)i, z0, z1);
            zbuf[ds * RL + 2 * i] = z0;
            if (2 * i + 1 < CTm) zbuf[ds * RL + 2 * i + 1] = z1;
        }
        for (int ds = pt; ds < ns; ds += nprod)
            dring[((s0 + ds) % RING) * RL + CT] = dg_log(dg_uniform(seed, sid, (uint32_t)(s0 + ds)));
        wide_sync<BLOCK>();
        for (int q = pt; q < ns * CTm; q += nprod) {
            const int ds = q / CTm, i = q - ds * CTm;
            const int c = i / Tc, t = i - c * Tc;
            const double *z = zbuf + ds * RL + c * Tc;
            double a = 0.0;
            for (int j = 0; j < Tc; ++j) a = fma(z[j], pfac[j * Tc + t], a);
            dring[((s0 + ds) % RING) * RL + i] = sqs[c] * a;
        }
    } else {
        for (int q = pt; q < ns * CTm; q += nprod) {
            const int ds = q / CTm, i = q - ds * CTm;
            dring[((s0 + ds) % RING) * RL + i] = dblk[(size_t)(s0 + ds - m_start) * CTm + i];
        }
        for (int ds = pt; ds < ns; ds += nprod)
            dring[((s0 + ds) % RING) * RL + CT] = dg_log(ublk[s0 + ds - m_start]);
    }
    wide_sync<BLOCK>();
    for (int q = pt; q < ns * CTm; q += nprod) {                   // per (step, c, t): (d K^-1)[t]
        const int ds = q / CTm, i = q - ds * CTm;
        const int c = i / Tc, t = i - c * Tc;
        const double *d = dring + ((s0 + ds) % RING) * RL + c * Tc;
        double r = 0.0;
        for (int j = 0; j < Tc; ++j) r = fma(d[j], kinv[j * Tc + t], r);
        zbuf[ds * RL + i] = r;                                     // the normals are consumed
    }
    wide_sync<BLOCK>();
    for (int q = pt; q < ns * (C - 1); q += nprod) {               // per (step, c): d K^-1 d -> proposal log-density term
        const int ds = q / (C - 1), c = q - ds * (C - 1);
        const double *d = dring + ((s0 + ds) % RING) * RL + c * Tc;
        double *rr = zbuf + ds * RL + c * Tc;
        double acc = 0.0;
        for (int t = 0; t < Tc; ++t) acc = fma(rr[t], d[t], acc);
        rr[0] = jcon[c] - (acc * ijs[c]) / 2.0;
    }
    wide_sync<BLOCK>();
    for (int ds = pt; ds < ns; ds += nprod) {
        double Q = 0.0;
        for (int c = 0; c < C - 1; ++c) Q += zbuf[ds * RL + c * Tc];
        dring[((s0 + ds) % RING) * RL + CT + 1] = Q;
    }
    wide_sync<BLOCK>();
}

// One task: 16 words (32 reads) of one time point, RJ candidates per lane.
//   wt       the task's words of isoform 0 in the stage (double2), isoform k at wt + k * kstride
//   ffrow    f[t][0][0]: f[t][k][j] at ffrow + k * J + j
//   nw       valid words of the task (1..16); padw = the word whose second read is the zero pad of an odd
//            read count, or -1
template <int KC, int RJ, bool LOGPATH>
__device__ __forceinline__ void wide_task(const double2 *__restrict__ wt, int kstride, int C, const double *__restrict__ ffrow,
                                          int lane, int nw, int padw, double (&mm)[RJ], int (&ee)[RJ], unsigned &bad,
                                          double (&ls)[RJ]) {
    constexpr int J = 4 * RJ;
    const int s8 = lane & 7, h = lane >> 3;
    const bool va = s8 < nw, vb = s8 + 8 < nw;
    const int wa = va ? s8 : 0, wb = vb ? s8 + 8 : 0;       // a word past the end recomputes word 0 and is dropped below
    double x[4][RJ];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int n = 0; n < RJ; ++n) x[r][n] = 0.0;
    }
    const double2 *fp = reinterpret_cast<const double2 *>(ffrow + h * RJ);
    const int kmax = KC > 0 ? KC : C;
#pragma unroll(KC > 0 ? KC : 1)
    for (int k = 0; k < kmax; ++k) {
        const double2 A = wt[k * kstride + wa], B = wt[k * kstride + wb];
        double f[RJ];
#pragma unroll
        for (int u = 0; u < RJ / 2; ++u) { const double2 t = fp[(k * J) / 2 + u]; f[2 * u] = t.x; f[2 * u + 1] = t.y; }
#pragma unroll
        for (int n = 0; n < RJ; ++n) {
            x[0][n] = fma(A.x, f[n], x[0][n]);
            x[1][n] = fma(A.y, f[n], x[1][n]);
            x[2][n] = fma(B.x, f[n], x[2][n]);
            x[3][n] = fma(B.y, f[n], x[3][n]);
        }
    }
    const bool pa = va && (s8 != padw), pb = vb && (s8 + 8 != padw);   // the second read of the word exists
    if (LOGPATH) {
#pragma unroll
        for (int n = 0; n < RJ; ++n) {
            if (va) ls[n] += log(x[0][n]);
            if (pa) ls[n] += log(x[1][n]);
            if (vb) ls[n] += log(x[2][n]);
            if (pb) ls[n] += log(x[3][n]);
        }
    } else {
#pragma unroll
        for (int n = 0; n < RJ; ++n) {
            // scaling by powers of two is exact, so mant(product) carries the bits of the product of the
            // mantissas; a product outside the positive normal range (a zero, negative or non-finite x, an
            // under/overflow) raises `bad`
            const double a0 = va ? x[0][n] : 1.0, a1 = pa ? x[1][n] : 1.0;
            const double b0 = vb ? x[2][n] : 1.0, b1 = pb ? x[3][n] : 1.0;
            const int sg = __double2hiint(a0) | __double2hiint(a1) | __double2hiint(b0) | __double2hiint(b1);   // two negative x would hide each other
            const double pr = (a0 * a1) * (b0 * b1);
            const int hh = __double2hiint(pr);
            bad |= ((unsigned)(hh - 0x00100000) >= 0x7fe00000u) ? 1u : 0u;
            bad |= (unsigned)sg >> 31;
            ee[n] += (hh >> 20) - 1023;
            mm[n] *= dg_mant(pr, hh);
        }
    }
}

// Per-candidate partials of a warp -> red[cw][j].  The 8 lanes s = lane & 7 of a group h hold partials of
// the same RJ candidates; a split butterfly (each stage halves the candidates a lane carries) leaves one
// candidate per lane in 3 stages.  PROD: (mantissa, exponent) products, else sums of logs.
template <int RJ, bool PROD>
__device__ __forceinline__ void wide_flush(double (&mm)[RJ], int (&ee)[RJ], int lane, double *redm, int *rede, int cw) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int J = 4 * RJ;
    const bool b0 = lane & 1, b1 = lane & 2;
    double km; int ke = 0, node;
    if (RJ == 4) {
        double s0 = b0 ? mm[0] : mm[2], s1 = b0 ? mm[1] : mm[3];
        double k0 = b0 ? mm[2] : mm[0], k1 = b0 ? mm[3] : mm[1];
        int se0 = b0 ? ee[0] : ee[2], se1 = b0 ? ee[1] : ee[3];
        int ke0 = b0 ? ee[2] : ee[0], ke1 = b0 ? ee[3] : ee[1];
        s0 = __shfl_xor_sync(FULL, s0, 1); s1 = __shfl_xor_sync(FULL, s1, 1);
        if (PROD) { k0 *= s0; k1 *= s1; ke0 += __shfl_xor_sync(FULL, se0, 1); ke1 += __shfl_xor_sync(FULL, se1, 1); }
        else { k0 += s0; k1 += s1; }
        double s2 = b1 ? k0 : k1;
        km = b1 ? k1 : k0;
        int se2 = b1 ? ke0 : ke1;
        ke = b1 ? ke1 : ke0;
        s2 = __shfl_xor_sync(FULL, s2, 2);
        if (PROD) { km *= s2; ke += __shfl_xor_sync(FULL, se2, 2); } else km += s2;
        node = (b0 ? 2 : 0) + (b1 ? 1 : 0);
    } else {
        double s0 = b0 ? mm[0] : mm[1];
        km = b0 ? mm[1] : mm[0];
        int se0 = b0 ? ee[0] : ee[1];
        ke = b0 ? ee[1] : ee[0];
        s0 = __shfl_xor_sync(FULL, s0, 1);
        if (PROD) { km *= s0; ke += __shfl_xor_sync(FULL, se0, 1); } else km += s0;
        const double s2 = __shfl_xor_sync(FULL, km, 2);
        if (PROD) { km *= s2; ke += __shfl_xor_sync(FULL, ke, 2); } else km += s2;
        node = b0 ? 1 : 0;
    }
    {
        const double s4 = __shfl_xor_sync(FULL, km, 4);
        if (PROD) { km *= s4; ke += __shfl_xor_sync(FULL, ke, 4); } else km += s4;
    }
    const bool writer = RJ == 4 ? ((lane & 4) == 0) : ((lane & 6) == 0);
    if (writer) {
        const int j = (lane >> 3) * RJ + node;
        if (PROD) {
            const int g = __double2hiint(km);
            redm[cw * J + j] = dg_mant(km, g);
            rede[cw * J + j] = ke + (g >> 20) - 1023;
        } else {
            redm[cw * J + j] = km;
        }
    }
}

// ---------------------------------------------------------------------------------------
// one chain, up to n_steps steps
// ---------------------------------------------------------------------------------------
template <int KC, int RJ>
__device__ void run_chain_wide(const GeneTab &gt, const ChainTab &ct, const StreamArgs &sa, int chain, int slot_id,
                               int n_steps, double *smem, int ring_bytes) {
    constexpr int J = 4 * RJ, RING = 2 * J;
    constexpr unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int NCW = nwarps - 1;
    const bool is_aux = warp == NCW;

    const int g = ct.gene[chain], t0 = ct.t0[chain];
    const int C = KC > 0 ? KC : gt.C[g];
    const int Tc = ct.Tc[chain], T = gt.T;
    const int CT = C * Tc, CTm = (C - 1) * Tc, RL = CT + 2, CTP = CT | 1;
    const WideLayout L = dg_wide_layout(C, Tc, NCW, J);
    double *state = smem + L.state, *ycand = smem + L.ycand, *ex = smem + L.ex, *pr = smem + L.pr, *ff = smem + L.ff;
    double *priorc = smem + L.priorc, *redm = smem + L.redm, *bc = smem + L.bc, *zbuf = smem + L.zbuf, *dring = smem + L.dring;
    double *kinv = smem + L.kinv, *pfac = smem + L.pfac, *ijs = smem + L.ijs, *jcon = smem + L.jcon, *sqs = smem + L.sqs;
    double *lens = smem + L.len, *gw = smem + L.gw;
    int *rede = reinterpret_cast<int *>(smem + L.rede), *redb = reinterpret_cast<int *>(smem + L.redb);
    int *rel = reinterpret_cast<int *>(smem + L.ints), *ncnt = rel + (Tc + 1), *cum = ncnt + Tc;
    unsigned long long *fullb = reinterpret_cast<unsigned long long *>(smem + L.bars);
    int *relcnt = reinterpret_cast<int *>(fullb + DG_WIDE_MAXSTAGES);
    double2 *ring2 = reinterpret_cast<double2 *>(smem + L.ring);
    const double prior_const = ct.prior_const[chain];

    const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t0;
    const int r_first = roff[0];
    const double *Wg = gt.W + gt.woff[g] + (size_t)C * r_first;

    // ---- per-chain constants ----
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
        const double *lg = gt.len + gt.lenoff[g] + (size_t)t0 * C;
        for (int i = tid; i < CT; i += nthreads) lens[i] = lg[i];             // len[t][c]
    }
    const int M = ct.M[chain], initial = ct.initial[chain], gap = ct.gap[chain];
    int m = ct.m_next[chain];
    double *st = ct.st + ct.st_off[chain];
    const int rowlen = 2 * CT + 2;
    double *gst = st + rowlen;                            // running Geweke sums (geweke_converged)
    if (m == 0 && tid < 4) gst[tid] = 0.0;
    int64_t *ptab = ct.ptab + ct.ptab_off[chain];
    const int pshift = ct.pshift[chain];
    const int pmask = (1 << pshift) - 1;
    TraceView tv;
    tv.pool = ct.trace; tv.tab = ptab; tv.shift = pshift; tv.rowlen = rowlen;
    const int64_t sid = ct.stream_id ? ct.stream_id[chain] : (int64_t)chain;
    const bool dev_stream = sa.mode != 0;
    const double *y0p = ct.y0_off[chain] >= 0 ? ct.pool + ct.y0_off[chain] : nullptr;

    // state buffers: [psi (CT) | y (CT)], the layout of a trace row
    int cur = 0;
    double P_now = 0.0, L_now = 0.0;
    int cnt = 0, flags = 0;
    if (m == 0) {
        for (int i = tid; i < CT; i += nthreads) { state[i] = 0.0; state[CT + i] = y0p ? y0p[i] : 0.0; }   // Y_now = Ymean (:275)
    } else {
        for (int i = tid; i < CT; i += nthreads) { state[CT + i] = st[i]; state[i] = st[CT + i]; }
        P_now = st[2 * CT]; L_now = st[2 * CT + 1];
        cnt = ct.cnt[chain]; flags = ct.flags[chain];
    }
    for (int i = tid; i < NCW * J; i += nthreads) { redm[i] = 1.0; rede[i] = 0; }
    if (tid < NCW) redb[tid] = 0;
    __syncthreads();                                      // rel / ncnt visible
    if (tid == 0) {
        int a = 0;
        for (int t = 0; t < Tc; ++t) { cum[t] = a; a += (((rel[t + 1] - rel[t]) >> 1) + 15) >> 4; }
        cum[Tc] = a;
    }
    __syncthreads();
    const int NT = cum[Tc];                               // tasks per round
    const int NS = (NT + NCW - 1) / NCW;                  // stages per round
    const int stage_w2 = NCW * C * 16;                    // double2 per stage
    int D = ring_bytes / (stage_w2 * 16);
    if (D > DG_WIDE_MAXSTAGES) D = DG_WIDE_MAXSTAGES;
    if (D > NS) D = NS;
    const bool resident = NS <= D;                        // the whole W sits in the ring: loaded once
    if (tid == 0) {
        for (int s = 0; s < D; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dg_smem_u32(fullb + s)));
            relcnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // copies of stage number q (round-periodic) into its ring slot, by the 32 lanes of one warp
    unsigned long long tma_bytes = 0;
    auto issue_stage = [&](int q) {
        const int s = q % NS, rs = q % D;
        const int task0 = s * NCW, task1 = min(NT, task0 + NCW);
        int tfirst = 0;
        while (cum[tfirst + 1] <= task0) ++tfirst;
        uint32_t words = 0;
        for (int a = task0, tt = tfirst; a < task1; ++tt) {
            const int b = min(task1, cum[tt + 1]);
            const int np = (rel[tt + 1] - rel[tt]) >> 1;
            const int w0 = (a - cum[tt]) * 16, w1 = min((b - cum[tt]) * 16, np);
            if (w1 > w0) words += (uint32_t)(w1 - w0);
            a = b;
        }
        const uint32_t bar = dg_smem_u32(fullb + rs);
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(words * 16u * (uint32_t)C) : "memory");
            tma_bytes += (unsigned long long)words * 16u * (unsigned)C;
        }
        __syncwarp();
        for (int a = task0, tt = tfirst; a < task1; ++tt) {
            const int b = min(task1, cum[tt + 1]);
            const int np = (rel[tt + 1] - rel[tt]) >> 1;
            const int w0 = (a - cum[tt]) * 16, w1 = min((b - cum[tt]) * 16, np);
            if (w1 > w0) {
                for (int k = lane; k < C; k += 32) {
                    const double2 *src = reinterpret_cast<const double2 *>(Wg + (size_t)C * rel[tt]) + (size_t)k * np + w0;
                    double2 *dst = ring2 + (size_t)rs * stage_w2 + (k * NCW + (a - task0)) * 16;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(dg_smem_u32(dst)), "l"(src), "r"((uint32_t)(w1 - w0) * 16u), "r"(bar) : "memory");
                }
            }
            a = b;
        }
    };

    const int m_start = m;
    const int m_end = (m_start + n_steps < M) ? (m_start + n_steps) : M;
    int state_code = 0, m_ret = 0;
    long long page_cur = -1;                              // aux lane 0
    const double *dblk = dev_stream ? nullptr : sa.delta + sa.delta_off[slot_id];
    const double *ublk = dev_stream ? nullptr : sa.u + sa.u_off[slot_id];
    if (is_aux) {
        for (int q = 0; q < D; ++q) issue_stage(q);       // prologue: the first D stages
    }
    int produced = min(m + RING, m_end);                  // ring holds steps [.., produced)
    wide_produce<true>(dev_stream, m, produced, tid, nthreads, zbuf, dring, RING, C, Tc, pfac, sqs, kinv, jcon, ijs,
                       sa.seed, sid, dblk, ublk, m_start);

    int qs = 0;                                           // stages consumed so far (uniform over the block)
    unsigned long long rounds = 0;
    bool init = (m == 0);                                 // evaluate the start value first (model_GP.py:275-292)
    int next_check = m > initial ? m : initial;
    next_check += (gap - next_check % gap) % gap;

#ifdef DG_PROFILE
    long long wprof[10];
    for (int i = 0; i < 10; ++i) wprof[i] = 0;
    long long wt0 = clock64();
#endif
    while (m < m_end || init) {
        const int nd_max = init ? 0 : min(J, min(m_end - m, next_check - m + 1));
        const double *snow = state + cur * 2 * CT;        // [psi | y]
        const double *ynow = snow + CT;
        // ================= S phase (two sub-phases, every thread) =================
        // S1, item (j, c, t): candidate y -> exp; (x K^-1)[t] * x[t] of the GP prior (:332).  One latent value per
        // item: the exp / fma chains of a thread's items are independent (fp64 latency, not throughput, is the cost)
        for (int v = tid; v < J * CT; v += nthreads) {
            const int j = v / CT, i = v - j * CT;
            const int c = i / Tc, t = i - c * Tc;
            const bool prop = !init && j < nd_max;
            const double *dl = dring + ((m + j) % RING) * RL;
            double y;
            if (init) y = ynow[i];
            else if (c < C - 1) y = (prop ? dl[i] : 0.0) + ynow[i];           // Y_try = delta + Y_now (:329)
            else y = 0.0;                                                      // Y_try[C-1,:] stays 0 (:295)
            ycand[v] = y;
            if (c < C - 1) {
                const double *dc = dl + c * Tc, *yn = ynow + c * Tc;
                double r = 0.0;
                for (int t2 = 0; t2 < Tc; ++t2) {
                    const double y2 = init ? yn[t2] : ((prop ? dc[t2] : 0.0) + yn[t2]);
                    r = fma(y2, kinv[t2 * Tc + t], r);
                }
                pr[v] = __dmul_rn(r, y);
            }
            ex[j * CTP + i] = exp(y);
        }
        __syncthreads();
        // S2, item (j, c, t): f = len e / sum(len e) (:335-337; the 1/sum(e) of psi cancels);  item (j, c): prior term
        for (int v = tid; v < J * CT + J * (C - 1); v += nthreads) {
            if (v < J * CT) {
                const int idx = v / J, j = v - idx * J;                        // candidates fastest: conflict-free stores of f
                const int t = idx / C, c = idx - t * C;
                const double *lt = lens + t * C, *et = ex + j * CTP + t;
                const double gsum = dg_np_sum_fn([&](int c2) { return __dmul_rn(lt[c2], et[c2 * Tc]); }, C);
                ff[idx * J + j] = __dmul_rn(lt[c], et[c * Tc]) / gsum;
            } else {
                const int q = v - J * CT;
                const int j = q / (C - 1), c = q - j * (C - 1);
                const double *pc = pr + j * CT + c * Tc;
                double a = 0.0;
                for (int t = 0; t < Tc; ++t) a += pc[t];
                priorc[j * C + c] = prior_const - a / 2.0;
            }
        }
        DG_WTICK(0);
        __syncthreads();                                                     // f, prior terms ready
        DG_WTICK(1);
        // ================= likelihood (consumer warps) || stream + trace pages (aux warp) =================
        bool redo = false;
        do {
            if (!is_aux) {
                double mm[RJ], ls[RJ];
                int ee[RJ];
                unsigned bad = 0u;
#pragma unroll
                for (int n = 0; n < RJ; ++n) { mm[n] = 1.0; ee[n] = 0; ls[n] = 0.0; }
                int tt = 0, it = 0;
                int rslot = D > 0 ? qs % D : 0;                                  // ring slot / barrier parity of stage qs
                uint32_t rphase = D > 0 ? (uint32_t)((qs / D) & 1) : 0u;
                for (int s = 0; s < NS; ++s) {
                    mbar_wait(dg_smem_u32(fullb + rslot), resident ? 0u : rphase);
                    const int task = s * NCW + warp;
                    if (task < NT) {
                        while (task >= cum[tt + 1]) ++tt;
                        const int np = (rel[tt + 1] - rel[tt]) >> 1;
                        const int w0 = (task - cum[tt]) * 16;
                        const int nw = min(16, np - w0);
                        const int padw = ((ncnt[tt] & 1) && (w0 + nw == np)) ? nw - 1 : -1;
                        const double2 *wt = ring2 + (size_t)rslot * stage_w2 + warp * 16;
                        if (redo) wide_task<KC, RJ, true>(wt, NCW * 16, C, ff + tt * C * J, lane, nw, padw, mm, ee, bad, ls);
                        else wide_task<KC, RJ, false>(wt, NCW * 16, C, ff + tt * C * J, lane, nw, padw, mm, ee, bad, ls);
                        if (!redo && (++it & 127) == 0) {                        // keep the running products in range
#pragma unroll
                            for (int n = 0; n < RJ; ++n) {
                                const int gq = __double2hiint(mm[n]);
                                ee[n] += (gq >> 20) - 1023;
                                mm[n] = dg_mant(mm[n], gq);
                            }
                        }
                    }
                    if (!resident) {
                        // release the slot; the warp that arrives last re-arms it with stage qs + D
                        __syncwarp();
                        int old = 0;
                        if (lane == 0)
                            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(dg_smem_u32(relcnt + rslot)) : "memory");
                        old = __shfl_sync(FULL, old, 0);
                        if (old == NCW - 1) {
                            if (lane == 0) relcnt[rslot] = 0;
                            issue_stage(qs + D);
                        }
                    }
                    ++qs;
                    if (++rslot == D) { rslot = 0; rphase ^= 1u; }
                }
                if (redo) {
                    wide_flush<RJ, false>(ls, ee, lane, redm, rede, warp);
                } else {
#pragma unroll
                    for (int n = 0; n < RJ; ++n) {
                        const int gq = __double2hiint(mm[n]);
                        ee[n] += (gq >> 20) - 1023;
                        mm[n] = dg_mant(mm[n], gq);
                    }
                    wide_flush<RJ, true>(mm, ee, lane, redm, rede, warp);
                    const unsigned anybad = __ballot_sync(FULL, bad != 0u);
                    if (lane == 0) redb[warp] = anybad ? 1 : 0;
                }
            } else {
                qs += NS;
#ifdef DG_PROFILE
                const long long at0 = clock64();
#endif
                if (!redo) {
                    if (lane == 0) {
                        // trace pages of the rows this round may write (first touch allocates)
                        int nd_ok = nd_max;
                        for (int d = 0; d < nd_max; ++d) {
                            const int step = m + d;
                            if (d == 0 || (step & pmask) == 0)
                                page_cur = trace_page(ptab, step, pshift, rowlen, ct.pool_head, ct.pool_cap);
                            if (page_cur < 0) { nd_ok = d; break; }
                            bc[32 + d] = __longlong_as_double(page_cur);
                        }
                        bc[8] = (double)nd_ok;
                    }
                    __syncwarp();
                    // stream rows of the next rounds: the ring holds steps [m, produced) -> up to m + 2J
                    const int target = min(m + RING, m_end);
                    if (!init && produced < target)
                        wide_produce<false>(dev_stream, produced, target, lane, 32, zbuf, dring, RING, C, Tc, pfac, sqs, kinv,
                                            jcon, ijs, sa.seed, sid, dblk, ublk, m_start);
                }
#ifdef DG_PROFILE
                if (lane == 0) wprof[7] += clock64() - at0;
#endif
            }
            DG_WTICK(2);
            __syncthreads();                                                 // partial likelihoods ready
            DG_WTICK(3);
            if (!redo) {
                int anyb = 0;
                for (int w = 0; w < NCW; ++w) anyb |= redb[w];
                if (anyb) { redo = true; __syncthreads(); continue; }        // rare: some x_n was not a positive normal
            }
            break;
        } while (true);
        if (!init) { const int target = min(m + RING, m_end); if (produced < target) produced = target; }
        ++rounds;
        // ================= decide (every warp, lane j = candidate j) =================
        const int nd_ok = init ? 0 : (int)bc[8];
        double lik = 0.0, P_try = 0.0;
        if (lane < J) {
            if (!redo) {
                double ma = 1.0, mb = 1.0;
                int e = 0;
                int w = 0;
                for (; w + 1 < NCW; w += 2) { ma *= redm[w * J + lane]; mb *= redm[(w + 1) * J + lane]; e += rede[w * J + lane] + rede[(w + 1) * J + lane]; }
                if (w < NCW) { ma *= redm[w * J + lane]; e += rede[w * J + lane]; }
                ma *= mb;
                const int gq = __double2hiint(ma);
                e += (gq >> 20) - 1023;
                lik = fma((double)e, 0.693147180559945309417232121458, log(dg_mant(ma, gq)));
            } else {
                for (int w = 0; w < NCW; ++w) lik += redm[w * J + lane];
            }
            double prior = 0.0;                                              // c ascending (:312-332)
            for (int c = 0; c < C - 1; ++c) prior += priorc[lane * C + c];
            P_try = prior + lik;
        }
        int first = -1, nd = 0;
        double P_new = 0.0, L_new = 0.0;
        if (init) {
            L_new = __shfl_sync(FULL, lik, 0);
            P_new = __shfl_sync(FULL, P_try, 0);                             // (:284-292)
            if (!isfinite(P_new)) flags |= 1;
            first = 0;
        } else {
            bool acc = false;
            if (lane < nd_ok) {
                const double *dl = dring + ((m + lane) % RING) * RL;
                const double lu = dl[CT], Q = dl[CT + 1];
                const double a = ((P_try + Q) - P_now) - Q;                  // (:348)
                acc = (a >= 0.0) || (lu < a);                                // u < exp(min(a, 0)); NaN a: both false
            }
            const unsigned am = __ballot_sync(FULL, acc);
            first = am ? (__ffs(am) - 1) : -1;
            nd = first >= 0 ? first + 1 : nd_ok;
            const unsigned nf = __ballot_sync(FULL, lane < nd && !isfinite(P_try));
            if (nf) flags |= 1;
            if (first >= 0) {
                P_new = __shfl_sync(FULL, P_try, first);
                L_new = __shfl_sync(FULL, lik, first);
            }
        }
        // ================= update: accepted state, trace rows =================
        if (first >= 0) {
            double *snew = state + (cur ^ 1) * 2 * CT;
            for (int i = tid; i < CT; i += nthreads) {
                const int c = i / Tc, t = i - c * Tc;
                snew[CT + i] = ycand[first * CT + i];
                snew[i] = ex[first * CTP + i] / dg_np_sum(ex + first * CTP + t, C, Tc);  // psi (:335)
            }
        }
        DG_WTICK(4);
        __syncthreads();                                                     // new state visible; candidates free
        DG_WTICK(5);
        if (nd > 0) {
            const double *sold = state + cur * 2 * CT, *snw = state + (cur ^ 1) * 2 * CT;
            const int total = nd * rowlen;
            for (int idx = tid; idx < total; idx += nthreads) {
                const int d = idx / rowlen, col = idx - d * rowlen;
                const bool nw_ = d == first;
                const double *sb = nw_ ? snw : sold;
                const double v = col < 2 * CT ? sb[col] : (col == 2 * CT ? (nw_ ? P_new : P_now) : (nw_ ? L_new : L_now));
                double *row = ct.trace + __double_as_longlong(bc[32 + d]) + (size_t)((m + d) & pmask) * rowlen;
                row[col] = v;                                                // Psi_all / Y_all / Pro_all / Lik_all [m+d]  (:362-366)
            }
        }
        if (first >= 0) { cur ^= 1; P_now = P_new; L_now = L_new; if (!init) ++cnt; }
        DG_WTICK(6);
        if (init) { init = false; continue; }
        const bool nomem = nd_ok < nd_max;
        m += nd;
        if (nomem) {                                                         // trace pool exhausted (uniform)
            state_code = DICEGP_CHAIN_NOMEM_;
            m_ret = m - 1;
            break;
        }
        // convergence diagnostics (:369-391) after step m-1 if it is a check step; rows [0, m-1)
        const int mlast = m - 1;
        if (nd > 0 && mlast >= initial && (mlast % gap) == 0) {
            __syncthreads();                                                 // the rows are visible
            const int conv = geweke_converged(gw, gst, CTm, nwarps, warp, lane, tv, mlast);
            if (conv) {
                state_code = DICEGP_CHAIN_CONVERGED_;
                m_ret = mlast;
                break;
            }
            next_check += gap;
        }
    }
    if (state_code == 0 && m >= M) { state_code = DICEGP_CHAIN_EXHAUSTED_; m_ret = M - 1; }
#ifdef DG_PROFILE
    if (lane == 0 && (warp == 0 || is_aux) && sa.prof) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(const_cast<double *>(sa.prof));
        for (int i = 0; i < 10; ++i) if (wprof[i]) atomicAdd(o + i, (unsigned long long)wprof[i]);
        if (warp == 0) { atomicAdd(o + 10, (unsigned long long)(m - m_start)); atomicAdd(o + 11, rounds); }
    }
#endif

    // ---- drain the ring (copies in flight into shared memory), save state ----
    if (!resident) {
        for (int q = qs; q < qs + D; ++q) mbar_wait(dg_smem_u32(fullb + q % D), (uint32_t)((q / D) & 1));
    } else if (D > 0) {
        for (int q = 0; q < D; ++q) mbar_wait(dg_smem_u32(fullb + q), 0u);
    }
    __syncthreads();
    {
        const double *snow = state + cur * 2 * CT;
        for (int i = tid; i < CT; i += nthreads) { st[i] = snow[CT + i]; st[CT + i] = snow[i]; }
    }
    if (lane == 0 && ct.ctr) {
        if (tma_bytes) atomicAdd(ct.ctr + 0, tma_bytes);
    }
    if (tid == 0) {
        st[2 * CT] = P_now; st[2 * CT + 1] = L_now;
        ct.m_next[chain] = m;
        ct.cnt[chain] = cnt;
        ct.flags[chain] = flags;
        ct.m_ret[chain] = m_ret;
        ct.state[chain] = state_code;
        if (ct.ctr) {
            atomicAdd(ct.ctr + 1, rounds);
            atomicAdd(ct.ctr + 2, rounds * (unsigned long long)J * (unsigned long long)(rel[Tc]));
            atomicAdd(ct.ctr + 3, (unsigned long long)(m - m_start));
            atomicAdd(ct.ctr + 4, rounds * (unsigned long long)J * (unsigned long long)(rel[Tc]) * (unsigned long long)C);
        }
        for (int s = 0; s < D; ++s)
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(dg_smem_u32(fullb + s)) : "memory");
    }
    __syncthreads();
}

}  // namespace dg

// Persistent blocks pull chains from `queue` (device-stream runs) or block b advances chain ids[b].
// ring_bytes: shared memory behind the largest fixed part of the launch's chains, for the W ring.
template <int KC, int RJ>
__global__ void __launch_bounds__(DG_WIDE_THREADS, 1)
dicegp_wide_kernel(GeneTab gt, ChainTab ct, StreamArgs sa, const int32_t *__restrict__ ids, int n_ids, int n_steps,
                   int ring_bytes, int *queue) {
    extern __shared__ __align__(128) double dg_smem[];
    __shared__ int next_slot;
    int slot = blockIdx.x;
    while (true) {
        if (queue != nullptr) {
            if (threadIdx.x == 0) next_slot = atomicAdd(queue, 1);
            __syncthreads();
            slot = next_slot;
            __syncthreads();
        }
        if (slot >= n_ids) break;
        if (ct.state[ids[slot]] == 0) dg::run_chain_wide<KC, RJ>(gt, ct, sa, ids[slot], slot, n_steps, dg_smem, ring_bytes);
        __syncthreads();
        if (queue == nullptr) break;
    }
}
