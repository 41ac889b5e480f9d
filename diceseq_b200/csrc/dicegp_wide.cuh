// dicegp_wide.cuh -- the MH loop of a joint (time-series) chain, kernel v3: WIDE rejection-chain rounds
// without warp specialisation, W streamed through per-warp TMA rings.
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
// One block = NCW consumer warps + 1..6 producer ("auxiliary") warps; every consumer warp does every phase:
//   S      two sub-phases over all threads: item (j, c, t): candidate y -> exp and its GP-prior term; then
//          item (j, c, t): length-weighted softmax f[t][c][j], item (j, c): prior of one isoform
//   LIK    reads in TASKS of 16 words (32 reads) of one time point.  W is read from a TILED copy (GeneTab::Wt,
//          made on the device after the upload): a tile = the C rows of one task back to back, tiles in
//          (time, position) order, so the contiguous run of tasks a warp owns is ONE byte range.  Each consumer
//          warp streams its run through its OWN ring of two-task slots with 1-D TMA bulk copies
//          (cp.async.bulk + mbarrier complete_tx, one copy per slot, issued by lane 0 when the slot has been
//          consumed; the ring cursor is kept in running counters, no divisions); the sequence repeats every round, so the first slots of the next round are in flight
//          while this round decides; a run that fits the ring is loaded once (W resident).  No block-level
//          synchronisation inside the pass.
//          Lane (s, h), s = lane & 7, h = lane >> 3: words s and s + 8 of the task (4 reads), candidates
//          4h .. 4h+3 (RJ = 4) -- a 4 x RJ register tile per lane.
//          prod_n x_n per candidate as (mantissa product, exponent sum); anything that is not a positive
//          normal number sends the round to the real-log path (-inf / NaN like np.log(...).sum()).
//   AUX    the stream of the NEXT rounds: the consumer warps make the Philox / Box-Muller normals (one pair per
//          thread: ~150 dependent fp64 operations each) before their tasks, the auxiliary warp turns them into
//          increments, log u and the proposal log-density Q, and reserves the trace pages of this round's rows.
//   DECIDE every warp redundantly: lane j combines the per-warp partial products of candidate j, takes one
//          log, adds the prior, tests  log u < ((P_try + Q) - P_now) - Q ; a ballot gives d*.
//   UPDATE psi of the accepted state, trace rows (coalesced, all threads), Geweke rule when due.
// Five block barriers per round.  Same stream, same saved state, same paged trace and incremental Geweke
// sums as dg::run_chain -- a chain can be advanced by either kernel in turn.
//
// Measured (profiles/): per round at C = 8, 15 consumer warps + 1 producer at 128 registers: S 8k,
// likelihood 48k (tasks 38k: the tile loop is bound by the shared-memory and fp64 pipes together), decide + update 5k,
// rows 3k cycles for ~13 steps at the acceptance rates of the straggler pass = 2.7 us per step.  Used for the straggler
// pass (one chain per SM, W L2 resident) and, with few chains per SM, from step 1024 on; in the throughput pass of a full
// batch the general kernel's 2-3 chains per SM at 4 candidates per round spend a third less fp64 work per committed
// step and win.
#pragma once
#include "dicegp_chain.cuh"

#define DG_WIDE_SLOTS 8             // ring slots (tasks in flight) per consumer warp
#define DG_WIDE_MAXCT 128           // latent values (C * Tc) of a chain this kernel takes
#ifndef DG_WIDE_THREADS
#define DG_WIDE_THREADS 512         // launch bound: 16 warps x 128 registers (registers come in units of 16 per thread: 17-18
                                    // warps would get 96 and spill into the task loop; measured 98 -> 90 ms on the straggler pass)
#endif
#define DG_WIDE_MAXCS 16            // blocks of a thread-block cluster that share one chain (16 = non-portable size, the largest genes)

#ifdef DG_PROFILE
// warp 0 lane 0: 0 S, 1 wait f, 2 likelihood, 3 wait partials, 4 decide + new state, 5 wait state, 6 rows; aux lane 0: 7 busy
#define DG_WTICK(i) do { if (lane == 0 && warp == 0) { long long t__ = clock64(); wprof[i] += t__ - wt0; wt0 = t__; } } while (0)
#else
#define DG_WTICK(i) do { } while (0)
#endif

namespace dg {

struct WideLayout {
    int state, ycand, ex, pr, ff, priorc, redm, rede, redb, xm, xe, xb, bc, zbuf, dring, kinv, pfac, ijs, jcon, sqs, len, ints, gw, bars;
    int ring;                   // doubles before the ring (128-byte aligned)
};

__host__ __device__ inline WideLayout dg_wide_layout(int C, int Tc, int ncw, int J, int npw = 1) {
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
    // cluster launches: per-block partial likelihoods of the J candidates, exchanged through DSMEM (two buffers)
    L.xm = o; o += 2 * DG_WIDE_MAXCS * J;
    L.xe = o; o += DG_WIDE_MAXCS * J;                  // 2 x MAXCS x J ints
    L.xb = o; o += DG_WIDE_MAXCS;                      // 2 x MAXCS ints
    L.bc = o; o += 96;
    L.zbuf = o; o += 2 * J * RL;
    L.dring = o; o += 2 * J * RL;
    L.kinv = o; o += Tc * Tc;
    L.pfac = o; o += Tc * Tc;
    L.ijs = o; o += C; L.jcon = o; o += C; L.sqs = o; o += C;
    L.len = o; o += CT;
    L.ints = o; o += (3 * (Tc + 1) + 1) / 2 + 1;       // rel[Tc+1], ncnt[Tc], cum[Tc+1]
    L.gw = o; o += 2 * (ncw + npw) * (ctm < DG_GW_COLS ? (ctm > 0 ? ctm : 1) : DG_GW_COLS);
    L.bars = o; o += (ncw + npw) * DG_WIDE_SLOTS;                 // per consumer warp: one mbarrier per ring slot
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
// BLOCK: all threads of the block produce (start-up); else the `nprod` threads of the producer warps: one warp
// synchronises with __syncwarp, several through named barrier 2
template <bool BLOCK>
__device__ __forceinline__ void wide_sync(int nprod) {
    if (BLOCK) __syncthreads();
    else if (nprod <= 32) __syncwarp();
    else dg_bar(2, nprod);
}

// The Philox / Box-Muller normals of steps [s0, s1) into zbuf rows 0 .. s1-s0 (one item per pair: ~150 dependent
// fp64 operations, so the items are spread over as many threads as there are).
__device__ __noinline__ void wide_normals(int s0, int s1, int pt, int nprod, double *zbuf, int C, int Tc, uint64_t seed, int64_t sid) {
    const int CT = C * Tc, CTm = CT - Tc, RL = CT + 2, ns = s1 - s0;
    const int npair = (CTm + 1) >> 1;
    for (int q = pt; q < ns * npair; q += nprod) {
        const int ds = q / npair, i = q - ds * npair;
        double z0, z1;
        dg_normal_pair(seed, sid, (uint32_t)(s0 + ds), (uint32_t)i, z0, z1);
        zbuf[ds * RL + 2 * i] = z0;
        if (2 * i + 1 < CTm) zbuf[ds * RL + 2 * i + 1] = z1;
    }
}

template <bool BLOCK>
__device__ __noinline__ void wide_produce(bool dev_stream, int s0, int s1, int pt, int nprod, double *zbuf, double *dring,
                                          int RING, int C, int Tc, const double *pfac, const double *sqs, const double *kinv,
                                          const double *jcon, const double *ijs, uint64_t seed, int64_t sid,
                                          const double *dblk, const double *ublk, int m_start, bool have_z) {
    const int CT = C * Tc, CTm = CT - Tc, RL = CT + 2, ns = s1 - s0;
    if (ns <= 0) return;
    if (dev_stream) {
        if (!have_z) wide_normals(s0, s1, pt, nprod, zbuf, C, Tc, seed, sid);
        for (int ds = pt; ds < ns; ds += nprod)
            dring[((s0 + ds) % RING) * RL + CT] = dg_log(dg_uniform(seed, sid, (uint32_t)(s0 + ds)));
        wide_sync<BLOCK>(nprod);
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
    wide_sync<BLOCK>(nprod);
    for (int q = pt; q < ns * CTm; q += nprod) {                   // per (step, c, t): (d K^-1)[t]
        const int ds = q / CTm, i = q - ds * CTm;
        const int c = i / Tc, t = i - c * Tc;
        const double *d = dring + ((s0 + ds) % RING) * RL + c * Tc;
        double r = 0.0;
        for (int j = 0; j < Tc; ++j) r = fma(d[j], kinv[j * Tc + t], r);
        zbuf[ds * RL + i] = r;                                     // the normals are consumed
    }
    wide_sync<BLOCK>(nprod);
    for (int q = pt; q < ns * (C - 1); q += nprod) {               // per (step, c): d K^-1 d -> proposal log-density term
        const int ds = q / (C - 1), c = q - ds * (C - 1);
        const double *d = dring + ((s0 + ds) % RING) * RL + c * Tc;
        double *rr = zbuf + ds * RL + c * Tc;
        double acc = 0.0;
        for (int t = 0; t < Tc; ++t) acc = fma(rr[t], d[t], acc);
        rr[0] = jcon[c] - (acc * ijs[c]) / 2.0;
    }
    wide_sync<BLOCK>(nprod);
    for (int ds = pt; ds < ns; ds += nprod) {
        double Q = 0.0;
        for (int c = 0; c < C - 1; ++c) Q += zbuf[ds * RL + c * Tc];
        dring[((s0 + ds) % RING) * RL + CT + 1] = Q;
    }
    wide_sync<BLOCK>(nprod);
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
// CLUSTER: a thread-block cluster of cs blocks advances ONE chain (the last stragglers of a batch, when there are
// fewer chains than SMs): every block runs the same serial phases on the same inputs (same bits), the likelihood
// tasks are split over the consumer warps of all blocks, every block folds its warps' partial products into one
// (mantissa, exponent) pair per candidate and stores it into the shared memory of all blocks (DSMEM, two buffers),
// one barrier.cluster per pass over W.  Only the lead block writes trace rows, runs the Geweke rule (its verdict is
// broadcast through DSMEM) and saves the state; trace pages are pre-assigned by the host.
template <int KC, int RJ, bool CLUSTER>
__device__ void run_chain_wide(const GeneTab &gt, const ChainTab &ct, const StreamArgs &sa, int chain, int slot_id,
                               int n_steps, double *smem, int ring_bytes, int npw, int cs_rt, int crank_rt) {
    constexpr int J = 4 * RJ, RING = 2 * J;
    constexpr unsigned FULL = 0xffffffffu;
    const int cs = CLUSTER ? cs_rt : 1, crank = CLUSTER ? crank_rt : 0;
    const bool lead = crank == 0;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    // warps [0, NCW) consume W; warps [NCW, nwarps) produce the stream of the next rounds (the first of them also
    // reserves the trace pages)
    const int NCW = nwarps - npw;
    const bool is_aux = warp >= NCW;
    const bool aux0 = warp == NCW;
    int xbuf = 0;                                         // exchange buffer of the next DSMEM exchange

    const int g = ct.gene[chain], t0 = ct.t0[chain];
    const int C = KC > 0 ? KC : gt.C[g];
    const int Tc = ct.Tc[chain], T = gt.T;
    const int CT = C * Tc, CTm = (C - 1) * Tc, RL = CT + 2, CTP = CT | 1;
    const WideLayout L = dg_wide_layout(C, Tc, NCW, J, npw);
    double *state = smem + L.state, *ycand = smem + L.ycand, *ex = smem + L.ex, *pr = smem + L.pr, *ff = smem + L.ff;
    double *priorc = smem + L.priorc, *redm = smem + L.redm, *bc = smem + L.bc, *zbuf = smem + L.zbuf, *dring = smem + L.dring;
    double *kinv = smem + L.kinv, *pfac = smem + L.pfac, *ijs = smem + L.ijs, *jcon = smem + L.jcon, *sqs = smem + L.sqs;
    double *lens = smem + L.len, *gw = smem + L.gw;
    int *rede = reinterpret_cast<int *>(smem + L.rede), *redb = reinterpret_cast<int *>(smem + L.redb);
    double *xm = smem + L.xm;
    int *xe = reinterpret_cast<int *>(smem + L.xe), *xb = reinterpret_cast<int *>(smem + L.xb);
    int *rel = reinterpret_cast<int *>(smem + L.ints), *ncnt = rel + (Tc + 1), *cum = ncnt + Tc;
    unsigned long long *mybars = reinterpret_cast<unsigned long long *>(smem + L.bars) + warp * DG_WIDE_SLOTS;
    double2 *ring2 = reinterpret_cast<double2 *>(smem + L.ring);
    const double prior_const = ct.prior_const[chain];

    const int32_t *roff = gt.roff + (size_t)g * (T + 1) + t0;
    const int r_first = roff[0];

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
    if (m == 0 && lead && tid < 4) gst[tid] = 0.0;
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
    const int32_t *toffg = gt.toff + (size_t)g * (T + 1) + t0;
    for (int i = tid; i <= Tc; i += nthreads) cum[i] = toffg[i] - toffg[0];      // first tile (task) of every time point
    __syncthreads();
    // ---- this warp's share of W: a contiguous run of tasks, streamed through its OWN ring of TMA slots ----
    const int NT = cum[Tc];                               // tasks per round
    const int per = (NT + cs * NCW - 1) / (cs * NCW);     // ... split over the consumer warps of the whole cluster
    const int ta = is_aux ? 0 : min((crank * NCW + warp) * per, NT), tb = is_aux ? 0 : min(ta + per, NT);
    const int ntw = tb - ta;                              // tasks of this warp per round
    const int task_w2 = C * 16;                           // double2 per task (= tile): [isoform][16 words]
    const int warp_ring_w2 = ((ring_bytes / NCW) / 16) & ~7;
    double2 *myring = ring2 + (size_t)warp * warp_ring_w2;
    const int cap = warp_ring_w2 / task_w2;               // tasks the ring holds
    const bool resident = ntw <= cap;                     // all of them: loaded once, one barrier, one copy
    // streaming: chunks of G consecutive tasks (one bulk copy each: the tiles of a gene are contiguous), SL slots
    const int G = resident ? max(ntw, 1) : max(1, min(4, cap / 2));
    const int SL = resident ? 1 : min(cap / G, DG_WIDE_SLOTS);
    const int NC = (ntw + G - 1) / G;                     // chunks per round
    const double2 *mysrc = reinterpret_cast<const double2 *>(gt.Wt + gt.wtoff[g]) + (size_t)(toffg[0] + ta) * task_w2;
    unsigned long long tma_chunks = 0;                    // bulk copies issued by this warp (lane 0)
    // chunk `c` of this warp's run into ring slot `slot`: ONE bulk copy, issued by lane 0
    auto issue_chunk = [&](int c, int slot) {
        if (lane == 0) {
            const int first = c * G, n = min(G, ntw - first);
            const uint32_t bytes = (uint32_t)(n * task_w2) * 16u;
            const uint32_t bar = dg_smem_u32(mybars + slot);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dg_smem_u32(myring + (size_t)slot * G * task_w2)), "l"(mysrc + (size_t)first * task_w2), "r"(bytes), "r"(bar)
                : "memory");
            ++tma_chunks;
        }
    };
    if (!is_aux && ntw > 0) {
        if (lane == 0) {
            for (int s2 = 0; s2 < SL; ++s2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dg_smem_u32(mybars + s2)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        for (int i = 0; i < SL; ++i) issue_chunk(i % NC, i);
    }
    // ring cursor, kept as running counters (no divisions in the loop): slot and mbarrier phase of the next chunk
    // this warp consumes, and the index of the chunk that will be issued into the slot it frees
    int rslot = 0, rphase = 0, cnext = SL % max(NC, 1);

    const int m_start = m;
    const int m_end = (m_start + n_steps < M) ? (m_start + n_steps) : M;
    int state_code = 0, m_ret = 0;
    long long page_cur = -1;                              // aux lane 0
    const double *dblk = dev_stream ? nullptr : sa.delta + sa.delta_off[slot_id];
    const double *ublk = dev_stream ? nullptr : sa.u + sa.u_off[slot_id];
    int produced = min(m + RING, m_end);                  // ring holds steps [.., produced)
    wide_produce<true>(dev_stream, m, produced, tid, nthreads, zbuf, dring, RING, C, Tc, pfac, sqs, kinv, jcon, ijs,
                       sa.seed, sid, dblk, ublk, m_start, false);

    unsigned long long rounds = 0;
    bool init = (m == 0);                                 // evaluate the start value first (model_GP.py:275-292)
    int next_check = m > initial ? m : initial;
    next_check += (gap - next_check % gap) % gap;

#ifdef DG_PROFILE
    long long wprof[16];
    for (int i = 0; i < 16; ++i) wprof[i] = 0;
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
        int xcur = 0;                                                        // buffer of the last DSMEM exchange
        do {
            // the stream rows of the next rounds (ring holds [m, produced) -> up to m + 2J): the normals are made by
            // the consumer warps (one pair per thread) before their likelihood tasks, the rest by the aux warp
            const int target = min(m + RING, m_end);
            const bool top_up = !redo && !init && produced < target;
            if (!is_aux) {
#ifdef DG_PROFILE
                long long nt0 = clock64();
#endif
                if (top_up) {
                    if (dev_stream) wide_normals(produced, target, tid, NCW * 32, zbuf, C, Tc, sa.seed, sid);
                    __threadfence_block();
                    dg_bar_arrive(1, nthreads);                              // the aux warp waits for them
                }
#ifdef DG_PROFILE
                if (lane == 0 && warp == 0) wprof[12] += clock64() - nt0;
#endif
                double mm[RJ], ls[RJ];
                int ee[RJ];
                unsigned bad = 0u;
#pragma unroll
                for (int n = 0; n < RJ; ++n) { mm[n] = 1.0; ee[n] = 0; ls[n] = 0.0; }
                int tt = 0, it = 0;
                for (int c = 0; c < NC; ++c) {
                    const int slot = resident ? 0 : rslot;
#ifdef DG_PROFILE
                    long long lt0 = clock64();
#endif
                    mbar_wait(dg_smem_u32(mybars + slot), resident ? 0u : (uint32_t)rphase);
#ifdef DG_PROFILE
                    long long lt1 = clock64();
                    if (lane == 0 && warp == 0) wprof[8] += lt1 - lt0;
#endif
                    const int first = c * G, n = min(G, ntw - first);
                    for (int i = 0; i < n; ++i) {
                        const int task = ta + first + i;
                        while (task >= cum[tt + 1]) ++tt;
                        const int np = (rel[tt + 1] - rel[tt]) >> 1;
                        const int w0 = (task - cum[tt]) * 16;
                        const int nw = min(16, np - w0);
                        const int padw = ((ncnt[tt] & 1) && (w0 + nw == np)) ? nw - 1 : -1;
                        const double2 *wt = myring + (size_t)(slot * G + i) * task_w2;
                        if (redo) wide_task<KC, RJ, true>(wt, 16, C, ff + tt * C * J, lane, nw, padw, mm, ee, bad, ls);
                        else wide_task<KC, RJ, false>(wt, 16, C, ff + tt * C * J, lane, nw, padw, mm, ee, bad, ls);
                        if (!redo && (++it & 127) == 0) {                    // keep the running products in range
#pragma unroll
                            for (int n2 = 0; n2 < RJ; ++n2) {
                                const int gq = __double2hiint(mm[n2]);
                                ee[n2] += (gq >> 20) - 1023;
                                mm[n2] = dg_mant(mm[n2], gq);
                            }
                        }
                    }
#ifdef DG_PROFILE
                    long long lt2 = clock64();
                    if (lane == 0 && warp == 0) wprof[9] += lt2 - lt1;
#endif
                    if (!resident) {
                        // the slot is free: refill it with the chunk SL ahead (the sequence repeats every round, so the
                        // first chunks of the next round are on their way while this round decides)
                        __syncwarp();
                        issue_chunk(cnext, slot);
                        if (++cnext == NC) cnext = 0;
                        if (++rslot == SL) { rslot = 0; rphase ^= 1; }
#ifdef DG_PROFILE
                        if (lane == 0 && warp == 0) wprof[13] += clock64() - lt2;
#endif
                    }
                }
#ifdef DG_PROFILE
                long long ft0 = clock64();
#endif
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
#ifdef DG_PROFILE
                if (lane == 0 && warp == 0) wprof[14] += clock64() - ft0;
#endif
            } else {
#ifdef DG_PROFILE
                const long long at0 = clock64();
#endif
                if (!redo) {
                    if (aux0 && lane == 0) {
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
                    if (top_up) {
                        dg_bar(1, nthreads);                                 // normals in zbuf
                        wide_produce<false>(dev_stream, produced, target, tid - NCW * 32, npw * 32, zbuf, dring, RING, C, Tc, pfac, sqs,
                                            kinv, jcon, ijs, sa.seed, sid, dblk, ublk, m_start, true);
                    }
                }
#ifdef DG_PROFILE
                if (aux0 && lane == 0) wprof[7] += clock64() - at0;
#endif
            }
            DG_WTICK(2);
            __syncthreads();                                                 // partial likelihoods ready
            DG_WTICK(3);
            if (CLUSTER && cs > 1) {
                // fold this block's partials into one (mantissa, exponent) pair -- or, on the log path, one sum -- per
                // candidate and store it into every block of the cluster; the buffers alternate between exchanges
                if (warp == 0) {
                    if (lane < J) {
                        double v = 0.0;
                        int e = 0;
                        if (!redo) {
                            double ma = 1.0, mb = 1.0;
                            int w = 0;
                            for (; w + 1 < NCW; w += 2) { ma *= redm[w * J + lane]; mb *= redm[(w + 1) * J + lane]; e += rede[w * J + lane] + rede[(w + 1) * J + lane]; }
                            if (w < NCW) { ma *= redm[w * J + lane]; e += rede[w * J + lane]; }
                            ma *= mb;
                            const int gq = __double2hiint(ma);
                            e += (gq >> 20) - 1023;
                            v = dg_mant(ma, gq);
                        } else {
                            for (int w = 0; w < NCW; ++w) v += redm[w * J + lane];
                        }
                        double *dm = xm + (xbuf * DG_WIDE_MAXCS + crank) * J + lane;
                        int *de = xe + (xbuf * DG_WIDE_MAXCS + crank) * J + lane;
                        for (int r = 0; r < cs; ++r) { st_cluster(dm, (unsigned)r, v); st_cluster(de, (unsigned)r, e); }
                    }
                    if (lane == 0) {
                        int anyb = 0;
                        if (!redo) for (int w = 0; w < NCW; ++w) anyb |= redb[w];
                        for (int r = 0; r < cs; ++r) st_cluster(xb + xbuf * DG_WIDE_MAXCS + crank, (unsigned)r, anyb);
                    }
                }
                cluster_barrier();
                xcur = xbuf;
                xbuf ^= 1;
            }
            if (!redo) {
                int anyb = 0;
                if (CLUSTER && cs > 1) { for (int r = 0; r < cs; ++r) anyb |= xb[xcur * DG_WIDE_MAXCS + r]; }
                else { for (int w = 0; w < NCW; ++w) anyb |= redb[w]; }
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
            if (CLUSTER && cs > 1) {
                const double *pm = xm + xcur * DG_WIDE_MAXCS * J + lane;
                const int *pe = xe + xcur * DG_WIDE_MAXCS * J + lane;
                if (!redo) {
                    double ma = 1.0;
                    int e = 0;
                    for (int r = 0; r < cs; ++r) { ma *= pm[r * J]; e += pe[r * J]; }          // <= 8 mantissas: < 2^8
                    const int gq = __double2hiint(ma);
                    e += (gq >> 20) - 1023;
                    lik = fma((double)e, 0.693147180559945309417232121458, log(dg_mant(ma, gq)));
                } else {
                    for (int r = 0; r < cs; ++r) lik += pm[r * J];
                }
            } else if (!redo) {
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
        if (nd > 0 && lead) {
            const double *sold = state + cur * 2 * CT, *snw = state + (cur ^ 1) * 2 * CT;
            // one warp per row: the row address is formed once, the lanes walk the columns (coalesced, no divisions)
            for (int d = warp; d < nd; d += nwarps) {
                const bool nw_ = d == first;
                const double *sb = nw_ ? snw : sold;
                double *row = ct.trace + __double_as_longlong(bc[32 + d]) + (size_t)((m + d) & pmask) * rowlen;
                for (int col = lane; col < rowlen; col += 32)                    // Psi_all / Y_all / Pro_all / Lik_all [m+d]  (:362-366)
                    row[col] = col < 2 * CT ? sb[col] : (col == 2 * CT ? (nw_ ? P_new : P_now) : (nw_ ? L_new : L_now));
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
            int conv;
            if (CLUSTER && cs > 1) {                                         // the lead block decides for the cluster
                if (lead) {
                    conv = geweke_converged(gw, gst, CTm, nwarps, warp, lane, tv, mlast);
                    if (tid == 0) for (int r = 0; r < cs; ++r) st_cluster(bc + 9, (unsigned)r, (double)conv);
                }
                cluster_barrier();
                conv = (int)bc[9];
            } else {
                conv = geweke_converged(gw, gst, CTm, nwarps, warp, lane, tv, mlast);
            }
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
    if (lane == 0 && (warp == 0 || aux0) && sa.prof) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(const_cast<double *>(sa.prof));
        for (int i = 0; i < 10; ++i) if (wprof[i]) atomicAdd(o + i, (unsigned long long)wprof[i]);
        for (int i = 12; i < 16; ++i) if (wprof[i]) atomicAdd(o + i, (unsigned long long)wprof[i]);
        if (warp == 0) { atomicAdd(o + 10, (unsigned long long)(m - m_start)); atomicAdd(o + 11, rounds); }
    }
#endif

    // ---- drain the ring (copies in flight into shared memory), save state ----
    if (!is_aux && ntw > 0) {
        if (resident) mbar_wait(dg_smem_u32(mybars), 0u);
        else for (int q = 0; q < SL; ++q) {                   // the SL chunks in flight, oldest first
            mbar_wait(dg_smem_u32(mybars + rslot), (uint32_t)rphase);
            if (++rslot == SL) { rslot = 0; rphase ^= 1; }
        }
        __syncwarp();
        if (lane == 0)
            for (int s2 = 0; s2 < SL; ++s2) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(dg_smem_u32(mybars + s2)) : "memory");
    }
    __syncthreads();
    if (lead) {
        const double *snow = state + cur * 2 * CT;
        for (int i = tid; i < CT; i += nthreads) { st[i] = snow[CT + i]; st[CT + i] = snow[i]; }
    }
    if (lane == 0 && ct.ctr) {
        // whole runs of NC chunks (ntw tasks) plus the first chunks (G tasks each) of the run that was cut
        if (tma_chunks) atomicAdd(ct.ctr + 0, ((tma_chunks / NC) * (unsigned long long)ntw + (tma_chunks % NC) * (unsigned long long)G) *
                                                  (unsigned long long)task_w2 * 16ull);
    }
    if (tid == 0 && lead) {
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
    }
    __syncthreads();
}

}  // namespace dg

// Persistent blocks (clusters) pull chains from `queue` (device-stream runs) or block (cluster) b advances chain ids[b].
// ring_bytes: shared memory behind the largest fixed part of the launch's chains, for the W ring.
template <int KC, int RJ, bool CLUSTER = false>
__global__ void __launch_bounds__(DG_WIDE_THREADS, 1)
dicegp_wide_kernel(GeneTab gt, ChainTab ct, StreamArgs sa, const int32_t *__restrict__ ids, int n_ids, int n_steps,
                   int ring_bytes, int npw, int *queue) {
    extern __shared__ __align__(128) double dg_smem[];
    __shared__ int next_slot;
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
                    for (int r = 0; r < cs; ++r) dg::st_cluster(&next_slot, (unsigned)r, sl);
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
        if (running) dg::run_chain_wide<KC, RJ, CLUSTER>(gt, ct, sa, ids[slot], slot, n_steps, dg_smem, ring_bytes, npw, cs, crank);
        if (cs == 1) __syncthreads(); else dg::cluster_barrier();
        if (queue == nullptr) break;
    }
}
