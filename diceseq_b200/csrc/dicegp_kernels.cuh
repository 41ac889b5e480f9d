// dicegp_kernels.cuh -- sm_100a device code of the DICEseq GP/MH sampler.
//
// One MH chain (= one Psi_GP_MH call, diceseq/models/model_GP.py:188-395) is advanced
// by ONE thread block.  The chain's read-by-isoform matrix W is staged once per launch
// into shared memory with a 1-D TMA bulk copy (cp.async.bulk + mbarrier) and re-read
// from there every step; chains whose W does not fit stream it from L2/HBM instead.
// All arithmetic is fp64.  No tensor cores: the inner product has length C <= 32 and is
// followed by a logarithm, it is not a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DG_MAXC 32
#define DG_MAXT 32
#define DG_MAXCT 512          // C * Tc of one chain
#define DG_MAXTHREADS 1024
#define DG_MAXWARPS 32
#define DG_CLUSTER_SLOTS 128   // likelihood warps of one chain spread over a thread-block cluster (<= 8 x 16)
#define DG_GW_COLS 32         // Geweke column tile
#define DG_STREAM_BLOCK 8      // random-stream steps produced per batch (ring = 2 batches)

// ---------------------------------------------------------------------------------------
// tables (device pointers, passed by value to the kernels)
// ---------------------------------------------------------------------------------------
struct GeneTab {
    const int32_t *C;        // [G]
    const int64_t *woff;     // [G]   first double of the gene's W
    const int32_t *roff;     // [G*(T+1)] padded read offset of time t inside the gene
    const int32_t *ncnt;     // [G*T] reads at time t
    const int64_t *lenoff;   // [G]   first double of the gene's len_iso (T x C)
    const double *W;
    const double *len;
    int32_t T;
    // tiled copy of W for the wide-round kernel (built on the device after the upload): the isoform-major block
    // of every time point cut into TILES of 16 words (32 reads); a tile holds its C rows back to back
    // ([k][16 words], zero padded), tiles follow each other in (time, position) order, so any run of
    // consecutive tiles of a gene is ONE contiguous range (one TMA bulk copy).
    const double *Wt;
    const int64_t *wtoff;    // [G]        first double of the gene's tiles
    const int32_t *toff;     // [G*(T+1)]  tiles of the gene before time t
};

struct ChainTab {
    // immutable description
    const int32_t *gene, *t0, *Tc, *M, *initial, *gap;
    const int64_t *kinv_off, *pf_off, *js_off, *jc_off, *y0_off;
    const double *prior_const;
    const double *pool;
    const int64_t *st_off;       // first double of the chain's saved state
    const int64_t *stream_id;
    // paged trace pool: a chain's rows live in pages of 2^shift rows; page p of chain i is
    // at trace[ptab[ptab_off[i] + p]] (-1 = not allocated yet; the first page always is).
    // Further pages are bump-allocated on the device as a chain grows.
    const int64_t *ptab_off;
    const int32_t *pshift;
    int64_t *ptab;
    unsigned long long *pool_head;   // next free double of the trace pool
    int64_t pool_cap;                // doubles
    // mutable
    double *trace;
    double *st;                  // ynow[C*Tc], psinow[C*Tc], P_now, L_now
    int32_t *m_next, *cnt, *state, *m_ret, *flags;
    // work counters (may be NULL): [0] bytes of W brought from L2 / HBM into shared memory or registers,
    // [1] rounds (passes over W), [2] candidate evaluations (rounds x candidates x reads), [3] steps,
    // [4] fp64 fused multiply-adds of those evaluations (x isoforms)
    unsigned long long *ctr;
};

struct StreamArgs {
    // host stream (mode 0)
    const double *delta;
    const int64_t *delta_off;    // per launch slot
    const double *u;
    const int64_t *u_off;        // per launch slot
    // device stream (mode 1)
    uint64_t seed;
    int32_t mode;
    const double *prof;          // DG_PROFILE builds: 16 x uint64 cycle counters (else NULL)
};

// Read-side view of one chain's paged trace.  The trace is written by the very kernel that
// reads it back for the Geweke test: loads go through ld.global.cg (L2), never through
// the non-coherent path.
struct TraceView {
    const double *pool;
    const int64_t *tab;
    int shift, rowlen;
#ifdef __CUDACC__
    __device__ __forceinline__ const double *row(int r) const {
        const int64_t base = __ldcg(tab + (r >> shift));
        return pool + base + (size_t)(r & ((1 << shift) - 1)) * rowlen;
    }
#endif
};

// ---------------------------------------------------------------------------------------
// shared memory carve-up (host and device must agree)
// ---------------------------------------------------------------------------------------
struct SmLayout {
    // offsets in doubles from the start of dynamic shared memory
    int ycand, psic, gs, ex, pr;                               // J*CT each (J = 2^D - 1 candidates)
    int ff;                                                    // CT*(J+1): f of all candidates, [t][k][j]
    int zring, dring;                                          // stream: one batch of z / scratch; ring of delta, log u, Q
    int kinv, pfac;                                            // Tc*Tc each
    int ijs, jcon, sqs;                                        // C each
    int red, redi, redb, bc, gw, rel, seg, mbar;               // scratch
    int w;                                                     // W (resident tiers)
    int total;                                                 // doubles without W
};

// rows of the stream ring for rounds that commit up to J steps: a batch of DG_STREAM_BLOCK steps
// is refilled while the next round already reads its own (up to J) rows
__host__ __device__ constexpr int dg_ring_rows(int J) {
    return ((DG_STREAM_BLOCK + 2 * J - 1 + DG_STREAM_BLOCK - 1) / DG_STREAM_BLOCK) * DG_STREAM_BLOCK;
}

// J = candidate nodes evaluated per round (per pass over W)
__host__ __device__ inline SmLayout dg_layout(int C, int Tc, int nwarps, int J, int cs = 1) {
    SmLayout L;
    // partial-likelihood slots: one per likelihood warp of the chain; a cluster of cs blocks
    // shares them (DG_CLUSTER_SLOTS) and double-buffers them across rounds
    const int GWS = cs > 1 ? DG_CLUSTER_SLOTS : DG_MAXWARPS, NBUF = cs > 1 ? 2 : 1;
    const int JP = (J + 1) & ~1, RING = dg_ring_rows(J);
    int CT = C * Tc, TT = Tc * Tc, o = 0;
    L.ycand = o; o += J * CT;  L.psic = o; o += J * CT;  L.gs = o; o += J * CT;  L.ex = o; o += J * CT;
    L.pr = o; o += J * CT;
    o = (o + 1) & ~1;
    L.ff = o; o += CT * JP;                    // 16-byte aligned rows of JP doubles
    L.zring = o; o += DG_STREAM_BLOCK * (CT + 2);  L.dring = o; o += RING * (CT + 2);
    L.kinv = o; o += TT;  L.pfac = o; o += TT;
    L.ijs = o; o += C;    L.jcon = o; o += C;  L.sqs = o; o += C;
    L.red = o; o += NBUF * J * GWS;
    L.redi = o; o += NBUF * J * GWS / 2 + 1;   // ints
    L.redb = o; o += NBUF * GWS / 2;           // unsigned
    L.bc = o; o += 64;
    { const int ctm = CT - Tc; L.gw = o; o += 2 * nwarps * (ctm < DG_GW_COLS ? (ctm > 0 ? ctm : 1) : DG_GW_COLS); }
    L.rel = o; o += Tc + 2;                    // ints: rel roff[Tc+1], ncnt[Tc]
    L.seg = o; o += (DG_MAXT + 2) / 2 + 1;     // ints: first 32-pair item of each time point
    L.mbar = o; o += 2;
    o = (o + 15) & ~15;                         // 128-byte align W
    L.w = o;
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t dg_smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// numpy's pairwise summation for n <= 128 (np.sum of a short 1-D array): plain loop
// below 8 elements, 8 interleaved accumulators otherwise (numpy loops_utils.h,
// probed in this image).  Used for the softmax normalisers so that psi follows the
// reference's rounding (model_GP.py:335-337).
__device__ __forceinline__ double dg_np_sum(const double *v, int n, int stride) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += v[i * stride];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = v[j * stride];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += v[(i + j) * stride];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += v[i * stride];
    return res;
}

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) ----------------------------------
__device__ __forceinline__ void dg_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 53-bit uniform in (0,1)
__device__ __forceinline__ double dg_u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double dg_normal(uint64_t seed, int64_t sid, uint32_t step, uint32_t elem) {
    uint32_t o[4];
    dg_philox(step, elem, (uint32_t)sid, (uint32_t)((uint64_t)sid >> 32),
              (uint32_t)seed, (uint32_t)(seed >> 32), o);
    double u1 = dg_u53(o[0], o[1]), u2 = dg_u53(o[2], o[3]);
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    (void)s;
    return sqrt(-2.0 * log(u1)) * c;
}
// two normals from one Philox block (Box-Muller, sine and cosine branch)
__device__ __forceinline__ void dg_normal_pair(uint64_t seed, int64_t sid, uint32_t step, uint32_t pair,
                                               double &z0, double &z1) {
    uint32_t o[4];
    dg_philox(step, pair, (uint32_t)sid, (uint32_t)((uint64_t)sid >> 32),
              (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double u1 = dg_u53(o[0], o[1]), u2 = dg_u53(o[2], o[3]);
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    const double r = sqrt(-2.0 * log(u1));
    z0 = r * c;
    z1 = r * s;
}
__device__ __forceinline__ double dg_uniform(uint64_t seed, int64_t sid, uint32_t step) {
    uint32_t o[4];
    dg_philox(step, 0xFFFFFFFFu, (uint32_t)sid, (uint32_t)((uint64_t)sid >> 32),
              (uint32_t)seed, (uint32_t)(seed >> 32), o);
    // [0,1) like numpy's random_sample
    return ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
}

// ---- mantissa/exponent product accumulator ------------------------------------------------
// log(prod x_n) = log(prod mant_n) + ln2 * sum exp_n.  One DMUL + a few integer ops per
// read instead of a ~30-instruction DP log; the relative error of the running product
// grows by 2^-53 per factor, i.e. the summed log is MORE accurate than adding rounded
// logs.  Anything that is not a positive normal double (0, denormal, negative, inf, NaN)
// raises `bad` and the step is redone with real logs (slow path) so that -inf / NaN
// propagate exactly like np.log(...).sum() (model_GP.py:344).
struct MantExp {
    double m;
    int e;
    unsigned bad;
};
__device__ __forceinline__ void dg_acc(MantExp &a, double x) {
    int hi = __double2hiint(x);
    a.bad |= ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) ? 1u : 0u;
    a.e += (hi >> 20) - 1023;
    a.m *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
}
__device__ __forceinline__ void dg_renorm(MantExp &a) {
    int hi = __double2hiint(a.m);
    a.e += (hi >> 20) - 1023;
    a.m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(a.m));
}

#endif  // __CUDACC__
