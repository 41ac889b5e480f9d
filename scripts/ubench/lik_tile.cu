// Microbenchmark for the likelihood tile of the wide-round kernel: which (reads per lane x candidates per lane)
// register tile keeps the fp64 pipe busy when W and f come from shared memory?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lik_tile lik_tile.cu
// Part 1: cycles per LDS.128 / LDS.64 for address patterns (uniform, uniform per quarter warp, 8 words
//         repeated over the quarters, 32 consecutive words).
// Part 2: DFMA lane-ops per clock per SM of a C = 8 tile loop for several tile shapes and warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK 8

template <int PAT, int WIDTH>
__global__ void lds_kernel(double *out, long long *cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int w;
    if (PAT == 0) w = 0;                       // one address for the whole warp
    else if (PAT == 1) w = lane >> 3;          // one address per quarter warp
    else if (PAT == 2) w = lane & 7;           // 8 consecutive words, the same in every quarter
    else if (PAT == 3) w = lane;               // 32 consecutive words
    else w = lane >> 4;                        // one address per half warp
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + w * WIDTH;
    double a0 = 0, a1 = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double x, y = 0;
            if (WIDTH == 16) asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(base + u * 512));
            else asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(x) : "r"(base + u * 512));
            a0 += x; a1 += y;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// tile: WPL words (16 B = 2 reads) per lane, NJ candidates per lane; LANEMAP 0: every lane its own words and all
// NJ candidates (f uniform); 1: lane = (s = lane & 7, h = lane >> 3), words s, s + 8, candidates h*NJ.. (f per quarter)
template <int WPL, int NJ, int LANEMAP>
__global__ void tile_kernel(double *out, long long *cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words_per_task = LANEMAP == 0 ? 32 * WPL : 16;
    const int kstride = nwarps * words_per_task + 8;             // double2 per isoform row
    double2 *W2 = reinterpret_cast<double2 *>(sm);
    double *ff = sm + 2 * CK * kstride;
    const int JT = LANEMAP == 0 ? NJ : 4 * NJ;
    for (int i = threadIdx.x; i < 2 * CK * kstride; i += blockDim.x) sm[i] = 1e-3 + 1e-6 * i;
    for (int i = threadIdx.x; i < 2 * CK * JT; i += blockDim.x) ff[i] = 0.1 + 1e-3 * i;
    __syncthreads();
    const double2 *wt0 = W2 + warp * words_per_task + (LANEMAP == 0 ? lane : (lane & 7));
    const double2 *fp0 = reinterpret_cast<const double2 *>(ff + (LANEMAP == 0 ? 0 : (lane >> 3) * NJ));
    double mm[NJ];
#pragma unroll
    for (int n = 0; n < NJ; ++n) mm[n] = 1.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const double2 *wt = wt0 + (it & 3) * 2;                  // iteration-dependent address: nothing can be hoisted
        const double2 *fp = fp0 + (it & 1) * (CK * JT) / 2;
        double x[2 * WPL][NJ];
#pragma unroll
        for (int r = 0; r < 2 * WPL; ++r)
#pragma unroll
            for (int n = 0; n < NJ; ++n) x[r][n] = 0.0;
#pragma unroll
        for (int k = 0; k < CK; ++k) {
            double2 w[WPL];
#pragma unroll
            for (int u = 0; u < WPL; ++u) w[u] = wt[k * kstride + (LANEMAP == 0 ? 32 * u : 8 * u)];
            double f[NJ];
#pragma unroll
            for (int u = 0; u < NJ / 2; ++u) { const double2 t = fp[(k * JT) / 2 + u]; f[2 * u] = t.x; f[2 * u + 1] = t.y; }
#pragma unroll
            for (int u = 0; u < WPL; ++u)
#pragma unroll
                for (int n = 0; n < NJ; ++n) {
                    x[2 * u][n] = fma(w[u].x, f[n], x[2 * u][n]);
                    x[2 * u + 1][n] = fma(w[u].y, f[n], x[2 * u + 1][n]);
                }
        }
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
            double pr = x[0][n] * x[1][n];
#pragma unroll
            for (int u = 1; u < WPL; ++u) pr *= x[2 * u][n] * x[2 * u + 1][n];
            const int h = __double2hiint(pr);
            mm[n] *= __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(pr));
            if ((it & 63) == 63) { const int g = __double2hiint(mm[n]); mm[n] = __hiloint2double((g & 0x000fffff) | 0x3ff00000, __double2loint(mm[n])); }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int n = 0; n < NJ; ++n) s += mm[n];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int PAT, int WIDTH>
void run_lds(const char *name, double *d_out, long long *d_cyc) {
    long long h;
    for (int warps : {1, 4, 16}) {
        lds_kernel<PAT, WIDTH><<<148, warps * 32, 65536>>>(d_out, d_cyc, 256);
        cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("LDS.%d %-28s warps/SM %2d: %.2f cycles per warp instruction, %.2f cycles per instruction per SM\n", WIDTH * 8, name, warps,
               (double)h / (256 * 16), (double)h / (256 * 16) / warps);
    }
}

template <int WPL, int NJ, int LANEMAP>
void run_tile(const char *name, double *d_out, long long *d_cyc) {
    long long h;
    cudaFuncSetAttribute(tile_kernel<WPL, NJ, LANEMAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int warps : {4, 8, 12, 16}) {
        const int wpt = LANEMAP == 0 ? 32 * WPL : 16;
        const size_t smem = (size_t)(2 * CK * (warps * wpt + 8) + 2 * CK * 4 * NJ + 64) * 8;
        const int iters = 2000;
        tile_kernel<WPL, NJ, LANEMAP><<<148, warps * 32, smem>>>(d_out, d_cyc, iters);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        const double dfma = (double)iters * CK * 2 * WPL * NJ * 32 * warps;      // lane-ops per SM
        printf("tile %-44s warps/SM %2d: %6.1f DFMA lane-ops/clk/SM (peak 64)\n", name, warps, dfma / (double)h);
    }
}

int main() {
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 8);
    cudaMalloc(&d_cyc, 64);
    cudaFuncSetAttribute(lds_kernel<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(lds_kernel<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    run_lds<0, 16>("uniform", d_out, d_cyc);
    run_lds<4, 16>("uniform per half warp", d_out, d_cyc);
    run_lds<1, 16>("uniform per quarter warp", d_out, d_cyc);
    run_lds<2, 16>("8 words, same in each quarter", d_out, d_cyc);
    run_lds<3, 16>("32 consecutive words", d_out, d_cyc);
    run_lds<0, 8>("uniform", d_out, d_cyc);
    run_lds<3, 8>("32 consecutive doubles", d_out, d_cyc);
    run_tile<2, 4, 1>("(s8,h): 2 words x 4 cand, 16 cand per warp", d_out, d_cyc);
    run_tile<1, 8, 0>("1 word x 8 cand", d_out, d_cyc);
    run_tile<1, 16, 0>("1 word x 16 cand", d_out, d_cyc);
    run_tile<2, 8, 0>("2 words x 8 cand", d_out, d_cyc);
    run_tile<2, 4, 0>("2 words x 4 cand", d_out, d_cyc);
    run_tile<4, 4, 0>("4 words x 4 cand", d_out, d_cyc);
    run_tile<2, 16, 0>("2 words x 16 cand", d_out, d_cyc);
    return 0;
}
