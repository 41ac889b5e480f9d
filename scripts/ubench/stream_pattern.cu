// Microbenchmark: bandwidth of the chain kernel's W streaming pattern (per-warp cp.async ring),
// without the MH arithmetic.  Each block owns one "gene" of C rows x NP double2 per time point,
// T time points; warps take 32-pair items round-robin; every pass re-reads the gene.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void cp16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int C, int NS>
__global__ void __launch_bounds__(512, 1) stream_kernel(const double2 *W, int T, int np, int passes, int nwarps_used, double *out, int genes) {
    extern __shared__ double2 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    if (warp < nwarps_used) {
        double2 *st = smem + (size_t)warp * NS * C * 32 + lane;
        const int items_per_t = np / 32, nitems = T * items_per_t;
        for (int gsel = blockIdx.x; gsel < genes; gsel += gridDim.x) {
            const double2 *Wg = W + (size_t)gsel * T * C * np;
            for (int pass = 0; pass < passes; ++pass) {
                int qi = warp;
                auto issue = [&](int slot) {
                    if (qi < nitems) {
                        const int t = qi / items_per_t, ch = qi - t * items_per_t;
                        const double2 *src = Wg + (size_t)t * C * np + ch * 32 + lane;
#pragma unroll
                        for (int k = 0; k < C; ++k) cp16(st + (slot * C + k) * 32, src + (size_t)k * np);
                        qi += nwarps_used;
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
#pragma unroll
                for (int s = 0; s < NS - 1; ++s) issue(s);
                int it = 0;
                for (int q = warp; q < nitems; q += nwarps_used, ++it) {
                    issue((it + NS - 1) % NS);
                    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
                    const double2 *s2 = st + ((it % NS) * C) * 32;
#pragma unroll
                    for (int k = 0; k < C; ++k) { double2 w = s2[k * 32]; acc += w.x * 1.0000001 + w.y; }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int C, int NS>
void run(int blocks, int threads, int nwarps_used, int genes, const double2 *W, double *out, int T, int np, int passes) {
    size_t smem = (size_t)nwarps_used * NS * C * 32 * sizeof(double2);
    cudaFuncSetAttribute(stream_kernel<C, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    stream_kernel<C, NS><<<blocks, threads, smem>>>(W, T, np, 1, nwarps_used, out, genes);
    cudaEventRecord(e0);
    stream_kernel<C, NS><<<blocks, threads, smem>>>(W, T, np, passes, nwarps_used, out, genes);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)genes * passes * T * C * np * 16.0;
    printf("C=%d NS=%d blocks=%d thr=%d warps=%d genes=%d (%.0f MB): %.2f ms  %.0f GB/s  (%s)\n", C, NS, blocks, threads, nwarps_used, genes,
           genes * (double)T * C * np * 16 / 1e6, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int T = 8, np = 512, C = 8;                    // 512 KB per gene
    const int genes_max = 1184;
    size_t n = (size_t)genes_max * T * C * np;
    double2 *W; double *out;
    cudaMalloc(&W, n * sizeof(double2)); cudaMemset(W, 0, n * sizeof(double2)); cudaMalloc(&out, 8);
    // L2-resident working set (148 genes = 78 MB) and HBM (592 genes = 310 MB)
    for (int genes : {148, 592}) {
        int passes = genes == 148 ? 40 : 10;
        run<8, 4>(genes <= 148 ? 148 : 296, 256, 6, genes, W, out, T, np, passes);
        run<8, 4>(genes <= 148 ? 148 : 296, 512, 14, genes, W, out, T, np, passes);
        run<8, 2>(genes <= 148 ? 148 : 296, 512, 14, genes, W, out, T, np, passes);
        run<8, 8>(genes <= 148 ? 148 : 148, 256, 6, genes, W, out, T, np, passes);
        run<8, 4>(genes <= 148 ? 148 : 592, 128, 3, genes, W, out, T, np, passes);
    }
    return 0;
}
