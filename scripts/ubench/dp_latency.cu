// Microbenchmark: fp64 dependent-chain latency and per-SM throughput on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dp_latency dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 4096

template <int OP>
__device__ __forceinline__ double op(double x, double a, double b) {
    if (OP == 0) return fma(x, a, b);
    if (OP == 1) return x * a;
    if (OP == 2) return x + b;
    if (OP == 3) return a / x + b;              // division
    if (OP == 4) return exp(x * 1e-3) ;         // exp (argument stays small)
    if (OP == 5) return log(x + 2.0);
    if (OP == 6) return sqrt(x + 2.0);
    if (OP == 7) return 1.0 / x + b;
    if (OP == 8) return __drcp_rn(x) + b;
    return x;
}

// latency: one warp, one dependent chain
template <int OP>
__global__ void lat_kernel(double *out, long long *cyc, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT; ++i) x = op<OP>(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// throughput: many warps, 8 independent chains per thread
template <int OP>
__global__ void thr_kernel(double *out, long long *cyc, double a, double b) {
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = out[threadIdx.x] + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = op<OP>(x[j], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char *name, double a, double b, double *d_out, long long *d_cyc) {
    long long h;
    lat_kernel<OP><<<1, 32>>>(d_out, d_cyc, a, b);
    cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    double lat = (double)h / N_IT;
    thr_kernel<OP><<<148, 1024>>>(d_out, d_cyc, a, b);
    cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    // ops per SM per cycle: 1024 threads * N_IT ops / cycles
    double thr = 1024.0 * N_IT / (double)h;
    printf("%-10s latency %7.1f cyc   throughput %6.1f lane-ops/clk/SM\n", name, lat, thr);
}

int main() {
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 8);
    cudaMemset(d_out, 0, 148 * 1024 * 8);
    cudaMalloc(&d_cyc, 8);
    run<0>("dfma", 0.999, 1e-3, d_out, d_cyc);
    run<1>("dmul", 0.999999, 0, d_out, d_cyc);
    run<2>("dadd", 0, 1e-3, d_out, d_cyc);
    run<3>("a/x+b", 1.5, 0.7, d_out, d_cyc);
    run<7>("1/x+b", 1.5, 0.7, d_out, d_cyc);
    run<8>("drcp+b", 1.5, 0.7, d_out, d_cyc);
    run<4>("exp", 0, 0, d_out, d_cyc);
    run<5>("log", 0, 0, d_out, d_cyc);
    run<6>("sqrt", 0, 0, d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("clock rate attr %d kHz\n", clk);
    return 0;
}
