// Microbenchmark: read bandwidth of the L2 (working set 32 / 64 / 96 MB, re-read many times) and of HBM
// (4 GB, read once per pass) with 16-byte loads from every SM -- the L2 / HBM denominators of bench.py's roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_bw l2_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void read_kernel(const double2 *__restrict__ buf, size_t n2, int passes, double *out) {
    double a = 0.0, b = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += 4 * stride) {
            double2 v0 = __ldcg(buf + i), v1 = i + stride < n2 ? __ldcg(buf + i + stride) : make_double2(0, 0);
            double2 v2 = i + 2 * stride < n2 ? __ldcg(buf + i + 2 * stride) : make_double2(0, 0);
            double2 v3 = i + 3 * stride < n2 ? __ldcg(buf + i + 3 * stride) : make_double2(0, 0);
            a += v0.x + v1.x + v2.x + v3.x;
            b += v0.y + v1.y + v2.y + v3.y;
        }
    }
    if (a + b == 123.456) out[0] = a;
}

int main() {
    double2 *buf; double *out;
    const size_t big = (size_t)4 << 30;
    cudaMalloc(&buf, big);
    cudaMalloc(&out, 8);
    cudaMemset(buf, 0, big);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t sizes[4] = {(size_t)32 << 20, (size_t)64 << 20, (size_t)96 << 20, big};
    for (int k = 0; k < 4; ++k) {
        const size_t n2 = sizes[k] / 16;
        const int passes = k < 3 ? 200 : 4;
        read_kernel<<<148 * 4, 512>>>(buf, n2, 2, out);                  // warm the L2
        cudaEventRecord(e0);
        read_kernel<<<148 * 4, 512>>>(buf, n2, passes, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("read %5zu MB x %3d passes: %.1f GB/s\n", sizes[k] >> 20, passes, (double)sizes[k] * passes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
