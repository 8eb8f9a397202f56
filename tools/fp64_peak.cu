// Measured FP64 FMA throughput of the device (the compute roof of the S-build pass): 8 independent DFMA chains per
// thread, 1024 threads per SM-sized block set.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_peak.cu -o /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<<<blocks, threads>>>(out, 1000, 0.999999, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 8.0 * iters * blocks * threads;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm\": %.2f}\n",
           p.name, p.multiProcessorCount, p.clockRate / 1000, fma / (ms * 1e-3), 2 * fma / (ms * 1e-3) / 1e12,
           fma / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
    return 0;
}
