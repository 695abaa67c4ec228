// Micro-benchmarks of what bounds the sweep kernels on B200: DFMA throughput and
// latency, LDS.64 latency, bar.red cost.  nvcc -arch=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_tput(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

__global__ void dfma_lat(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b);
        x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x + 1] = x0;
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0);
}

__global__ void ffma_lat(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x0 = fmaf(x0, a, b); x0 = fmaf(x0, a, b); x0 = fmaf(x0, a, b);
        x0 = fmaf(x0, a, b); x0 = fmaf(x0, a, b); x0 = fmaf(x0, a, b); x0 = fmaf(x0, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x + 1] = x0;
    if (threadIdx.x == 0) out[0] = (float)(t1 - t0);
}

__global__ void lds_lat(double *out, int iters) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { p = idx[p]; p = idx[p]; p = idx[p]; p = idx[p]; }
    long long t1 = clock64();
    out[threadIdx.x + 1] = p;
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0);
}

__global__ void bar_cost(double *out, int iters) {
    int v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { v = __syncthreads_or(v > i); }
    long long t1 = clock64();
    out[threadIdx.x + 1] = v;
    if (threadIdx.x == 0) out[0] = (double)(t1 - t0);
}

int main() {
    double *d; cudaMalloc(&d, 1 << 24);
    double h;
    const int it = 4096;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        dfma_tput<<<1, warps * 32>>>(d, it, 1.0000001, 1e-9);
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("dfma tput  warps=%2d: %.2f cycles per warp-DFMA per SM  (%.1f lanes/clk/SM)\n", warps,
               h / (it * 8.0 * warps), 32.0 * it * 8.0 * warps / h);
    }
    dfma_lat<<<1, 32>>>(d, it, 1.0000001, 1e-9);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("dfma dependent latency: %.2f cycles\n", h / (it * 8.0));
    float hf; ffma_lat<<<1, 32>>>((float *)d, it, 1.0000001f, 1e-9f);
    cudaMemcpy(&hf, d, 4, cudaMemcpyDeviceToHost);
    printf("ffma dependent latency: %.2f cycles\n", hf / (it * 8.0));
    lds_lat<<<1, 32>>>(d, it);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("lds dependent latency: %.2f cycles\n", h / (it * 4.0));
    for (int warps : {1, 4, 8, 16, 32}) {
        bar_cost<<<1, warps * 32>>>(d, it);
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("bar.red.or warps=%2d: %.1f cycles per barrier\n", warps, h / it);
    }
    // full-chip DFMA throughput
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_tput<<<148 * 2, 1024>>>(d, it, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_tput<<<148 * 2, 1024>>>(d, it * 4, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("chip DFMA: %.2f TFLOP/s\n", 2.0 * 148 * 2 * 1024 * 8.0 * it * 4 / (ms * 1e-3) / 1e12);
    return 0;
}
