// Do SHFL and LDS share one throughput limit?  Time N LDS.64, N x (2 SHFL.32), and both together,
// all warps of a full SM busy.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: LDS only, 1: SHFL only, 2: both
__global__ void k(double *out, int iters) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double acc = 0.0, v = threadIdx.x;
    const int lane = threadIdx.x & 31;
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) { acc += sm[(idx + u * 32) & 2047]; }
            if (MODE == 1 || MODE == 2) { v = __shfl_sync(0xffffffffu, v, (lane + 1 + u) & 31); acc += v; }
        }
        idx += 17;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

int main() {
    double *d; cudaMalloc(&d, 1 << 22);
    double h; const int it = 2048;
    for (int warps : {8, 16, 32}) {
        k<0><<<1, warps * 32>>>(d, it); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); double a = h / (it * 8.0 * warps);
        k<1><<<1, warps * 32>>>(d, it); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); double b = h / (it * 8.0 * warps);
        k<2><<<1, warps * 32>>>(d, it); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); double c = h / (it * 8.0 * warps);
        printf("warps=%2d cycles per warp-op per SM: LDS.64 %.2f | double shuffle (2 SHFL.32) %.2f | both %.2f (sum would be %.2f)\n",
               warps, a, b, c, a + b);
    }
    return 0;
}
