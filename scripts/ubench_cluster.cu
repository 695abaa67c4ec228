// Ping-pong latency between two CTAs of a thread-block cluster on sm_100a:
//   mode 0: st.async (remote store + mbarrier complete_tx) / mbarrier.try_wait
//   mode 1: plain remote store (st.volatile.shared::cluster) / polling ld.volatile.shared
//   mode 2: st.relaxed.cluster remote store / polling ld.relaxed.cluster
//   mode 3: barrier.cluster arrive.release + wait.acquire (all CTAs), per round
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_cluster scripts/ubench_cluster.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    return ok;
}

__global__ void pingpong(int mode, int rounds, int peer_dist, long long *out) {
    __shared__ __align__(16) unsigned long long bar[2];
    __shared__ __align__(16) uint32_t word[4];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = cl.block_rank(), n = cl.num_blocks();
    const uint32_t sb = smem_u32(bar), sw = smem_u32(word);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb + 8));
        asm volatile("fence.mbarrier_init.release.cluster;");
        word[0] = word[1] = word[2] = word[3] = 0;
    }
    cl.sync();
    const int peer = rank == 0 ? peer_dist : 0;
    const bool active = (rank == 0 || rank == peer_dist) && threadIdx.x == 0;
    const uint32_t rbar = mapa_u32(sb, peer), rword = mapa_u32(sw, peer);
    long long t0 = clock64();
    if (mode == 3) {
        for (int i = 0; i < rounds; ++i) {
            asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    } else if (active) {
        // rank 0 sends round i, waits for the echo; the peer waits, then echoes
        for (int i = 0; i < rounds; ++i) {
            const uint32_t par = i & 1;
            if (mode == 0) {
                if (rank == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(sb) : "memory");
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(rword), "r"(i + 1), "r"(rbar) : "memory");
                    while (!try_wait(sb, par)) {}
                } else {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(sb) : "memory");
                    while (!try_wait(sb, par)) {}
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(rword), "r"(i + 1), "r"(rbar) : "memory");
                }
            } else {
                uint32_t v;
                if (rank == 0) {
                    if (mode == 1) asm volatile("st.volatile.shared::cluster.u32 [%0], %1;" ::"r"(rword), "r"(i + 1) : "memory");
                    else asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(rword), "r"(i + 1) : "memory");
                    do {
                        if (mode == 1) asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(sw) : "memory");
                        else asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(sw) : "memory");
                    } while (v != (uint32_t)(i + 1));
                } else {
                    do {
                        if (mode == 1) asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(sw) : "memory");
                        else asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(sw) : "memory");
                    } while (v != (uint32_t)(i + 1));
                    if (mode == 1) asm volatile("st.volatile.shared::cluster.u32 [%0], %1;" ::"r"(rword), "r"(i + 1) : "memory");
                    else asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(rword), "r"(i + 1) : "memory");
                }
            }
        }
    }
    long long t1 = clock64();
    if (rank == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    cl.sync();
}

// dependent-load latency of the three flavours of shared-memory load used for polling
__global__ void ldlat(int flavour, int iters, long long *out) {
    __shared__ uint32_t w[32];
    w[threadIdx.x & 31] = 0;
    __syncthreads();
    uint32_t a = smem_u32(w), v = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (flavour == 0) asm volatile("ld.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a + v) : "memory");
        else if (flavour == 1) asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a + v) : "memory");
        else if (flavour == 2) asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a + v) : "memory");
        else asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a + v) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0 + v;
}

int main() {
    {
        long long *d0; cudaMalloc(&d0, 8);
        const char *names[4] = {"ld.shared (weak)", "ld.volatile.shared", "ld.relaxed.cluster.shared", "ld.relaxed.cta.shared"};
        for (int f = 0; f < 4; ++f) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) { ldlat<<<1, 32>>>(f, 10000, d0); cudaDeviceSynchronize(); cudaMemcpy(&h, d0, 8, cudaMemcpyDeviceToHost); }
            printf("%-28s %6.1f cycles per dependent load\n", names[f], (double)h / 10000);
        }
    }
    long long *d; cudaMalloc(&d, 8);
    const int rounds = 20000;
    cudaFuncSetAttribute(pingpong, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int csize : {2, 8, 16}) for (int threads : {32, 256}) for (int mode = 0; mode < 4; ++mode) {
        for (int dist : {1, csize - 1}) {
            if (dist != 1 && (csize == 2 || mode == 3)) continue;
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(csize); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaError_t e = cudaLaunchKernelEx(&cfg, pingpong, mode, rounds, dist, d);
                if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); break; }
                e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sync failed: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            }
            printf("cluster %2d threads %3d mode %d peer +%2d: %7.1f cycles per %s\n", csize, threads, mode, dist,
                   (double)h / rounds, mode == 3 ? "barrier.cluster" : "round trip (2 flights)");
        }
    }
    return 0;
}
