// Micro-benchmark of tcgen05.mma issue/commit cost on one SM (timing experiment, not a test).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// mode 0: 1 thread, back-to-back; mode 1: commit every `grp` MMAs to rotating barriers; nthr issuing warps
template <int N>
__global__ void __launch_bounds__(128) bench(int total, int grp, int nthr, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ __align__(8) uint64_t done[4];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + N) * 64 * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1 << 20); for (int i = 0; i < 4; ++i) mbar_init(&done[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (lane == 0 && warp < nthr) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFF) | ((128u * 16 >> 4) << 16);
        const uint32_t b0 = (((smem_u32(smem) + 128 * 64 * 2) >> 4) & 0x3FFF) | ((uint32_t)N << 16);
        const uint32_t d = tmem + warp * (N <= 128 ? 128 : 256);
        long long t0 = clock64();
        int bi = 0;
        for (int i = 0; i < total; i += grp) {
            for (int u = 0; u < grp; ++u) {
                const int kk = u & 3;
                umma(d, ((uint64_t)hi << 32) | (a0 + kk * 256), ((uint64_t)hi << 32) | (b0 + kk * 2 * N), idesc, (i | u) ? 1u : 0u);
            }
            if (grp < total) { commit(&bars[bi]); bi = (bi + 1) & 7; }
        }
        long long t1 = clock64();
        commit(&done[warp]);
        long long t2 = clock64();
        mbar_wait(&done[warp], 0);
        long long t3 = clock64();
        out[warp * 4 + 0] = t1 - t0; out[warp * 4 + 1] = t2 - t1; out[warp * 4 + 2] = t3 - t2; out[warp * 4 + 3] = t3 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
template <int N> void run(int total, int grp, int nthr) {
    long long* d; CK(cudaMalloc(&d, 16 * 8)); CK(cudaMemset(d, 0, 128));
    size_t smem = (128 + N) * 64 * 2 + 1024;
    CK(cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) { bench<N><<<1, 128, smem>>>(total, grp, nthr, d); CK(cudaDeviceSynchronize()); }
    long long h[16]; CK(cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost));
    for (int w = 0; w < nthr; ++w)
        printf("N=%3d total=%4d grp=%4d thr=%d/%d: issue %6lld cyc (%.1f/mma)  final commit %4lld  drain %6lld  all %6lld (%.1f/mma)\n", N, total, grp, w, nthr, h[w*4], (double)h[w*4] / total, h[w*4+1], h[w*4+2], h[w*4+3], (double)h[w*4+3] / total);
    cudaFree(d);
}
int main() {
    run<128>(256, 256, 1); run<128>(256, 4, 1); run<128>(256, 8, 1); run<128>(256, 1, 1);
    run<256>(256, 256, 1); run<256>(256, 4, 1);
    run<128>(256, 256, 2); run<128>(256, 4, 2); run<128>(128, 4, 4);
    run<64>(256, 256, 1); run<16>(256, 256, 1);
    return 0;
}
