// Standalone probe of the tcgen05 conventions logprob_tc.cu relies on (run on a B200: `nvcc ... && ./umma_probe`):
//   * no-swizzle K-major shared-memory operand layout  [k/8][row][8 elems]  (LBO = rows*16 B, SBO = 128 B)
//   * kind::f16 instruction descriptor (bf16 x bf16 -> fp32), M = 128, N in {16, 48, 128}
//   * accumulation over several K=16 MMAs by advancing the descriptor start address
//   * cp.async.bulk global->shared with mbarrier complete_tx, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b
// Prints max |error| against a CPU product; exits non-zero on mismatch.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int N, int KTOT>
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ Apacked,
                                                    const __nv_bfloat16* __restrict__ Bpacked, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);                       // [KTOT/8][128][8]
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + KTOT * 128 * 2);      // [KTOT/8][N][8]
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    constexpr uint32_t A_BYTES = KTOT * 128 * 2, B_BYTES = KTOT * N * 2;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_load)), "r"(A_BYTES + B_BYTES));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sA)), "l"(Apacked), "r"(A_BYTES), "r"(smem_u32(&bar_load)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(Bpacked), "r"(B_BYTES), "r"(smem_u32(&bar_load)) : "memory");
        mbar_wait(&bar_load, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        // instruction descriptor: D fp32, A/B bf16, both K-major, N, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int kk = 0; kk < KTOT / 16; ++kk) {
            const uint64_t da = make_desc(smem_u32(sA) + kk * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t db = make_desc(smem_u32(sB) + kk * 2 * (N * 16), N * 16, 128);
            const uint32_t acc = kk > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // each warp reads its 32 lanes: row = 32*warp + lane, N columns in groups of 16
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

template <int N, int KTOT>
static int run_case() {
    std::vector<float> A(128 * KTOT), B(N * KTOT);
    srand(1234 + N + KTOT);
    for (auto& v : A) v = (float)((rand() % 9) - 4);
    for (auto& v : B) v = (float)((rand() % 7) - 3) * 0.5f;
    std::vector<__nv_bfloat16> Ap(128 * KTOT), Bp(N * KTOT);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < KTOT; ++k) Ap[(size_t)(k / 8) * 128 * 8 + r * 8 + (k % 8)] = __float2bfloat16(A[r * KTOT + k]);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < KTOT; ++k) Bp[(size_t)(k / 8) * N * 8 + n * 8 + (k % 8)] = __float2bfloat16(B[n * KTOT + k]);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, Ap.size() * 2));
    CK(cudaMalloc(&dB, Bp.size() * 2));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
    size_t smem = (size_t)KTOT * (128 + N) * 2 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<N, KTOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N, KTOT><<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KTOT; ++k) ref += (double)A[r * KTOT + k] * B[n * KTOT + k];
            maxerr = std::max(maxerr, std::abs(ref - D[r * N + n]));
        }
    printf("umma probe N=%d K=%d: max |err| = %g %s\n", N, KTOT, maxerr, maxerr == 0 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr == 0 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run_case<128, 16>();
    bad += run_case<128, 64>();
    bad += run_case<16, 64>();
    bad += run_case<48, 256>();
    bad += run_case<128, 256>();
    return bad;
}
