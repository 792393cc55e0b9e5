// Standalone probe of the tcgen05 conventions train_tc.cu relies on in addition to tests/umma_probe.cu
// (run on a B200: `nvcc -gencode arch=compute_100a,code=sm_100a tests/umma_probe_mn.cu -o build/umma_probe_mn`):
//   * MN-major (transposed) no-swizzle operands: the SAME shared-memory image  [feature/8][row][8 elems]  read as
//       K-major  (contraction over the features):  LBO = rows*16 B (k-chunk stride), SBO = 128 B (8-row group)
//       MN-major (contraction over the rows):      SBO = rows*16 B (feature-chunk stride), LBO = 128 B (8-row group),
//                                                  instruction-descriptor bits 15 (A) / 16 (B) set
//   * N = 256 in one instruction (M = 128), fp16 operands
// Prints max |error| against a CPU product; exits non-zero on mismatch.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

struct Op { uint32_t bytes, lbo, sbo, step; };  // image size, descriptor strides, start-address advance per K=16

__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ Aimg, const __half* __restrict__ Bimg,
                                                    float* __restrict__ D, Op oa, Op ob, int N, int ksteps, uint32_t idesc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + oa.bytes;
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
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_load)), "r"(oa.bytes + ob.bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sA)), "l"(Aimg), "r"(oa.bytes), "r"(smem_u32(&bar_load)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(Bimg), "r"(ob.bytes), "r"(smem_u32(&bar_load)) : "memory");
        mbar_wait(&bar_load, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t da = make_desc(smem_u32(sA) + kk * oa.step, oa.lbo, oa.sbo);
            const uint64_t db = make_desc(smem_u32(sB) + kk * ob.step, ob.lbo, ob.sbo);
            const uint32_t acc = kk > 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
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

// image [feature/8][rows][8] of a [rows][features] matrix
static std::vector<__half> image(const std::vector<float>& X, int rows, int feats) {
    std::vector<__half> img((size_t)rows * feats);
    for (int r = 0; r < rows; ++r)
        for (int f = 0; f < feats; ++f) img[(size_t)(f / 8) * rows * 8 + r * 8 + (f % 8)] = __float2half(X[(size_t)r * feats + f]);
    return img;
}

// mode 0: D[128][N] = A[128][K] * B[N][K]^T, both K-major (contraction over the features of both images)
// mode 1: D[128][N] = P[R][128]^T * Q[R][N], both MN-major (contraction over the R rows of both images)
static int run_case(int mode, int N, int KorR) {
    const int M = 128;
    std::vector<float> A, B;
    std::vector<__half> Ai, Bi;
    Op oa{}, ob{};
    int ksteps;
    srand(99 + mode * 7 + N + KorR);
    auto rnd = [](float s) { return (float)((rand() % 9) - 4) * s; };
    if (mode == 0) {
        const int K = KorR;
        A.resize((size_t)M * K); B.resize((size_t)N * K);
        for (auto& v : A) v = rnd(1.f);
        for (auto& v : B) v = rnd(0.5f);
        Ai = image(A, M, K); Bi = image(B, N, K);
        oa = {(uint32_t)(M * K * 2), (uint32_t)(M * 16), 128u, (uint32_t)(2 * M * 16)};
        ob = {(uint32_t)(N * K * 2), (uint32_t)(N * 16), 128u, (uint32_t)(2 * N * 16)};
        ksteps = K / 16;
    } else {
        const int R = KorR;
        A.resize((size_t)R * M); B.resize((size_t)R * N);   // P [R][M], Q [R][N]
        for (auto& v : A) v = rnd(1.f);
        for (auto& v : B) v = rnd(0.5f);
        Ai = image(A, R, M); Bi = image(B, R, N);
        oa = {(uint32_t)(R * M * 2), 128u, (uint32_t)(R * 16), 256u};   // LBO = 8-row group, SBO = feature-chunk stride
        ob = {(uint32_t)(R * N * 2), 128u, (uint32_t)(R * 16), 256u};
        ksteps = R / 16;
    }
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 x fp16 -> fp32
    if (mode == 1) idesc |= (1u << 15) | (1u << 16);
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, Ai.size() * 2)); CK(cudaMalloc(&dB, Bi.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, Ai.data(), Ai.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bi.data(), Bi.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = oa.bytes + ob.bytes + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, oa, ob, N, ksteps, idesc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            if (mode == 0) for (int k = 0; k < KorR; ++k) ref += (double)A[(size_t)m * KorR + k] * B[(size_t)n * KorR + k];
            else for (int r = 0; r < KorR; ++r) ref += (double)A[(size_t)r * M + m] * B[(size_t)r * N + n];
            maxerr = std::max(maxerr, std::abs(ref - D[(size_t)m * N + n]));
        }
    printf("umma probe %s N=%d %s=%d: max |err| = %g %s\n", mode ? "MN-major" : "K-major ", N, mode ? "R" : "K", KorR, maxerr,
           maxerr == 0 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr == 0 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run_case(0, 256, 64);
    bad += run_case(0, 256, 256);
    bad += run_case(1, 128, 16);
    bad += run_case(1, 128, 128);
    bad += run_case(1, 256, 128);
    bad += run_case(1, 16, 128);
    bad += run_case(1, 48, 128);
    return bad;
}
