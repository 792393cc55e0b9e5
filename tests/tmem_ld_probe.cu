// Probe of tcgen05.ld fragment layouts (run on a B200).  TMEM is filled through the 32x32b shape (thread t of warp w owns
// lane 32*(w%4)+t; register j = column j) with value = lane*1000 + column; then one warp reads 16 lanes back with the
// 16x256b / 16x128b / 16x64b shapes and prints which (lane, column) every thread received in every register.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    // fill 64 columns: lane L, column c -> L*1000 + c
    {
        const uint32_t L = warp * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 64; c0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = L * 1000 + c0 + j;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c0), "r"(v[0]),
                         "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 1) {   // lanes 32..63; read the 16 lanes starting at lane 32 and at lane 48
        for (int sub = 0; sub < 2; ++sub) {
            const uint32_t taddr = tmem + ((uint32_t)(32 + 16 * sub) << 16);
            uint32_t r[8];
            // 16x256b.x2: 16 lanes x 16 columns, 8 registers per thread
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) out[(0 * 2 + sub) * 256 + lane * 8 + j] = r[j];
            uint32_t q[4];
            // 16x128b.x2: 16 lanes x 8 columns, 4 registers per thread
            asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 4; ++j) out[(1 * 2 + sub) * 256 + lane * 8 + j] = q[j];
            uint32_t u[4];
            // 16x64b.x4: 16 lanes x 8 columns, 4 registers per thread
            asm volatile("tcgen05.ld.sync.aligned.16x64b.x4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 4; ++j) out[(2 * 2 + sub) * 256 + lane * 8 + j] = u[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

int main() {
    uint32_t* d;
    CK(cudaMalloc(&d, 6 * 256 * 4));
    CK(cudaMemset(d, 0xFF, 6 * 256 * 4));
    probe<<<1, 128>>>(d);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    uint32_t h[6 * 256];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    const char* names[3] = {"16x256b.x2", "16x128b.x2", "16x64b.x4"};
    const int nreg[3] = {8, 4, 4};
    for (int s = 0; s < 3; ++s)
        for (int sub = 0; sub < 2; ++sub) {
            printf("%s base lane %d:\n", names[s], 32 + 16 * sub);
            for (int t = 0; t < 32; ++t) {
                printf("  t%02d:", t);
                for (int j = 0; j < nreg[s]; ++j) {
                    uint32_t v = h[(s * 2 + sub) * 256 + t * 8 + j];
                    printf(" (%u,%u)", v / 1000, v % 1000);
                }
                printf("\n");
            }
        }
    return 0;
}
