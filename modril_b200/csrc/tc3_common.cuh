// Device helpers shared by the CTA-pair (cta_group::2) integrators logprob_tc3.cu (one-pass 16-bit operands) and
// logprob_tc3x.cu (3-term fp16 operand split): 16-bit conversions, stmatrix, the 16x256b TMEM loads, cluster / DSMEM
// barrier primitives and the pair-wide tcgen05.mma / commit wrappers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace frl {
namespace tc3h {

using namespace tcc;

template <typename T> struct Cvt;
template <> struct Cvt<__half> {
    static constexpr uint32_t kFmt = 0;
    __device__ static __forceinline__ uint32_t pack2_relu(uint32_t lo, uint32_t hi) {
        uint32_t d;
        asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
        return d;
    }
    __device__ static __forceinline__ uint32_t pack2_sat(uint32_t lo, uint32_t hi) {
        uint32_t d;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
        return d;
    }
    __device__ static __forceinline__ uint16_t one(float x) {
        x = fminf(fmaxf(x, -65504.f), 65504.f);
        __half h = __float2half_rn(x);
        return *reinterpret_cast<uint16_t*>(&h);
    }
};
template <> struct Cvt<__nv_bfloat16> {
    static constexpr uint32_t kFmt = 1;
    __device__ static __forceinline__ uint32_t pack2_relu(uint32_t lo, uint32_t hi) {
        uint32_t d;
        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
        return d;
    }
    __device__ static __forceinline__ uint32_t pack2_sat(uint32_t lo, uint32_t hi) {
        uint32_t d;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
        return d;
    }
    __device__ static __forceinline__ uint16_t one(float x) {
        __nv_bfloat16 h = __float2bfloat16_rn(x);
        return *reinterpret_cast<uint16_t*>(&h);
    }
};

// 0xFFFF in each half whose source float has its sign bit set (z < 0 or -0): PRMT with sign replication
__device__ __forceinline__ uint32_t neg_mask16x2(uint32_t z_lo, uint32_t z_hi) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(d) : "r"(z_lo), "r"(z_hi));
    return d;
}

// four 8x8 b16 matrices from the MMA-fragment layout (thread t: row t/4, columns 2(t%4), +1) to shared memory; lane i
// supplies the address of row i%8 of matrix i/8 (16 bytes per row)
__device__ __forceinline__ void stmatrix_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}

struct Tracer {
    unsigned long long* buf;
    int n;
    __device__ __forceinline__ void init(unsigned long long* base, int role, int cta = 0) {
        buf = (base != nullptr && blockIdx.x == cta) ? base + (size_t)role * 8192 : nullptr;
        n = 0;
    }
    __device__ __forceinline__ void mark(unsigned tag) {
        if (buf && n < 8190) buf[n++] = ((unsigned long long)tag << 48) | (clock64() & 0xFFFFFFFFFFFFull);
    }
};

__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// 16 lanes x (8 * N) columns: thread t receives, for k = 0..N-1,  r[4k..4k+3] = D[base + t/4][8k + 2(t%4) + {0,1}],
// D[base + 8 + t/4][8k + 2(t%4) + {0,1}]  (layout probed by tests/tmem_ld_probe.cu)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait_dep8(uint32_t (&a)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                 :
                 : "memory");
}

// operand row of sample i (0..63): groups of 16 rows = 8 primal rows followed by the 8 tangent rows of the same samples,
// so that one 16x256b TMEM load hands a thread the primal AND the tangent of its sample
__device__ __forceinline__ int row_primal(int i) { return ((i >> 3) << 4) | (i & 7); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster (release at cluster scope: the
// shared-memory operand rows written before it must be visible to the MMA the leader issues afterwards)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(bar), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs of this thread -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

}  // namespace tc3h
}  // namespace frl
