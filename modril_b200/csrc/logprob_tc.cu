// tcgen05 tensor-core fused log-density integrator (sm_100a).
//
// One persistent CTA per SM; a tile is 128 samples x one role.  The primal rows (tile "p") and the forward-mode
// tangent rows (tile "t") of the SAME 128 samples are two M=128 UMMA operands that live at the same TMEM lane /
// shared-memory row, so the epilogue thread of row r applies the ReLU mask of the primal to the tangent in
// registers.  All n_steps Euler steps of pretrain.py:192-214 / :112-134 run on chip:
//
//   warp 0      producer : streams pre-packed 16 KB weight blocks (16-bit, UMMA canonical no-swizzle K-major
//                          layout) from L2 into a shared-memory ring with cp.async.bulk + mbarrier complete_tx
//   warp 1      MMA      : one elected thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM); every block
//                          is used for the primal and the tangent tile back to back (weights cross L2->SM once)
//   warps 2..9  epilogue : tcgen05.ld -> +bias, ReLU, mask the tangent -> 16-bit -> next layer's A operand in shared
//                          memory (in place); heads -> tanh, divergence, Euler update of the fp32 state
//
// Each layer's N is split in two halves with separate accumulators D[p|t][h]; the block order
//   (h0,K-lo) (h1,K-lo) (h0,K-hi) (h1,K-hi)
// makes the K-lo half of the activation buffer dead before the epilogue of h0 overwrites it, so activations are
// updated IN PLACE (128 KB for both tiles) while the epilogue of one half overlaps the MMAs of the other.
//
// Shared-memory operand layout (A, X0, weight blocks):  [k/8][row][8 elements]  -> 8-row x 16-byte core matrices
// are contiguous (SBO = 128 B), consecutive k-chunks are rows*16 B apart (LBO).  Verified by tests/umma_probe.cu.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "frl_internal.h"

namespace frl {

namespace tc {

constexpr int kThreads = 416;          // 13 warps: producer (+TMEM owner), 4 MMA issuers, 8 epilogue
constexpr int kMmaWarps = 4;           // (accumulator half) x (primal | tangent)
constexpr int kEpiWarps = 8;
constexpr int kStageBytes = 16384;
constexpr int kRows = 128;

struct Blk {
    uint32_t goff;      // byte offset of the block in the packed weight buffer
    uint32_t bytes;
    uint16_t nrows;     // N of the UMMA (rows of the weight block)
    uint16_t a_koff;    // first k-chunk (of 8) of the A operand this block multiplies
    uint8_t n_k16;      // K=16 MMAs per tile for the primal
    uint8_t n_k16_t;    // ... for the tangent (layer 0: only the action columns carry a tangent)
    uint8_t src;        // A operand: 0 activation buffer, 1 layer-0 input buffer, 2 the constant "ones" tile (bias)
    uint8_t h;          // accumulator half
    uint8_t first;      // bit0: first primal MMA overwrites the accumulator, bit1: same for the tangent
    uint8_t commit;     // 0: none, 1: commit d_full[0], 2: commit d_full[1]
    uint8_t wait;       // bit0: wait a_ready[0], bit1: wait a_ready[1] before issuing
    uint8_t pad;
};

struct Params {
    const uint8_t* wpack[2];
    const float* bias[2];       // [depth*H + NOp]
    float sign[2];
    float* out[2];
    const float* s;
    const float* a;
    const float* probe;
    long long probe_step_stride;
    long long B;
    int n_roles, n_heads, depth, s_dim, a_dim, d_in, K0p, KT, NOp, H;
    int n_steps, basis, accumulate, n_blocks, n_stages;
    unsigned long long* trace;  // FRL_TC_TRACE: [role][8192] clock stamps of CTA 0 (timing experiments only)
    int debug;  // FRL_TC_DEBUG env: bit0 skip trunk-epilogue work, bit1 skip MMA issue (timing experiments only)
    float delta;
    // shared-memory byte offsets
    uint32_t off_a, off_x0p, off_x0t, off_ones, off_bias, off_state, off_rec, off_ring;
    Blk blk[kMaxBlocks];
    float t_mid[kMaxSteps];
};

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    // no-swizzle K-major: SBO (8-row group stride) = 128 B, LBO (k-chunk stride) = lbo_bytes, version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that also names the destination registers, so the compiler cannot hoist their uses above the wait
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                   "+r"(a[16]), "+r"(a[17]), "+r"(a[18]), "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]),
                   "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]), "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]),
                   "+r"(b[16]), "+r"(b[17]), "+r"(b[18]), "+r"(b[19]), "+r"(b[20]), "+r"(b[21]), "+r"(b[22]), "+r"(b[23]),
                   "+r"(b[24]), "+r"(b[25]), "+r"(b[26]), "+r"(b[27]), "+r"(b[28]), "+r"(b[29]), "+r"(b[30]), "+r"(b[31])
                 :
                 : "memory");
}
// 0xFFFF in each half whose source float has its sign bit set (z < 0 or -0): PRMT with sign replication
__device__ __forceinline__ uint32_t neg_mask16x2(uint32_t z_lo, uint32_t z_hi) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(d) : "r"(z_lo), "r"(z_hi));
    return d;
}

struct Tracer {
    unsigned long long* buf;
    int n;
    __device__ __forceinline__ void init(unsigned long long* base, int role) {
        buf = (base != nullptr && blockIdx.x == 0) ? base + (size_t)role * 8192 : nullptr;
        n = 0;
    }
    __device__ __forceinline__ void mark(unsigned tag) {
        if (buf && n < 8190) {
            buf[n++] = ((unsigned long long)tag << 48) | (clock64() & 0xFFFFFFFFFFFFull);
        }
    }
};

// 16-bit operand types
template <typename T> struct Cvt;
template <> struct Cvt<__half> {
    static constexpr uint32_t kFmt = 0;  // F16
    // relu + round + saturate + pack in ONE instruction (F2FP.SATFINITE.RELU.F16.F32.PACK_AB)
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
    __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
        lo = fminf(fmaxf(lo, -65504.f), 65504.f);
        hi = fminf(fmaxf(hi, -65504.f), 65504.f);
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ static __forceinline__ uint16_t one(float x) {
        x = fminf(fmaxf(x, -65504.f), 65504.f);
        __half h = __float2half_rn(x);
        return *reinterpret_cast<uint16_t*>(&h);
    }
};
template <> struct Cvt<__nv_bfloat16> {
    static constexpr uint32_t kFmt = 1;  // BF16
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
    __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ static __forceinline__ uint16_t one(float x) {
        __nv_bfloat16 h = __float2bfloat16_rn(x);
        return *reinterpret_cast<uint16_t*>(&h);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename T16, int H>
__global__ void __launch_bounds__(kThreads, 1) logprob_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // header: barriers
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0..7] w_full, [8..15] w_empty, [16,17] a_ready, [18,19] d_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    const uint32_t bar0 = smem_u32(bars);
    auto w_full = [&](int s) { return bar0 + 8u * s; };
    auto w_empty = [&](int s) { return bar0 + 8u * (8 + s); };
    auto a_ready = [&](int i) { return bar0 + 8u * (16 + i); };
    auto d_full = [&](int i) { return bar0 + 8u * (18 + i); };

    uint8_t* sA = smem + p.off_a;        // [2 tiles][H/8][128][8] T16
    uint8_t* sX0p = smem + p.off_x0p;    // [K0p/8][128][8]
    uint8_t* sX0t = smem + p.off_x0t;    // [KT/8][128][8]
    float* sBias = reinterpret_cast<float*>(smem + p.off_bias);    // [2 roles][NOp]: head biases (trunk biases ride in the MMA)
    float* sState = reinterpret_cast<float*>(smem + p.off_state);  // a [a_dim][128], probe [a_dim][128]
    uint8_t* sRing = smem + p.off_ring;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NH = H / 2;
    constexpr uint32_t a_tile_bytes = (uint32_t)H * kRows * 2;  // one activation tile
    uint8_t* sOnes = smem + p.off_ones;  // [2][128][8]: columns 0..2 are 1.0 (multiplies the bias hi/mid/lo rows)

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), kMmaWarps);   // one arrival per MMA-issuing warp
        }
        mbar_init(a_ready(0), kEpiWarps);
        mbar_init(a_ready(1), kEpiWarps);
        mbar_init(d_full(0), kMmaWarps);
        mbar_init(d_full(1), kMmaWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // biases of both roles, zero the (padded) layer-0 input buffers
    for (int i = threadIdx.x; i < p.NOp * p.n_roles; i += kThreads)
        sBias[i] = p.bias[i / p.NOp][p.depth * H + i % p.NOp];
    for (int i = threadIdx.x; i < (p.K0p + p.KT) * kRows * 2 / 16; i += kThreads)
        reinterpret_cast<uint4*>(sX0p)[i] = make_uint4(0, 0, 0, 0);  // X0t follows X0p contiguously
    for (int i = threadIdx.x; i < 2 * kRows; i += kThreads) {
        const uint32_t one = Cvt<T16>::one(1.f);
        reinterpret_cast<uint4*>(sOnes)[i] = i < kRows ? make_uint4(one | (one << 16), one, 0, 0) : make_uint4(0, 0, 0, 0);
    }
    // per-block issue records of the four MMA warps (two 16-byte shared-memory loads per block in the hot loop).
    // Work split: MMA warp w = 2*h + x issues tile x (0 primal, 1 tangent) of accumulator half h for the trunk /
    // layer-0 / bias blocks; in the heads blocks warp 0 issues the primal tile and warp 1 the tangent tile.  One thread
    // pays ~60 cycles per tcgen05.mma and ~230 per tcgen05.commit (tests/umma_bench.cu) while an N = 128 MMA holds the
    // tensor pipe for 64, so the issue work is spread over four threads; primal and tangent go to different
    // accumulators, so their relative order does not matter and the results stay deterministic.
    uint4* sRec = reinterpret_cast<uint4*>(smem + p.off_rec);  // [4 warps][n_blocks][2]
    for (int i = threadIdx.x; i < kMmaWarps * p.n_blocks; i += kThreads) {
        const int w = i / p.n_blocks, b = i % p.n_blocks;
        const Blk& k = p.blk[b];
        const bool heads = k.nrows != NH;  // only the heads blocks have N != H/2 (NOp <= 64 < H/2)
        const int x = w & 1;
        const bool mine = heads ? (w < 2) : ((int)k.h == (w >> 1));
        uint32_t a_lo = 0, dcol = 0, nk = 0, first = 0;
        if (mine) {
            const uint32_t src = k.src == 0 ? smem_u32(smem + p.off_a) + x * a_tile_bytes
                                 : (k.src == 1 ? smem_u32(x == 0 ? sX0p : sX0t) : smem_u32(sOnes));
            a_lo = (((src + (uint32_t)k.a_koff * (kRows * 16)) >> 4) & 0x3FFF) | ((uint32_t)(kRows * 16 >> 4) << 16);
            dcol = (uint32_t)k.h * 256 + (uint32_t)x * 128;
            nk = x == 0 ? k.n_k16 : k.n_k16_t;
            first = (k.first >> x) & 1;
        }
        const uint32_t idesc = (1u << 4) | (Cvt<T16>::kFmt << 7) | (Cvt<T16>::kFmt << 10) | ((uint32_t)(kRows >> 4) << 24) |
                               ((uint32_t)(k.nrows >> 3) << 17);
        sRec[2 * i] = make_uint4(a_lo, dcol, idesc, (uint32_t)k.nrows | ((uint32_t)k.commit << 16) | ((uint32_t)k.wait << 24));
        sRec[2 * i + 1] = make_uint4(nk | (first << 8), 0, 0, 0);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const long long tiles_per_role = (p.B + kRows - 1) / kRows;
    const long long n_tiles = tiles_per_role * p.n_roles;

    if (warp == 0) {
        // ------------------------------------------------ producer ------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            Tracer tr; tr.init(p.trace, 0);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint8_t* wbase = p.wpack[tile % p.n_roles];
                for (int step = 0; step < p.n_steps; ++step) {
                    for (int b = 0; b < p.n_blocks; ++b) {
                        mbar_wait(w_empty(stage), phase ^ 1);
                        tr.mark(b);
                        if (p.debug & 4) {
                            mbar_arrive(w_full(stage));  // timing experiment: no weight traffic
                        } else {
                            mbar_expect_tx(w_full(stage), p.blk[b].bytes);
                            bulk_g2s(smem_u32(sRing + (size_t)stage * kStageBytes), wbase + p.blk[b].goff,
                                     p.blk[b].bytes, w_full(stage));
                        }
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 1 && warp <= kMmaWarps) {
        // ------------------------------ MMA issuers (see the record table above for the work split) ----------------
        if (lane == 0) {
            const int j = warp - 1;
            int stage = 0;
            uint32_t phase = 0, par_a0 = 0, par_a1 = 0;
            const uint4* rec = sRec + (size_t)j * p.n_blocks * 2;
            const uint32_t ring_lo = (smem_u32(sRing) >> 4);
            constexpr uint32_t kDescHi = (uint32_t)(128 >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1
            Tracer tr; tr.init(j < 2 ? p.trace : nullptr, 1 + j);
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int step = 0; step < p.n_steps; ++step) {
                    for (int b = 0; b < p.n_blocks; ++b) {
                        const uint4 r0 = rec[2 * b], r1 = rec[2 * b + 1];
                        tr.mark(0x100 + b);
                        // every issuer observes every readiness phase of the activations, whether or not it issues here
                        if (r0.w & (1u << 24)) { mbar_wait(a_ready(0), par_a0); par_a0 ^= 1; }
                        if (r0.w & (2u << 24)) { mbar_wait(a_ready(1), par_a1); par_a1 ^= 1; }
                        const uint32_t nk = (p.debug & 2) ? 0 : (r1.x & 0xFF);
                        const uint32_t cm = (r0.w >> 16) & 0xFF;
                        // every issuer observes every fill of the ring (a warp that ran ahead through blocks it does
                        // not own could otherwise arrive twice on one w_empty phase and release a stage too early)
                        mbar_wait(w_full(stage), phase);
                        if (nk) {
                            tc_fence_after();
                            tr.mark(0x300 + b);
                            const uint32_t nrows = r0.w & 0xFFFF;
                            uint32_t a_lo = r0.x;
                            uint32_t b_lo = ((ring_lo + (uint32_t)stage * (kStageBytes >> 4)) & 0x3FFF) | (nrows << 16);
                            uint32_t acc = ((r1.x >> 8) & 1) ? 0u : 1u;
                            for (uint32_t u = 0; u < nk; ++u) {
                                umma(tmem + r0.y, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, r0.z, acc);
                                acc = 1u;
                                a_lo += 2 * (kRows * 16 >> 4);   // two k-chunks of 8 per K=16 MMA
                                b_lo += 2 * nrows;
                            }
                            umma_commit(w_empty(stage));
                            if (cm) umma_commit(d_full(cm - 1));
                        } else {
                            // not this warp's block: release the stage / accumulator slot without touching the pipe
                            mbar_arrive(w_empty(stage));
                            if (cm) mbar_arrive(d_full(cm - 1));
                        }
                        tr.mark(0x400 + b);
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp > kMmaWarps) {
        // ------------------------------------------------ epilogue ------------------------------------------------
        const int ew = warp - (kMmaWarps + 1);   // 0..7
        const int q = warp & 3;             // TMEM lane quarter this warp may access
        const int ch = ew >> 2;             // which half of the N-half's columns
        const int row = q * 32 + lane;      // tile row == TMEM lane
        const int et = ew * 32 + lane;      // 0..255 epilogue thread id
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        float* sA_state = sState;
        float* sP_state = sState + p.a_dim * kRows;
        const int t_col = p.a_dim + p.s_dim;
        uint32_t par_d0 = 0, par_d1 = 0;
        Tracer tr; tr.init((lane == 0 && ew == 0) ? p.trace : nullptr, 3);

        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int role = (int)(tile % p.n_roles);
            const long long row0 = (tile / p.n_roles) * kRows;
            const float sign = p.sign[role];
            const float* bias_heads = sBias + role * p.NOp;

            // ---- tile init: layer-0 inputs + fp32 state ----
            for (int idx = et; idx < kRows * p.s_dim; idx += kEpiWarps * 32) {
                const int r = idx / p.s_dim, c = idx % p.s_dim;
                const long long g = row0 + r;
                const float v = g < p.B ? p.s[g * p.s_dim + c] : 0.f;
                const int kk = p.a_dim + c;
                *reinterpret_cast<uint16_t*>(sX0p + (kk >> 3) * (kRows * 16) + r * 16 + (kk & 7) * 2) = Cvt<T16>::one(v);
            }
            for (int idx = et; idx < kRows * p.a_dim; idx += kEpiWarps * 32) {
                const int r = idx / p.a_dim, c = idx % p.a_dim;
                const long long g = row0 + r;
                const float v = g < p.B ? p.a[g * p.a_dim + c] : 0.f;
                float pr = 0.f;
                if (g < p.B) pr = p.basis >= 0 ? (c == p.basis ? 1.f : 0.f) : p.probe[g * p.a_dim + c];
                sA_state[c * kRows + r] = v;
                sP_state[c * kRows + r] = pr;
                *reinterpret_cast<uint16_t*>(sX0p + (c >> 3) * (kRows * 16) + r * 16 + (c & 7) * 2) = Cvt<T16>::one(v);
                *reinterpret_cast<uint16_t*>(sX0t + (c >> 3) * (kRows * 16) + r * 16 + (c & 7) * 2) = Cvt<T16>::one(pr);
            }
            if (et < kRows)
                *reinterpret_cast<uint16_t*>(sX0p + (t_col >> 3) * (kRows * 16) + et * 16 + (t_col & 7) * 2) =
                    Cvt<T16>::one(p.t_mid[0]);
            float log_det = 0.f;
            // all epilogue warps must have finished the init before anyone signals (named barrier 1, 256 threads)
            fence_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) { mbar_arrive(a_ready(0)); mbar_arrive(a_ready(1)); }

            for (int step = 0; step < p.n_steps; ++step) {
                // ---- trunk layers ----
                for (int l = 0; l < p.depth; ++l) {
                    for (int h = 0; h < 2; ++h) {
                        tr.mark(0x100 + l * 2 + h);
                        if (h == 0) { mbar_wait(d_full(0), par_d0); par_d0 ^= 1; }
                        else        { mbar_wait(d_full(1), par_d1); par_d1 ^= 1; }
                        tc_fence_after();
                        tr.mark(0x200 + l * 2 + h);
                        if (!(p.debug & 1)) {
                            // this warp's CPW columns of the half: all TMEM loads are issued up front, one wait
                            constexpr int CPW = NH / 2, NCH = CPW / 32;
                            uint32_t zp[NCH][32], zt[NCH][32];
                            const uint32_t tb = lane_addr + (uint32_t)h * 256 + (uint32_t)(ch * CPW);
                            tmem_ld32(tb, zp[0]);
                            tmem_ld32(tb + 128, zt[0]);
                            if (NCH == 2) {
                                tmem_ld32(tb + 32, zp[NCH - 1]);
                                tmem_ld32(tb + 128 + 32, zt[NCH - 1]);
                            }
                            tmem_ld_wait_dep(zp[0], zt[0]);
                            if (NCH == 2) tmem_ld_wait_dep(zp[NCH - 1], zt[NCH - 1]);
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                const int n0 = h * NH + ch * CPW + c * 32;  // output column == next layer's k index
                                uint32_t op[16], ot[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    // bias is already in the accumulator (folded into the MMA); relu+round+pack is one
                                    // F2FP; the tangent is masked by the primal's sign bits
                                    op[j] = Cvt<T16>::pack2_relu(zp[c][2 * j], zp[c][2 * j + 1]);
                                    ot[j] = Cvt<T16>::pack2_sat(zt[c][2 * j], zt[c][2 * j + 1]) &
                                            ~neg_mask16x2(zp[c][2 * j], zp[c][2 * j + 1]);
                                }
                                uint8_t* dp = sA + (size_t)(n0 >> 3) * (kRows * 16) + row * 16;
                                uint8_t* dt = dp + a_tile_bytes;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    *reinterpret_cast<uint4*>(dp + (size_t)i * (kRows * 16)) =
                                        make_uint4(op[4 * i], op[4 * i + 1], op[4 * i + 2], op[4 * i + 3]);
                                    *reinterpret_cast<uint4*>(dt + (size_t)i * (kRows * 16)) =
                                        make_uint4(ot[4 * i], ot[4 * i + 1], ot[4 * i + 2], ot[4 * i + 3]);
                                }
                            }
                        }
                        tc_fence_before();
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a_ready(h));
                        tr.mark(0x300 + l * 2 + h);
                    }
                }
                tr.mark(0x1FF);
                // ---- heads: tanh, divergence, Euler update (pretrain.py:205-211) ----
                mbar_wait(d_full(0), par_d0); par_d0 ^= 1;
                tc_fence_after();
                tr.mark(0x2FF);
                const long long g = row0 + row;
                if (ch == 0) {
                    const float* bh = bias_heads;
                    float div = 0.f;
                    for (int g16 = 0; g16 < p.NOp; g16 += 16) {
                        uint32_t vp[16], vt[16];
                        tmem_ld16(lane_addr + (uint32_t)g16, vp);
                        tmem_ld16(lane_addr + 128 + (uint32_t)g16, vt);
                        tmem_ld_wait();
                        if (p.n_heads == 2) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int j = (g16 >> 1) + i;
                                if (j < p.a_dim) {
                                    const float vc = __uint_as_float(vp[2 * i]) + bh[g16 + 2 * i];
                                    const float rp = __uint_as_float(vp[2 * i + 1]) + bh[g16 + 2 * i + 1];
                                    const float th = tanhf(rp);
                                    const float v = vc + sign * th;
                                    const float dv = __uint_as_float(vt[2 * i]) +
                                                     sign * (1.f - th * th) * __uint_as_float(vt[2 * i + 1]);
                                    div = fmaf(dv, sP_state[j * kRows + row], div);
                                    const float an = sA_state[j * kRows + row] + v * p.delta;
                                    sA_state[j * kRows + row] = an;
                                    *reinterpret_cast<uint16_t*>(sX0p + (j >> 3) * (kRows * 16) + row * 16 + (j & 7) * 2) =
                                        Cvt<T16>::one(an);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int j = g16 + i;
                                if (j < p.a_dim) {
                                    const float v = __uint_as_float(vp[i]) + bh[g16 + i];
                                    const float dv = __uint_as_float(vt[i]);
                                    div = fmaf(dv, sP_state[j * kRows + row], div);
                                    const float an = sA_state[j * kRows + row] + v * p.delta;
                                    sA_state[j * kRows + row] = an;
                                    *reinterpret_cast<uint16_t*>(sX0p + (j >> 3) * (kRows * 16) + row * 16 + (j & 7) * 2) =
                                        Cvt<T16>::one(an);
                                }
                            }
                        }
                    }
                    log_det -= p.delta * div;
                    if (step + 1 < p.n_steps) {
                        *reinterpret_cast<uint16_t*>(sX0p + (t_col >> 3) * (kRows * 16) + row * 16 + (t_col & 7) * 2) =
                            Cvt<T16>::one(p.t_mid[step + 1]);
                        if (p.probe_step_stride != 0) {
                            for (int j = 0; j < p.a_dim; ++j) {
                                const float pr = g < p.B ? p.probe[(step + 1) * p.probe_step_stride + g * p.a_dim + j] : 0.f;
                                sP_state[j * kRows + row] = pr;
                                *reinterpret_cast<uint16_t*>(sX0t + (j >> 3) * (kRows * 16) + row * 16 + (j & 7) * 2) =
                                    Cvt<T16>::one(pr);
                            }
                        }
                    } else if (g < p.B) {
                        if (p.accumulate) {
                            p.out[role][g] += log_det;
                        } else {
                            float prior = 0.f;
                            for (int j = 0; j < p.a_dim; ++j) {
                                const float x = sA_state[j * kRows + row];
                                prior += -0.5f * x * x - 0.91893853320467274178f;
                            }
                            p.out[role][g] = prior + log_det;
                        }
                    }
                }
                if (step + 1 < p.n_steps) {
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(a_ready(0)); mbar_arrive(a_ready(1)); }
                    tr.mark(0x3FF);
                } else {
                    // the next tile's init rewrites X0 / state: every warp must be done with this tile first
                    tc_fence_before();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: fp32 flat parameters -> fp16 and bf16 block images + fp32 bias vector
// ---------------------------------------------------------------------------------------------------------------
struct PackParams {
    PackBlk blk[kMaxBlocks];
    int n_blocks;
    long long off_w[kMaxDepth], off_b[kMaxDepth], off_wh[2], off_bh[2];
    int depth, H, d_in, a_dim, n_heads, NOp;
    uint32_t image_bytes;   // bytes of one 16-bit image
    float scale;            // weights and trunk biases are multiplied by it (a power of two) before rounding / splitting
};

__global__ void tc_pack_kernel(const float* __restrict__ flat, uint8_t* __restrict__ img_f16,
                               uint8_t* __restrict__ img_bf16, float* __restrict__ bias, PackParams pp) {
    const int b = blockIdx.y;
    if (b == pp.n_blocks) {  // bias vector
        const int nb = pp.depth * pp.H + pp.NOp;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) {
            float v = 0.f;
            if (i < pp.depth * pp.H) {
                v = flat[pp.off_b[i / pp.H] + i % pp.H];
            } else {
                const int o = i - pp.depth * pp.H;
                const int j = pp.n_heads == 2 ? o >> 1 : o, head = pp.n_heads == 2 ? (o & 1) : 0;
                if (j < pp.a_dim) v = flat[pp.off_bh[head] + j];
            }
            bias[i] = v;
        }
        return;
    }
    const PackBlk& k = pp.blk[b];
    const int chunks = k.nrows * (k.klen / 8);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
        const int j = c / k.nrows, n = c % k.nrows;  // k-chunk, row
        __half hv[8];
        __nv_bfloat16 bv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kg = k.k0 + j * 8 + e, ng = k.n0 + n;
            float v = 0.f;
            if (k.k0 < 0) {
                // bias block: column e of k-chunk 0 holds the e-th term of the 16-bit split of bias[n]
                hv[e] = __float2half_rn(0.f);
                bv[e] = __float2bfloat16_rn(0.f);
                if (j == 0 && e < 3) {
                    const float b = flat[pp.off_b[k.layer] + ng] * pp.scale;
                    float rh = b, rb = b;
                    __half th = __float2half_rn(0.f);
                    __nv_bfloat16 tb = __float2bfloat16_rn(0.f);
                    for (int term = 0; term <= e; ++term) {
                        th = __float2half_rn(fminf(fmaxf(rh, -65504.f), 65504.f));
                        tb = __float2bfloat16_rn(rb);
                        rh -= __half2float(th);
                        rb -= __bfloat162float(tb);
                    }
                    hv[e] = th;
                    bv[e] = tb;
                }
                continue;
            }
            if (k.layer < pp.depth) {
                const int K = k.layer == 0 ? pp.d_in : pp.H;
                if (kg < K) v = flat[pp.off_w[k.layer] + (long long)ng * K + kg];
            } else {
                const int a = pp.n_heads == 2 ? ng >> 1 : ng, head = pp.n_heads == 2 ? (ng & 1) : 0;
                if (a < pp.a_dim) v = flat[pp.off_wh[head] + (long long)a * pp.H + kg];
            }
            v *= pp.scale;
            hv[e] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
            bv[e] = __float2bfloat16_rn(v);
            if (k.part == 2) {   // lo term of the 2-term split
                hv[e] = __float2half_rn(v - __half2float(hv[e]));
                bv[e] = __float2bfloat16_rn(v - __bfloat162float(bv[e]));
            }
        }
        const size_t o = (size_t)k.goff + (size_t)c * 16;
        *reinterpret_cast<uint4*>(img_f16 + o) = *reinterpret_cast<uint4*>(hv);
        *reinterpret_cast<uint4*>(img_bf16 + o) = *reinterpret_cast<uint4*>(bv);
    }
}

// Block schedule of one Euler step (shared by the packer and the kernel).
struct Schedule {
    std::vector<Blk> blk;
    std::vector<PackBlk> pack;
    uint32_t image_bytes = 0;
    int KT = 16;
};

static Schedule make_schedule(const frl_net* net) {
    Schedule sc;
    const int H = net->cfg.hidden, NH = H / 2, depth = net->cfg.depth, K0p = net->d_in_p, NOp = net->n_out_p;
    sc.KT = round_up(net->cfg.a_dim, 16);
    uint32_t goff = 0;
    auto add = [&](int layer, int n0, int nrows, int k0, int klen, int a_koff, int src, int h, int first, int commit,
                   int wait, int nk_t) {
        Blk b{};
        b.goff = goff; b.bytes = (uint32_t)nrows * klen * 2; b.nrows = (uint16_t)nrows; b.a_koff = (uint16_t)a_koff;
        b.n_k16 = (uint8_t)(klen / 16); b.n_k16_t = (uint8_t)nk_t; b.src = (uint8_t)src; b.h = (uint8_t)h;
        b.first = (uint8_t)first; b.commit = (uint8_t)commit; b.wait = (uint8_t)wait;
        sc.blk.push_back(b);
        sc.pack.push_back(PackBlk{goff, nrows, klen, n0, k0, layer});
        goff += b.bytes;
    };
    // bias block of (layer, half): ones[128 x 16] * [bias_hi; bias_mid; bias_lo; 0...]^T starts the primal accumulator
    auto add_bias = [&](int layer, int h, int wait) { add(layer, h * NH, NH, -1, 16, 0, 2, h, 1, 0, wait, 0); };
    // layer 0: A operand = X0 (K0p columns); the tangent only has the first KT columns
    for (int h = 0; h < 2; ++h) {
        add_bias(0, h, h == 0 ? 3 : 0);
        for (int k0 = 0; k0 < K0p; k0 += 64) {
            const int klen = std::min(64, K0p - k0);
            const int nk_t = std::max(0, std::min(klen, sc.KT - k0)) / 16;
            add(0, h * NH, NH, k0, klen, k0 / 8, 1, h, k0 == 0 ? 2 : 0, (k0 + 64 >= K0p) ? h + 1 : 0, 0, nk_t);
        }
    }
    // trunk layers 1..depth-1: (h0,K-lo) (h1,K-lo) (h0,K-hi) (h1,K-hi); the K-lo half of the in-place activation
    // buffer is dead once (h1,K-lo) has been issued, i.e. before d_full[0] can complete
    const int nkq = H / 64;
    for (int l = 1; l < depth; ++l) {
        for (int khalf = 0; khalf < 2; ++khalf) {
            for (int h = 0; h < 2; ++h) {
                if (khalf == 0) add_bias(l, h, h == 0 ? 1 : 2);  // h0: A k-lo ready & D[h0] free; h1: D[h1] free
                for (int kq = khalf * nkq / 2; kq < (khalf + 1) * nkq / 2; ++kq) {
                    const bool last = khalf == 1 && kq == nkq - 1;
                    add(l, h * NH, NH, kq * 64, 64, kq * 8, 0, h, (khalf == 0 && kq == 0) ? 2 : 0, last ? h + 1 : 0, 0, 4);
                }
            }
        }
    }
    // heads: N = NOp, all K (bias added in the epilogue in fp32)
    for (int kq = 0; kq < nkq; ++kq) {
        int wait = kq == 0 ? 1 : (kq == nkq / 2 ? 2 : 0);
        add(depth, 0, NOp, kq * 64, 64, kq * 8, 0, 0, kq == 0 ? 3 : 0, kq == nkq - 1 ? 1 : 0, wait, 4);
    }
    sc.image_bytes = goff;
    return sc;
}

struct TcPack {
    uint8_t* img_f16;
    uint8_t* img_bf16;
    float* bias;
};

}  // namespace tc

// fp16 + bf16 images of a block list (+ the fp32 bias vector) into *dst (allocated on first use):
// layout [fp16 image | bf16 image | bias], each image rounded up to 256 B
int tc_pack_images(frl_net* net, const std::vector<tc::PackBlk>& pack, uint32_t image_bytes, void** dst,
                   const float* params_dev, cudaStream_t stream, float scale) {
    using namespace tc;
    const int nb = net->cfg.depth * net->cfg.hidden + net->n_out_p;
    const size_t img = ((size_t)image_bytes + 255) / 256 * 256;
    const size_t total = 2 * img + (size_t)nb * sizeof(float);
    if (!*dst) FRL_CUDA(cudaMalloc(dst, total));
    PackParams pp{};
    pp.n_blocks = (int)pack.size();
    for (int i = 0; i < pp.n_blocks; ++i) pp.blk[i] = pack[i];
    for (int l = 0; l < net->cfg.depth; ++l) { pp.off_w[l] = net->off_w[l]; pp.off_b[l] = net->off_b[l]; }
    for (int h = 0; h < net->cfg.n_heads; ++h) { pp.off_wh[h] = net->off_wh[h]; pp.off_bh[h] = net->off_bh[h]; }
    pp.depth = net->cfg.depth; pp.H = net->cfg.hidden; pp.d_in = net->d_in; pp.a_dim = net->cfg.a_dim;
    pp.n_heads = net->cfg.n_heads; pp.NOp = net->n_out_p; pp.image_bytes = image_bytes; pp.scale = scale;
    uint8_t* base = static_cast<uint8_t*>(*dst);
    tc_pack_kernel<<<dim3(4, pp.n_blocks + 1), 256, 0, stream>>>(params_dev, base, base + img,
                                                                 reinterpret_cast<float*>(base + 2 * img), pp);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

int tc_pack(frl_net* net, const float* params_dev, cudaStream_t stream) {
    using namespace tc;
    Schedule sc = make_schedule(net);
    if ((int)sc.blk.size() > kMaxBlocks) return FRL_OK;
    return tc_pack_images(net, sc.pack, sc.image_bytes, &net->pack_tc, params_dev, stream);
}

template <typename T16, int H>
static int launch_tc(const tc::Params& p, size_t smem_bytes, int grid, cudaStream_t stream) {
    auto kern = tc::logprob_tc_kernel<T16, H>;
    FRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    kern<<<grid, tc::kThreads, smem_bytes, stream>>>(p);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

int launch_logprob_tc(const ScoreArgs& a, int precision, cudaStream_t stream) {
    using namespace tc;
    const frl_net* n0 = a.net[0];
    // Kernel selection.  Default: the CTA-pair integrator (logprob_tc3.cu, cta_group::2) whenever the net fits its plan
    // (H = 256, padded input width <= 192, <= 64 head columns, >= 4 ring slots); otherwise -- or with FRL_TC_V1=1 -- the
    // 128-sample single-CTA kernel below.
    if (!getenv("FRL_TC_V1")) {
        for (int r = 0; r < a.n_roles; ++r) {
            int rcp = ensure_pack(a.net[r], PACK_TC3, stream);
            if (rcp != FRL_OK) return rcp;
        }
        if (n0->pack_tc3) {
            const int rc3 = launch_logprob_tc3(a, precision, stream);
            if (rc3 <= 0) return rc3;
        }
    }
    for (int r = 0; r < a.n_roles; ++r) {
        int rcp = ensure_pack(a.net[r], PACK_TC1, stream);
        if (rcp != FRL_OK) return rcp;
    }
    if (!n0->pack_tc) {
        set_error("tensor-core weight image missing (network too deep for the block table)");
        return FRL_ERR_UNSUPPORTED;
    }
    Schedule sc = make_schedule(n0);
    const int H = n0->cfg.hidden;
    const int nb = n0->cfg.depth * H + n0->n_out_p;
    const size_t img = (sc.image_bytes + 255) / 256 * 256;
    const bool f16 = precision == FRL_PREC_FP16;

    Params p{};
    for (int r = 0; r < a.n_roles; ++r) {
        const uint8_t* base = static_cast<const uint8_t*>(a.net[r]->pack_tc);
        p.wpack[r] = f16 ? base : base + img;
        p.bias[r] = reinterpret_cast<const float*>(base + 2 * img);
        p.sign[r] = a.sign[r];
        p.out[r] = a.out_logp[r];
    }
    p.s = a.s; p.a = a.a; p.probe = a.probe;
    p.B = a.B; p.n_roles = a.n_roles; p.n_heads = n0->cfg.n_heads; p.depth = n0->cfg.depth;
    p.s_dim = n0->cfg.s_dim; p.a_dim = n0->cfg.a_dim; p.d_in = n0->d_in; p.K0p = n0->d_in_p; p.KT = sc.KT;
    p.NOp = n0->n_out_p; p.H = H;
    p.n_steps = a.n_steps;
    if (const char* dbg = getenv("FRL_TC_DEBUG")) p.debug = atoi(dbg);
    p.delta = (float)(1.0 / a.n_steps);
    for (int k = 0; k < a.n_steps; ++k) p.t_mid[k] = a.t_mid_host[k];
    p.probe_step_stride = a.probe_mode == FRL_PROBE_PER_STEP ? a.B * (long long)p.a_dim : 0;
    p.n_blocks = (int)sc.blk.size();
    for (int i = 0; i < p.n_blocks; ++i) p.blk[i] = sc.blk[i];

    // shared memory carve-up
    uint32_t off = 1024;
    p.off_a = off; off += 2u * H * kRows * 2;
    p.off_x0p = off; off += (uint32_t)p.K0p * kRows * 2;
    p.off_x0t = off; off += (uint32_t)p.KT * kRows * 2;
    p.off_ones = off; off += 2u * kRows * 16;
    p.off_bias = off; off += (uint32_t)(a.n_roles * n0->n_out_p) * 4;
    off = (off + 15) / 16 * 16;
    p.off_state = off; off += 2u * p.a_dim * kRows * 4;
    p.off_rec = off; off += 2u * tc::kMmaWarps * (uint32_t)p.n_blocks * 16;
    off = (off + 1023) / 1024 * 1024;
    p.off_ring = off;
    int max_smem = 0;
    FRL_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, n0->cfg.device));
    int stages = ((int)max_smem - (int)off) / kStageBytes;
    if (stages > 8) stages = 8;
    if (stages < 2) {
        set_error("tensor-core log-density: state/action dimensions too large for shared memory");
        return FRL_ERR_UNSUPPORTED;
    }
    p.n_stages = stages;
    const size_t smem_bytes = (size_t)off + (size_t)stages * kStageBytes;

    int sms = 0;
    FRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, n0->cfg.device));
    const long long n_tiles = ((a.B + kRows - 1) / kRows) * a.n_roles;
    const int grid = (int)std::min<long long>(n_tiles, sms);

    auto run = [&](const Params& q) {
        if (H == 256)
            return f16 ? launch_tc<__half, 256>(q, smem_bytes, grid, stream)
                       : launch_tc<__nv_bfloat16, 256>(q, smem_bytes, grid, stream);
        return f16 ? launch_tc<__half, 128>(q, smem_bytes, grid, stream)
                   : launch_tc<__nv_bfloat16, 128>(q, smem_bytes, grid, stream);
    };
    if (const char* tp = getenv("FRL_TC_TRACE")) {
        // timing experiment: stamp CTA 0's roles with clock64 and dump them (synchronises!)
        unsigned long long* d = nullptr;
        FRL_CUDA(cudaMalloc(&d, 4 * 8192 * sizeof(unsigned long long)));
        FRL_CUDA(cudaMemset(d, 0, 4 * 8192 * sizeof(unsigned long long)));
        p.trace = d;
        p.basis = -1;
        int rc = run(p);
        if (rc != FRL_OK) return rc;
        FRL_CUDA(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(4 * 8192);
        FRL_CUDA(cudaMemcpy(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(d);
        if (FILE* f = fopen(tp, "w")) {
            for (int r = 0; r < 4; ++r)
                for (int i = 0; i < 8192 && h[(size_t)r * 8192 + i]; ++i)
                    fprintf(f, "%d %llx %llu\n", r, h[(size_t)r * 8192 + i] >> 48, h[(size_t)r * 8192 + i] & 0xFFFFFFFFFFFFull);
            fclose(f);
        }
        return FRL_OK;
    }
    if (a.probe_mode == FRL_PROBE_EXACT) {
        for (int i = 0; i < p.a_dim; ++i) {
            p.basis = i;
            p.accumulate = i > 0;
            int rc = run(p);
            if (rc != FRL_OK) return rc;
        }
        return FRL_OK;
    }
    p.basis = -1;
    p.accumulate = 0;
    return run(p);
}

}  // namespace frl
