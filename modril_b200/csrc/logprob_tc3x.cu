// tcgen05 fused log-density integrator with FP32-GRADE products (FRL_PREC_FP16X3): the parity mode on tensor cores.
//
// Every product of the velocity net is evaluated as a 3-term fp16 operand split
//        x W  ~=  x_hi W_hi + x_lo W_hi + x_hi W_lo,     x = x_hi + x_lo,  W = W_hi + W_lo  (fp16 each, 22 bits together)
// accumulated in fp32 by the tensor core.  The dropped x_lo W_lo term is 2^-22 relative.  The weights are packed scaled by
// 2^6 so that W_lo stays a NORMAL fp16 number (|W| ~ 1/16: W_lo would be subnormal, 2^-20 relative instead of 2^-22);
// the epilogues multiply the accumulator by 2^-6 (exact).  Emulation and measurement: the per-sample error against the
// float64 oracle has the distribution of the FP32 FFMA kernel (median 5e-8, same ReLU-kink rate).
//
// Structure: the CTA-pair design of logprob_tc3.cu (cta_group::2, M = 256 over two SMs, accumulator halves, weight ring,
// 16x256b epilogue, DSMEM barrier relays) with ONE 64-sample tile per CTA instead of two: the tile's operand is held as
// two 64 KB images (hi, lo), which is all the shared memory there is.  The tensor pipe stays busy inside a layer because
// every trunk layer is 3x the MMA work of the one-pass kernel while the epilogue of accumulator half 0 runs under the
// MMAs of half 1 (and vice versa across layers).  A weight block (one ring slot, K = 64) holds [bias rows] [W_hi] [W_lo]
// and is read by two concurrent issue streams: [ones x bias rows] [x_hi x W_hi] into the main accumulator and
// [x_lo x W_hi] [x_hi x W_lo] into the small-terms accumulator (see make_schedule for the block order and why).
// Accumulators: the tensor core adds each K = 16 group into the fp32 TMEM accumulator with round-toward-zero, a bias
// that is coherent over all units of a layer (measured: 48 chained MMAs per product put the median error of a stress
// net at 8e-7, ten times the FFMA kernel's, and a numpy emulation of per-MMA truncation reproduces that figure to the
// digit).  So the two small products x_lo W_hi + x_hi W_lo go to a SECOND accumulator (TMEM columns 256..511; their
// truncation errors are 2^-11 of the main one's) and the epilogue adds the two in fp32 (round-to-nearest): the main
// accumulator sees a third of the truncations.
// Reference arithmetic: modril/flowril/pretrain.py:192-214 / :112-134 (fp32 throughout).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "frl_internal.h"
#include "tc3_common.cuh"

namespace frl {

namespace tc3x {

using namespace tcc;
using namespace tc3h;

constexpr int kEpiWarp0 = 6;    // warp 0 producer + TMEM owner, warps 1-5 MMA issuers (leader; main stream 2, small stream 3) / relay (peer: warp 1), 8 epilogue warps
constexpr int kIssuers0 = 2, kIssuers1 = 2;   // issue threads of the main / small-terms stream (up to 2 / 3)
constexpr int kEpiWarps = 8;    // four TMEM lane quarters x two 64-column slices of the accumulator half being drained
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kSlotBytes = 18432;   // [16 bias rows +] W_hi and W_lo, K = 64, of this CTA's 64 output features
constexpr int kMaxSlots = 8;
constexpr int kRows = 128;      // rows of the UMMA operand: 64 samples x (primal, tangent)
constexpr int kSamp = 64;
constexpr int H = 256;
constexpr int kMaxBlocks = 56;
constexpr float kWScale = 64.f;           // weights (and trunk biases) are packed x 2^6 ...
constexpr float kInvW = 1.f / 64.f;       // ... and every accumulator is scaled back in the epilogue (exact)

struct Seg {
    uint8_t a_img;      // A operand image: 0 hi, 1 lo, 2 the ones tile (bias rows)
    uint8_t n_k16;      // K = 16 MMAs of the segment (0: unused)
    uint16_t a_koff;    // first 8-feature chunk of the A operand
    uint16_t b_chunk;   // first 8-feature chunk of the ring slot the B operand starts at
    uint16_t over;      // the first MMA of the segment overwrites its accumulator
};
// wait bits of a block
constexpr int kWaitA = 1;        // << kq: operand columns 64 kq .. 64 kq + 63 written (a_ready[kq])
constexpr int kWaitFree = 16;    // << h:  accumulator half h (main and small) drained (acc_free[h])
// commit bits of a block
constexpr int kCommitD = 1;      // << h: d_full[h]
constexpr int kCommitKa = 4;     // << c: ka_done[c]
struct Blk {
    uint32_t goff, bytes;   // ONE CTA's half of the weight block
    uint16_t nrows;         // N of the UMMA (both CTAs together)
    uint8_t src;            // A operand: 0 activations, 1 layer-0 input
    uint8_t commit, wait;
    uint8_t dhalf;          // accumulator half (TMEM column offset 128 * dhalf; the small-terms accumulator is 256 columns further)
    Seg seg[2][2];          // [stream: 0 main accumulator, 1 small-terms accumulator][up to two segments]
};

struct Params {
    const uint8_t* wpack[2];
    const float* bias[2];       // [depth*H + NOp] fp32 (only the head biases are read here)
    float sign[2];
    float* out[2];
    const float* s;
    const float* a;
    const float* probe;
    long long probe_step_stride;
    long long B;
    int n_roles, n_heads, depth, s_dim, a_dim, K0p, NOp;
    int n_steps, basis, accumulate, n_blocks, n_stages;
    float delta;
    uint32_t rank_stride;        // bytes between the weight images of CTA rank 0 and rank 1
    uint32_t off_a, off_x0, off_ones, off_bias, off_state, off_rec, off_rrec, off_ring;
    uint32_t x0_bytes;           // one image (hi or lo) of the layer-0 operand; 0 when it lives inside the activation tile
    int x0_in_a;                 // wide inputs: see logprob_tc3.cu (the operand is rebuilt inside the activation images)
    const uint8_t* x0s;          // [tile][hi | lo][K0p/8][128][8] images of the state columns (x0_in_a only)
    unsigned long long* trace;   // FRL_TC_TRACE: [4 roles][kTraceCap] {tag, clock} stamps of cluster 0's leader (timing experiments only)
    int trace_step;              // the Euler step (of the first tile) that is recorded
    uint32_t off_trace;          // shared-memory staging of the stamps
    Blk blk[kMaxBlocks];
    float t_mid[kMaxSteps];
};

__device__ __forceinline__ uint32_t pack2_relu_f16(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack2_sat_f16(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 unpack2_f16(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

// (primal c, c+1; tangent c, c+1) pre-activations (x 2^6) -> the hi and lo operand words of the next layer: relu on the
// primal, the tangent (and both lo words) masked by the sign of the primal pre-activation
__device__ __forceinline__ void split4(float zp0, float zp1, float zt0, float zt1, uint32_t& ph, uint32_t& th, uint32_t& pl,
                                       uint32_t& tl) {
    const uint32_t keep = ~neg_mask16x2(__float_as_uint(zp0), __float_as_uint(zp1));
    const float p0 = zp0 * kInvW, p1 = zp1 * kInvW, t0 = zt0 * kInvW, t1 = zt1 * kInvW;
    ph = pack2_relu_f16(p0, p1);
    const uint32_t tr = pack2_sat_f16(t0, t1);
    const float2 pf = unpack2_f16(ph), tf = unpack2_f16(tr);
    pl = pack2_sat_f16(p0 - pf.x, p1 - pf.y) & keep;
    tl = pack2_sat_f16(t0 - tf.x, t1 - tf.y) & keep;
    th = tr & keep;
}
// one fp32 value -> its hi and lo fp16 entries at the same position of the two operand images
__device__ __forceinline__ void put_split(uint8_t* img_hi, uint32_t lo_off, uint32_t off, float v) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    const __half l = __float2half_rn(v - __half2float(h));
    *reinterpret_cast<__half*>(img_hi + off) = h;
    *reinterpret_cast<__half*>(img_hi + lo_off + off) = l;
}

// FRL_TC_TRACE (timing experiments): clock stamps of ONE Euler step of cluster 0's leader, recorded in shared memory
// (a stamp costs ~40 cycles; the global-memory tracer of logprob_tc3.cu distorts the issue threads) and copied out at exit.
constexpr int kTraceCap = 160;   // stamps per role
struct LTracer {
    uint2* buf;
    int n;
    bool on;
    __device__ __forceinline__ void init(uint8_t* base, int role, bool enable) {
        buf = enable ? reinterpret_cast<uint2*>(base) + (size_t)role * kTraceCap : nullptr;
        n = 0;
        on = false;
    }
    __device__ __forceinline__ void arm(bool a) { on = a && buf != nullptr; }
    __device__ __forceinline__ void mark(unsigned tag) {
        if (on && n < kTraceCap) buf[n++] = make_uint2(tag, (unsigned)clock());
    }
};

template <bool WIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) logprob_tc3x_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t bar0 = smem_u32(smem);
    auto w_full = [&](int s) { return bar0 + 8u * s; };                  // own half of the block has landed (tx)
    auto w_empty = [&](int s) { return bar0 + 8u * (kMaxSlots + s); };   // both streams' MMAs on the slot completed
    auto a_ready = [&](int kq) { return bar0 + 8u * (2 * kMaxSlots + kq); };         // leader only: operand columns 64 kq .. written
    auto acc_free = [&](int h) { return bar0 + 8u * (2 * kMaxSlots + 4 + h); };      // leader only: accumulator half h drained
    auto d_full = [&](int h) { return bar0 + 8u * (2 * kMaxSlots + 6 + h); };        // both streams' MMAs into half h completed
    // the layer's MMAs no longer read operand columns 64 c .. 64 c + 63 (c = 0, 1): the epilogue of accumulator half 0 may
    // overwrite them (the activation images are updated in place while the MMAs of half 1 are still running)
    auto ka_done = [&](int c) { return bar0 + 8u * (2 * kMaxSlots + 8 + c); };
    auto hand = [&](int st, int x) { return bar0 + 8u * (2 * kMaxSlots + 10 + st * 3 + x); };   // leader only: issue batons
    const uint32_t x0_full = bar0 + 8u * (2 * kMaxSlots + 16);   // wide inputs: the operand images have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);

    constexpr uint32_t a_tile_bytes = (uint32_t)H * kRows * 2;   // 64 KB per image
    constexpr uint32_t kChunk = kRows * 16;                       // bytes of one 8-feature chunk of an operand image
    uint8_t* sA = smem + p.off_a;         // [hi | lo][H/8][128][8]
    uint8_t* sX0 = smem + p.off_x0;       // [hi | lo][K0p/8][128][8]
    uint8_t* sOnes = smem + p.off_ones;   // [2][128][8]: feature 0..2 = 1.0 on the primal rows (bias hi/mid/lo terms)
    float* sBias = reinterpret_cast<float*>(smem + p.off_bias);     // [roles][NOp] head biases (trunk biases ride in the MMA)
    float* sState = reinterpret_cast<float*>(smem + p.off_state);   // a [a_dim][64], probe [a_dim][64]
    uint4* sRec = reinterpret_cast<uint4*>(smem + p.off_rec);       // [n_blocks][5]: block header + 2 streams x 2 segments
    uint2* sRingRec = reinterpret_cast<uint2*>(smem + p.off_rrec);  // [n_blocks] {offset, bytes} of the producer's loads
    uint8_t* sRing = smem + p.off_ring;
    uint8_t* sTrace = smem + p.off_trace;
    const bool tracing = p.trace != nullptr && blockIdx.x < 2;   // cluster 0: the leader's roles, and the peer's epilogue warp 0 (as role 4)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();          // 0 = leader (issues the MMAs of the pair)
    const long long cluster_id = cluster_id_x();
    const long long n_clusters = gridDim.x / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(w_full(s), cta_rank == 0 ? 2 : 1);   // leader: own producer + the peer's relay
            mbar_init(w_empty(s), 2);                      // one commit per stream
        }
        for (int kq = 0; kq < 4; ++kq) mbar_init(a_ready(kq), 2 * kEpiWarps);   // every epilogue warp of both CTAs writes a part of every operand quarter
        for (int h = 0; h < 2; ++h) {
            mbar_init(acc_free(h), 2 * kEpiWarps);   // every epilogue warp of both CTAs
            mbar_init(d_full(h), 2);
            mbar_init(ka_done(h), 2);
            for (int x = 0; x < 3; ++x) mbar_init(hand(h, x), 1);
        }
        mbar_init(x0_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.NOp * p.n_roles; i += kThreads) sBias[i] = p.bias[i / p.NOp][p.depth * H + i % p.NOp];
    for (int i = threadIdx.x; i < 2 * kRows; i += kThreads) {
        const uint32_t one = 0x3C00u;   // fp16 1.0; primal rows are rows 0..7 of every 16-row group
        reinterpret_cast<uint4*>(sOnes)[i] = (i < kRows && (i & 8) == 0) ? make_uint4(one | (one << 16), one, 0, 0)
                                                                           : make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < (int)(2 * p.x0_bytes / 16); i += kThreads)
        reinterpret_cast<uint4*>(sX0)[i] = make_uint4(0, 0, 0, 0);
    for (int b = threadIdx.x; b < p.n_blocks; b += kThreads) {
        const Blk& k = p.blk[b];
        const uint32_t nhalf = k.nrows >> 1;
        // M = 256 over the pair, N = nrows (each CTA holds nrows / 2 rows of B); fp16 x fp16 -> fp32
        const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 4) << 24) | ((uint32_t)(k.nrows >> 3) << 17);
        sRec[5 * b] = make_uint4(idesc, (uint32_t)k.nrows | ((uint32_t)k.commit << 16), (uint32_t)k.wait, (uint32_t)k.dhalf * 128);
        for (int st = 0; st < 2; ++st)
            for (int sg = 0; sg < 2; ++sg) {
                const Seg& g = k.seg[st][sg];
                uint32_t src;
                if (g.a_img == 2) src = smem_u32(sOnes);
                else if (k.src == 0 || WIDE) src = smem_u32(sA) + g.a_img * a_tile_bytes;
                else src = smem_u32(sX0) + g.a_img * p.x0_bytes;
                const uint32_t a_lo = (((src + (uint32_t)g.a_koff * kChunk) >> 4) & 0x3FFF) | ((kChunk >> 4) << 16);
                sRec[5 * b + 1 + 2 * st + sg] = make_uint4(a_lo, (uint32_t)g.b_chunk * nhalf, g.n_k16, g.over);
            }
        sRingRec[b] = make_uint2(k.goff, k.bytes);
    }
    if (tracing)
        for (int i = threadIdx.x; i < 4 * kTraceCap; i += kThreads) reinterpret_cast<uint2*>(sTrace)[i] = make_uint2(0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // barrier inits of both CTAs are visible before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // work decomposition: a "duo" is two consecutive 64-sample tiles of one role, one per CTA of the pair
    const long long tiles_per_role = (p.B + kSamp - 1) / kSamp;
    const long long n_duos = ((tiles_per_role + 1) / 2) * p.n_roles;
    auto arrive_leader = [&](uint32_t bar) {
        if (cta_rank == 0) mbar_arrive(bar);
        else mbar_arrive_remote(bar, 0);
    };

    if (warp == 0) {
        // ------------------------------- producer: this CTA's half of every weight block --------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            LTracer tr; tr.init(sTrace, 3, tracing && blockIdx.x == 0);
            for (long long duo = cluster_id; duo < n_duos; duo += n_clusters) {
                const uint8_t* wbase = p.wpack[duo % p.n_roles] + (size_t)cta_rank * p.rank_stride;
                for (int step = 0; step < p.n_steps; ++step) {
                    tr.arm(duo == cluster_id && step == p.trace_step);
                    for (int i = 0; i < p.n_blocks; ++i) {
                        const uint2 k = sRingRec[i];
                        mbar_wait(w_empty(stage), phase ^ 1);
                        tr.mark(0x200 + i);
                        mbar_expect_tx(w_full(stage), k.y);
                        bulk_g2s(smem_u32(sRing + (size_t)stage * kSlotBytes), wbase + k.x, k.y, w_full(stage));
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && cta_rank == 1) {
        // ------------------------------- peer: tell the leader when my half of a block has landed -----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long duo = cluster_id; duo < n_duos; duo += n_clusters)
                for (int step = 0; step < p.n_steps; ++step)
                    for (int i = 0; i < p.n_blocks; ++i) {
                        mbar_wait(w_full(stage), phase);
                        mbar_arrive_remote(w_full(stage), 0);
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp >= 1 && warp <= 5 && cta_rank == 0) {
        // ------------------------------- leader: the MMA issuers (for both CTAs) ----------------------------------
        // Two STREAMS run concurrently on every weight block: stream 0 issues the products that go to the main accumulator
        // ([ones x bias rows], x_hi W_hi), stream 1 those of the small-terms accumulator (x_lo W_hi, x_hi W_lo).  The streams
        // write different accumulators, so their relative order in the pipe does not matter (deterministic).  A thread pays
        // ~60-80 cycles per tcgen05.mma, ~250 per tcgen05.commit and ~100-500 per barrier wait, against 64 pipe cycles per
        // N = 128 MMA: inside a stream the threads take the blocks round-robin and pass a baton once their MMAs are
        // issued, so the commits and waits of one run in the shadow of the others' MMAs while the baton keeps the accumulation
        // order of the stream fixed.  EVERY issuer observes every phase of every readiness barrier, in block order: a thread
        // that asked for phase n before phase n-1 had completed would see the parity of phase n-1's predecessor and pass.
        // (Measured: three threads in the small-terms stream, or polling with mbarrier.test_wait instead of the suspending
        // try_wait, are both SLOWER -- the kernel is bound by shared-memory operand traffic, which polling adds to.)
        if (lane == 0) {
            const int st = warp <= kIssuers0 ? 0 : 1;
            const int n_x = st == 0 ? kIssuers0 : kIssuers1, x = st == 0 ? warp - 1 : warp - 1 - kIssuers0;
            uint32_t par_a = 0, par_f = 0, par_h = 0;   // parity bits of a_ready[0..3] / acc_free[0..1]
            long long seq = 0;
            int turn = 0;                               // seq % n_x
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t ring_lo = smem_u32(sRing) >> 4;
            constexpr uint32_t kDescHi = (uint32_t)(128 >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
            LTracer tr; tr.init(sTrace, 1 + st, tracing && blockIdx.x == 0 && x == 0);           // roles 1 / 2: first issuer of stream 0 / 1
            for (long long duo = cluster_id; duo < n_duos; duo += n_clusters) {
                for (int step = 0; step < p.n_steps; ++step) {
                    tr.arm(duo == cluster_id && step == p.trace_step);
                    for (int b = 0; b < p.n_blocks; ++b, ++seq) {
                        const uint4 r0 = sRec[5 * b];
                        const bool mine = turn == x;
                        if (r0.z) {
#pragma unroll
                            for (int kq = 0; kq < 4; ++kq)
                                if (r0.z & (kWaitA << kq)) {
                                    mbar_wait_cluster(a_ready(kq), (par_a >> kq) & 1);
                                    par_a ^= 1u << kq;
                                }
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                if (r0.z & (kWaitFree << h)) {
                                    mbar_wait_cluster(acc_free(h), (par_f >> h) & 1);
                                    par_f ^= 1u << h;
                                }
                        }
                        if (mine) {
                            tr.mark(0x600 + b);   // readiness of the block's operands / accumulator observed
                            if (seq > 0) { mbar_wait(hand(st, x), par_h); par_h ^= 1; }
                            mbar_wait(w_full(stage), phase);   // both halves of the block are in the two CTAs' rings
                            tr.mark(0x300 + b);
                            tc_fence_after();
                            const uint32_t nhalf = (r0.y & 0xFFFF) >> 1;
                            const uint32_t b_base = ((ring_lo + (uint32_t)stage * (kSlotBytes >> 4)) & 0x3FFF) | (nhalf << 16);
                            const uint32_t d_addr = tmem + r0.w + (st ? 256u : 0u);
#pragma unroll
                            for (int sg = 0; sg < 2; ++sg) {
                                const uint4 r = sRec[5 * b + 1 + 2 * st + sg];
                                uint32_t a_lo = r.x, b_lo = b_base + r.y;
                                uint32_t acc = r.w ? 0u : 1u;
                                for (uint32_t u = 0; u < r.z; ++u) {
                                    umma2_f16(d_addr, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, r0.x, acc);
                                    acc = 1u;
                                    a_lo += 2 * (kChunk >> 4);
                                    b_lo += 2 * nhalf;
                                }
                            }
                            mbar_arrive(hand(st, x + 1 == n_x ? 0 : x + 1));
                            tr.mark(0x700 + b);
                            umma2_commit_both(w_empty(stage));
                            const uint32_t cm = (r0.y >> 16) & 0xFF;
                            if (cm & kCommitKa) umma2_commit_both(ka_done(0));
                            if (cm & (kCommitKa << 1)) umma2_commit_both(ka_done(1));
                            if (cm & kCommitD) umma2_commit_both(d_full(0));
                            if (cm & (kCommitD << 1)) umma2_commit_both(d_full(1));
                            tr.mark(0x400 + b);
                        }
                        if (++turn == n_x) turn = 0;
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------ epilogue ------------------------------------------------
        // Every warp works on BOTH accumulator halves in turn: warp (q, cs) owns TMEM lane quarter q and the 64-column
        // slice cs of each half, i.e. operand columns 128 h + 64 cs .. + 63 = K quarter 2 h + cs of the next layer.  The heads
        // (<= 64 columns) are split by 16-row group instead.
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;                 // TMEM lane quarter: operand rows 32q .. 32q+31 = samples 16q .. 16q+15
        const int cs = ew >> 2;                 // trunk: 64-column slice of the half;  heads: 16-row group
        const int et = ew * 32 + lane;
        const int tq = lane & 3, tr8 = lane >> 2;   // position inside the 16x256b fragment: column pair, row
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int t_col = p.a_dim + p.s_dim;
        uint32_t par_d[2] = {0, 0}, par_k = 0, par_x0 = 0;
        float log_det = 0.f;   // of this warp's 16-row group (heads: group cs of lane quarter q)
        const uint32_t lo_off = WIDE ? a_tile_bytes : p.x0_bytes;   // hi -> lo image of the layer-0 operand
        uint8_t* const x0 = WIDE ? sA : sX0;
        float* const sA_state = sState;
        float* const sP_state = sState + p.a_dim * kSamp;
        const uint32_t x0_img_bytes = (uint32_t)p.K0p * kRows * 2;
        LTracer tr; tr.init(sTrace, 0, tracing && lane == 0 && ew == 0);

        // wide inputs: bulk copies of a tile's state-column images (hi, lo) into the activation images (one thread) ...
        auto x0_fetch = [&](long long tile) {
            mbar_expect_tx(x0_full, 2 * x0_img_bytes);
            const uint8_t* src = p.x0s + (size_t)tile * 2 * x0_img_bytes;
            for (int im = 0; im < 2; ++im)
                for (uint32_t off = 0; off < x0_img_bytes; off += 16384u)
                    bulk_g2s(smem_u32(sA) + im * a_tile_bytes + off, src + (size_t)im * x0_img_bytes + off,
                             min(16384u, x0_img_bytes - off), x0_full);
        };
        // ... and, once they have landed, the a / t (primal rows) and probe (tangent rows) entries of a sample
        auto x0_sample = [&](int srow, int j0, int jstep, bool write_t, float tval) {
            const int rowp = row_primal(srow);
            for (int j = j0; j < p.a_dim; j += jstep) {
                put_split(x0, lo_off, (j >> 3) * kChunk + rowp * 16 + (j & 7) * 2, sA_state[j * kSamp + srow]);
                put_split(x0, lo_off, (j >> 3) * kChunk + (rowp + 8) * 16 + (j & 7) * 2, sP_state[j * kSamp + srow]);
            }
            if (write_t) put_split(x0, lo_off, (t_col >> 3) * kChunk + rowp * 16 + (t_col & 7) * 2, tval);
        };
        // the layer-0 operand is complete and the head accumulators are drained: every barrier the issuers wait on gets one
        // phase
        auto arrive_all = [&]() {
            if (lane == 0) {
                arrive_leader(acc_free(0));
                arrive_leader(acc_free(1));
                for (int kq = 0; kq < 4; ++kq) arrive_leader(a_ready(kq));
            }
        };

        for (long long duo = cluster_id; duo < n_duos; duo += n_clusters) {
            const int role = (int)(duo % p.n_roles);
            const long long tile = 2 * (duo / p.n_roles) + cta_rank;   // padding tiles run masked
            const long long row0 = tile * kSamp;
            const float sign = p.sign[role];
            const float* bh = sBias + role * p.NOp;

            // ---- tile init: fp32 state + layer-0 operand rows (X0 was zeroed once; padding / tangent s,t columns stay 0) ----
            for (int idx = et; idx < kSamp * p.a_dim; idx += kEpiWarps * 32) {
                const int r = idx / p.a_dim, c = idx % p.a_dim;
                const long long g = row0 + r;
                const float v = g < p.B ? p.a[g * p.a_dim + c] : 0.f;
                float pr = 0.f;
                if (g < p.B) pr = p.basis >= 0 ? (c == p.basis ? 1.f : 0.f) : p.probe[g * p.a_dim + c];
                sA_state[c * kSamp + r] = v;
                sP_state[c * kSamp + r] = pr;
                if (!WIDE) {
                    put_split(x0, lo_off, (c >> 3) * kChunk + row_primal(r) * 16 + (c & 7) * 2, v);
                    put_split(x0, lo_off, (c >> 3) * kChunk + (row_primal(r) + 8) * 16 + (c & 7) * 2, pr);
                }
            }
            if (!WIDE) {
                for (int idx = et; idx < kSamp * p.s_dim; idx += kEpiWarps * 32) {
                    const int r = idx / p.s_dim, c = idx % p.s_dim;
                    const long long g = row0 + r;
                    const int kk = p.a_dim + c;
                    put_split(x0, lo_off, (kk >> 3) * kChunk + row_primal(r) * 16 + (kk & 7) * 2, g < p.B ? p.s[g * p.s_dim + c] : 0.f);
                }
                if (et < kSamp) put_split(x0, lo_off, (t_col >> 3) * kChunk + row_primal(et) * 16 + (t_col & 7) * 2, p.t_mid[0]);
            } else {
                named_sync(1, kEpiWarps * 32);   // the fp32 state of the tile is complete
                if (et == 0) x0_fetch(tile);
                mbar_wait(x0_full, par_x0);
                par_x0 ^= 1;
                if (et < kSamp) x0_sample(et, 0, 1, true, p.t_mid[0]);
            }
            log_det = 0.f;
            fence_async_smem();
            named_sync(1, kEpiWarps * 32);
            arrive_all();

            for (int step = 0; step < p.n_steps; ++step) {
                tr.arm(duo == cluster_id && step == p.trace_step);
                // ---- trunk layers: accumulator half h (main + small terms) -> columns 128 h .. 128 h + 127 of the operand images ----
                for (int l = 0; l < p.depth; ++l) {
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        tr.mark(0x100 + l * 2 + h);
                        mbar_wait(d_full(h), par_d[h]);
                        par_d[h] ^= 1;
                        tc_fence_after();
                        tr.mark(0x200 + l * 2 + h);
                        // z = main + small accumulator (fp32, round-to-nearest) of both 16-row groups of this warp's lane quarter.
                        // The 128 columns of the half are the operand quarters 2h and 2h + 1; ALL eight warps take 32 columns of
                        // quarter 2h first and of quarter 2h + 1 afterwards (slice cs of each), so that the next layer's MMAs over
                        // quarter 2h start while the second quarter is still being written.
                        float z[2][2][16];   // [quarter part][row group][4 x (primal c, c+1, tangent c, c+1)]
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
                            const uint32_t colp = (uint32_t)(h * 128 + part * 64 + cs * 32);
#pragma unroll
                            for (int grp = 0; grp < 2; ++grp) {
                                uint32_t zm[16], zs[16];
                                const uint32_t tb = lane_addr + ((uint32_t)(grp * 16) << 16) + colp;
                                tmem_ld_16x256b_x4(tb, zm);
                                tmem_ld_16x256b_x4(tb + 256, zs);
                                tmem_ld_wait_dep16(zm);
                                tmem_ld_wait_dep16(zs);
#pragma unroll
                                for (int i = 0; i < 16; ++i) z[part][grp][i] = __uint_as_float(zm[i]) + __uint_as_float(zs[i]);
                            }
                        }
                        // the accumulator half is in registers: the next layer's MMAs may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_leader(acc_free(h));
                        tr.mark(0x500 + l * 2 + h);
#pragma unroll
                        for (int part = 0; part < 2; ++part) {
                            if (h == 0) {   // in-place update: wait until half 1's MMAs are past this operand quarter
                                mbar_wait(ka_done(part), (par_k >> part) & 1);
                                par_k ^= 1u << part;
                            }
                            const uint32_t colp = (uint32_t)(h * 128 + part * 64 + cs * 32);
#pragma unroll
                            for (int grp = 0; grp < 2; ++grp) {
                                // lane i addresses row i%8 of matrix i/8: matrices 0/1 = primal/tangent rows of chunk k,
                                // matrices 2/3 = the same of chunk k+1
                                const uint32_t rowa = (uint32_t)(q * 32 + grp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8);
                                uint32_t addr = smem_u32(sA) + (uint32_t)((colp >> 3) + (lane >> 4)) * kChunk + rowa * 16;
#pragma unroll
                                for (int k = 0; k < 4; k += 2) {
                                    uint32_t wh[4], wl[4];
#pragma unroll
                                    for (int e = 0; e < 2; ++e)
                                        split4(z[part][grp][4 * (k + e)], z[part][grp][4 * (k + e) + 1], z[part][grp][4 * (k + e) + 2],
                                               z[part][grp][4 * (k + e) + 3], wh[2 * e], wh[2 * e + 1], wl[2 * e], wl[2 * e + 1]);
                                    stmatrix_x4(addr, wh[0], wh[1], wh[2], wh[3]);
                                    stmatrix_x4(addr + a_tile_bytes, wl[0], wl[1], wl[2], wl[3]);
                                    addr += 2 * kChunk;
                                }
                            }
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0) arrive_leader(a_ready(2 * h + part));
                        }
                        tr.mark(0x300 + l * 2 + h);
                    }
                }
                // ---- heads: tanh, divergence, Euler update (pretrain.py:205-211) ----
                tr.mark(0x1F0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // the last heads block commits both barriers: every warp follows both
                    mbar_wait(d_full(h), par_d[h]);
                    par_d[h] ^= 1;
                }
                tc_fence_after();
                tr.mark(0x2F0);
                // wide inputs: the activation images are free (their last reader, the heads MMA, is done): fetch the next
                // step's layer-0 operand images while the heads are evaluated
                if (WIDE && step + 1 < p.n_steps && et == 0) x0_fetch(tile);
                {
                    // the heads have only NOp <= 64 columns: the two warps of a lane quarter split its two 16-row groups
                    const int grp = cs;
                    const int srow = q * 16 + grp * 8 + tr8;             // sample inside the tile
                    const int rowp = q * 32 + grp * 16 + tr8;
                    const long long g = row0 + srow;
                    const uint32_t hb = lane_addr + ((uint32_t)(grp * 16) << 16);
                    float div = 0.f;
                    for (int g16 = 0; g16 < p.NOp; g16 += 16) {
                        uint32_t vv[8], ww[8];
                        tmem_ld_16x256b_x2(hb + (uint32_t)g16, vv);
                        tmem_ld_16x256b_x2(hb + 256u + (uint32_t)g16, ww);
                        tmem_ld_wait_dep8(vv);
                        tmem_ld_wait_dep8(ww);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int c = g16 + 8 * k + 2 * tq;   // this thread's column pair (c, c+1): primal [4k], [4k+1]; tangent [4k+2], [4k+3]
                            const float o0 = (__uint_as_float(vv[4 * k]) + __uint_as_float(ww[4 * k])) * kInvW;
                            const float o1 = (__uint_as_float(vv[4 * k + 1]) + __uint_as_float(ww[4 * k + 1])) * kInvW;
                            const float d0 = (__uint_as_float(vv[4 * k + 2]) + __uint_as_float(ww[4 * k + 2])) * kInvW;
                            const float d1 = (__uint_as_float(vv[4 * k + 3]) + __uint_as_float(ww[4 * k + 3])) * kInvW;
                            if (p.n_heads == 2) {
                                const int j = c >> 1;             // columns (2j, 2j+1) = (v_c, rho) of action dim j
                                if (j < p.a_dim) {
                                    const float th = tanhf(o1 + bh[c + 1]);
                                    const float v = o0 + bh[c] + sign * th;
                                    const float dv = d0 + sign * (1.f - th * th) * d1;
                                    div = fmaf(dv, sP_state[j * kSamp + srow], div);
                                    const float an = sA_state[j * kSamp + srow] + v * p.delta;
                                    sA_state[j * kSamp + srow] = an;
                                    if (!WIDE) put_split(x0, lo_off, (j >> 3) * kChunk + rowp * 16 + (j & 7) * 2, an);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int j = c + e;
                                    if (j < p.a_dim) {
                                        const float v = (e ? o1 : o0) + bh[j];
                                        div = fmaf(e ? d1 : d0, sP_state[j * kSamp + srow], div);
                                        const float an = sA_state[j * kSamp + srow] + v * p.delta;
                                        sA_state[j * kSamp + srow] = an;
                                        if (!WIDE) put_split(x0, lo_off, (j >> 3) * kChunk + rowp * 16 + (j & 7) * 2, an);
                                    }
                                }
                            }
                        }
                    }
                    // the four threads of a sample hold different action dims: sum their divergence terms
                    div += __shfl_xor_sync(0xffffffffu, div, 1);
                    div += __shfl_xor_sync(0xffffffffu, div, 2);
                    log_det -= p.delta * div;
                    __syncwarp();   // the state of this sample (written by its four threads) is complete
                    if (step + 1 < p.n_steps) {
                        if (tq == 0 && !WIDE)
                            put_split(x0, lo_off, (t_col >> 3) * kChunk + rowp * 16 + (t_col & 7) * 2, p.t_mid[step + 1]);
                        if (p.probe_step_stride != 0) {
                            for (int j = tq; j < p.a_dim; j += 4) {
                                const float pr = g < p.B ? p.probe[(step + 1) * p.probe_step_stride + g * p.a_dim + j] : 0.f;
                                sP_state[j * kSamp + srow] = pr;
                                if (!WIDE) put_split(x0, lo_off, (j >> 3) * kChunk + (rowp + 8) * 16 + (j & 7) * 2, pr);
                            }
                        }
                    } else if (tq == 0 && g < p.B) {
                        if (p.accumulate) {
                            p.out[role][g] += log_det;
                        } else {
                            float prior = 0.f;
                            for (int j = 0; j < p.a_dim; ++j) {
                                const float xx = sA_state[j * kSamp + srow];
                                prior += -0.5f * xx * xx - 0.91893853320467274178f;
                            }
                            p.out[role][g] = prior + log_det;
                        }
                    }
                    if (step + 1 < p.n_steps) {
                        if (WIDE) {   // the images have landed: this thread's entries of its sample (a, probe; t by the first of the four)
                            __syncwarp();
                            mbar_wait(x0_full, par_x0);
                            par_x0 ^= 1;
                            x0_sample(srow, tq, 4, tq == 0, p.t_mid[step + 1]);
                        }
                        tc_fence_before();
                        fence_async_smem();
                        __syncwarp();
                        arrive_all();
                    }
                }
                tr.mark(0x3F0);
            }
            // the next tile's init rewrites X0 / state: every epilogue warp must be done with this tile first
            tc_fence_before();
            named_sync(1, kEpiWarps * 32);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tracing)
        for (int i = threadIdx.x; i < 4 * kTraceCap; i += kThreads) {
            const uint2 v = reinterpret_cast<const uint2*>(sTrace)[i];
            p.trace[(size_t)blockIdx.x * 4 * kTraceCap + i] = ((unsigned long long)v.x << 32) | v.y;
        }
    cluster_sync_all();   // the leader's last MMAs read the peer's shared memory: nobody leaves or frees TMEM early
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// Wide inputs: hi and lo operand images of the state columns, one pair per 64-sample tile ([tile][hi | lo][K0p/8][128][8];
// primal rows carry s in columns [a_dim, a_dim + s_dim), everything else is zero).  Built once per launch.
__global__ void tc3x_state_image_kernel(const float* __restrict__ s, uint8_t* __restrict__ img, long long B, long long n_tiles,
                                        int s_dim, int a_dim, int K0p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int chunks = K0p / 8;
    if (i >= n_tiles * chunks * kRows) return;
    const int rowi = (int)(i % kRows);
    const int chunk = (int)((i / kRows) % chunks);
    const long long tile = i / ((long long)kRows * chunks);
    uint32_t wh[4] = {0, 0, 0, 0}, wl[4] = {0, 0, 0, 0};
    if ((rowi & 8) == 0) {
        const long long g = tile * kSamp + (((rowi >> 4) << 3) | (rowi & 7));
        if (g < B) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = chunk * 8 + e - a_dim;
                if (c >= 0 && c < s_dim) {
                    const float v = s[g * s_dim + c];
                    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
                    const __half l = __float2half_rn(v - __half2float(h));
                    wh[e >> 1] |= (uint32_t)__half_as_ushort(h) << (16 * (e & 1));
                    wl[e >> 1] |= (uint32_t)__half_as_ushort(l) << (16 * (e & 1));
                }
            }
        }
    }
    const size_t img_bytes = (size_t)K0p * kRows * 2;
    const size_t o = (size_t)tile * 2 * img_bytes + ((size_t)chunk * kRows + rowi) * 16;
    *reinterpret_cast<uint4*>(img + o) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    *reinterpret_cast<uint4*>(img + o + img_bytes) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

// Block stream of one Euler step (shared by the packer and the kernel).  One block = one ring slot of each CTA = one
// w_empty commit.  Blk::goff / Blk::bytes describe ONE CTA's half.
struct Schedule {
    std::vector<Blk> blk;
    std::vector<tc::PackBlk> pack[2];   // per CTA rank
    uint32_t image_bytes = 0;           // of one rank's image
};

static bool supported(const frl_net* net) {
    return net->cfg.hidden == H && net->d_in_p <= 192 && net->n_out_p <= 64;
}
static int block_count(const frl_net* net, bool wide) {
    const int l0 = wide ? 2 * ((net->d_in_p + 63) / 64) : 2;
    return l0 + 8 * (net->cfg.depth - 1) + 2;
}
static uint32_t front_bytes(const frl_net* net, int n_roles, bool wide) {
    uint32_t off = 1024;
    off += 2u * H * kRows * 2;
    if (!wide) off += 2u * (uint32_t)net->d_in_p * kRows * 2;
    off += 2u * kRows * 16;
    off += (uint32_t)(n_roles * net->n_out_p) * 4;
    off = (off + 15) / 16 * 16;
    off += (2u * net->cfg.a_dim * kSamp) * 4;
    off = (off + 15) / 16 * 16;
    off += (uint32_t)block_count(net, wide) * (5u * 16 + 8);
    return (off + 1023) / 1024 * 1024;
}
// 0: the layer-0 operand has its own tiles; 1: it is rebuilt inside the activation images (inputs wider than 32 columns,
// or when that is what leaves room for four ring slots); -1: no plan fits
static int choose_layout(const frl_net* net) {
    const uint32_t max_smem = 232448;   // sm_100a opt-in maximum of dynamic shared memory per CTA
    if (net->d_in_p <= 32 && front_bytes(net, 2, false) + 4u * kSlotBytes <= max_smem) return 0;
    if (front_bytes(net, 2, true) + 4u * kSlotBytes <= max_smem) return 1;
    return -1;
}

// One Euler step as a list of weight blocks.  A block is one ring slot: [16 bias rows] [W_hi piece] [W_lo piece] of one
// (layer, accumulator half, K piece), read by both streams.  Trunk layers run in K = 64 pieces ("K quarters" kq = the 64
// operand columns one epilogue warp slice writes) in the order
//        (h0,k0) (h0,k1) (h1,k0) (h0,k2) (h0,k3) (h1,k1) (h1,k2) (h1,k3)
// chosen against the dependency chain of a single tile: the previous layer's epilogue of half 1 (operand quarters k2, k3)
// ends ~2000 cycles after that layer's last MMA, and three pieces (2300 pipe cycles) need only k0, k1; half 0 completes
// after the fifth piece, so its own epilogue (quarters k0, k1 of the NEXT layer) runs under the last three.
static Schedule make_schedule(const frl_net* net) {
    Schedule sc;
    const int depth = net->cfg.depth, K0p = net->d_in_p, NOp = net->n_out_p;
    const bool wide = choose_layout(net) == 1;
    uint32_t goff = 0;
    // k0 < 0: no piece.  The block's slot holds [bias16 if `bias`] [W_hi k0..k0+klen) [W_lo k0..k0+klen)].
    auto add = [&](int layer, int nrows, int n_base, int src, bool bias, int k0, int klen, bool open_small, bool open_main,
                   int commit, int wait, int dhalf) {
        Blk b{};
        b.goff = goff; b.nrows = (uint16_t)nrows; b.src = (uint8_t)src; b.commit = (uint8_t)commit;
        b.wait = (uint8_t)wait; b.dhalf = (uint8_t)dhalf;
        auto piece = [&](int part, int pk0, int pklen) {
            for (int r = 0; r < 2; ++r) {
                tc::PackBlk pb{goff, nrows / 2, pklen, n_base + r * (nrows / 2), pk0, layer};
                pb.part = part;
                sc.pack[r].push_back(pb);
            }
            goff += (uint32_t)(nrows / 2) * pklen * 2;
        };
        const int nk = klen / 16, c_hi = bias ? 2 : 0, c_lo = c_hi + klen / 8;
        if (bias) piece(0, -1, 16);
        piece(1, k0, klen);
        piece(2, k0, klen);
        int m = 0;
        if (bias) b.seg[0][m++] = Seg{2, 1, 0, 0, 1};                                              // ones x bias rows opens the main accumulator
        b.seg[0][m++] = Seg{0, (uint8_t)nk, (uint16_t)(k0 / 8), (uint16_t)c_hi, (uint16_t)((open_main && !bias) ? 1 : 0)};   // x_hi W_hi
        b.seg[1][0] = Seg{1, (uint8_t)nk, (uint16_t)(k0 / 8), (uint16_t)c_hi, (uint16_t)(open_small ? 1 : 0)};              // x_lo W_hi
        b.seg[1][1] = Seg{0, (uint8_t)nk, (uint16_t)(k0 / 8), (uint16_t)c_lo, 0};                                           // x_hi W_lo
        b.bytes = goff - b.goff;
        sc.blk.push_back(b);
    };
    const int wait_all_a = kWaitA * 15;
    // layer 0: the whole (narrow) input per accumulator half, or K = 64 pieces of a wide one
    for (int h = 0; h < 2; ++h) {
        const int step = wide ? 64 : K0p;
        for (int k0 = 0; k0 < K0p; k0 += step) {
            const int klen = std::min(step, K0p - k0);
            const bool first = k0 == 0, last = k0 + klen >= K0p;
            add(0, 128, 128 * h, 1, first, k0, klen, first, false,
                (last ? kCommitD << h : 0) | ((last && h == 1) ? kCommitKa * 3 : 0),
                first ? (kWaitFree << h) | (h == 0 ? wait_all_a : 0) : 0, h);
        }
    }
    // trunk layers 1 .. depth-1
    static const int order[8][2] = {{0, 0}, {0, 1}, {1, 0}, {0, 2}, {0, 3}, {1, 1}, {1, 2}, {1, 3}};
    for (int l = 1; l < depth; ++l)
        for (int i = 0; i < 8; ++i) {
            const int h = order[i][0], kq = order[i][1];
            int wait = 0, commit = 0;
            if (h == 0) wait |= kWaitA << kq;              // first reader of operand quarter kq
            if (kq == 0) wait |= kWaitFree << h;           // first MMAs into accumulator half h
            if (kq == 3) commit |= kCommitD << h;          // half h complete
            if (h == 1 && kq < 2) commit |= kCommitKa << kq;   // last reader of operand quarter kq (kq = 0, 1)
            add(l, 128, 128 * h, 0, kq == 0, 64 * kq, 64, kq == 0, false, commit, wait, h);
        }
    // heads: two K = 128 blocks [W_hi | W_lo]; biases are added in the epilogue in fp32
    for (int k0 = 0; k0 < H; k0 += 128) {
        const bool first = k0 == 0;
        add(depth, NOp, 0, 0, false, k0, 128, first, first, first ? 0 : kCommitD * 3,
            (first ? kWaitA * 3 | kWaitFree : kWaitA * 12 | (kWaitFree << 1)), 0);
    }
    sc.image_bytes = goff;
    return sc;
}

static size_t rank_image_stride(const frl_net* net, uint32_t image_bytes) {
    const size_t img = ((size_t)image_bytes + 255) / 256 * 256;
    const size_t nb = (size_t)net->cfg.depth * net->cfg.hidden + net->n_out_p;
    return (2 * img + nb * sizeof(float) + 255) / 256 * 256;
}

}  // namespace tc3x

bool tc3x_supported(const frl_net* net) { return tc3x::supported(net) && tc3x::choose_layout(net) >= 0; }

int tc3x_pack(frl_net* net, const float* params_dev, cudaStream_t stream) {
    using namespace tc3x;
    if (!tc3x_supported(net)) return FRL_OK;
    Schedule sc = make_schedule(net);
    if ((int)sc.blk.size() > kMaxBlocks || (int)sc.pack[0].size() > tc::kMaxBlocks) return FRL_OK;
    const size_t stride = rank_image_stride(net, sc.image_bytes);
    if (!net->pack_tc3x) FRL_CUDA(cudaMalloc(&net->pack_tc3x, 2 * stride));
    for (int r = 0; r < 2; ++r) {
        void* dst = static_cast<uint8_t*>(net->pack_tc3x) + r * stride;
        int rc = tc_pack_images(net, sc.pack[r], sc.image_bytes, &dst, params_dev, stream, kWScale);
        if (rc != FRL_OK) return rc;
    }
    return FRL_OK;
}

template <bool WIDE>
static int launch_tc3x(const tc3x::Params& p, size_t smem_bytes, int grid, int device, cudaStream_t stream) {
    auto kern = tc3x::logprob_tc3x_kernel<WIDE>;
    static std::mutex mu;
    static bool done[64] = {};
    {
        std::lock_guard<std::mutex> lock(mu);
        if (device < 0 || device >= 64 || !done[device]) {
            int max_smem = 0;
            FRL_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
            FRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
            if (device >= 0 && device < 64) done[device] = true;
        }
    }
    kern<<<grid, tc3x::kThreads, smem_bytes, stream>>>(p);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

// returns > 0 when the net does not fit the plan (the caller falls back to the FP32 FFMA kernel, same accuracy class)
int launch_logprob_tc3x(const ScoreArgs& a, cudaStream_t stream) {
    using namespace tc3x;
    const frl_net* n0 = a.net[0];
    if (!tc3x_supported(n0)) return 1;
    Schedule sc = make_schedule(n0);
    if ((int)sc.blk.size() > kMaxBlocks) return 1;
    const size_t img = ((size_t)sc.image_bytes + 255) / 256 * 256;

    Params p{};
    for (int r = 0; r < a.n_roles; ++r) {
        const uint8_t* base = static_cast<const uint8_t*>(a.net[r]->pack_tc3x);
        if (!base) {
            set_error("split-precision weight image missing for one of the nets");
            return FRL_ERR_UNSUPPORTED;
        }
        p.wpack[r] = base;   // the fp16 image
        p.bias[r] = reinterpret_cast<const float*>(base + 2 * img);
        p.sign[r] = a.sign[r];
        p.out[r] = a.out_logp[r];
    }
    p.rank_stride = (uint32_t)rank_image_stride(n0, sc.image_bytes);
    p.s = a.s; p.a = a.a; p.probe = a.probe;
    p.B = a.B; p.n_roles = a.n_roles; p.n_heads = n0->cfg.n_heads; p.depth = n0->cfg.depth;
    p.s_dim = n0->cfg.s_dim; p.a_dim = n0->cfg.a_dim; p.K0p = n0->d_in_p; p.NOp = n0->n_out_p;
    p.n_steps = a.n_steps;
    p.delta = (float)(1.0 / a.n_steps);
    for (int k = 0; k < a.n_steps; ++k) p.t_mid[k] = a.t_mid_host[k];
    p.probe_step_stride = a.probe_mode == FRL_PROBE_PER_STEP ? a.B * (long long)p.a_dim : 0;
    p.n_blocks = (int)sc.blk.size();
    for (int i = 0; i < p.n_blocks; ++i) p.blk[i] = sc.blk[i];

    uint32_t off = 1024;   // barriers + TMEM slot
    p.off_a = off; off += 2u * H * kRows * 2;
    p.x0_in_a = choose_layout(n0) == 1 ? 1 : 0;
    p.x0_bytes = p.x0_in_a ? 0u : (uint32_t)p.K0p * kRows * 2;
    p.off_x0 = off; off += 2u * p.x0_bytes;
    p.off_ones = off; off += 2u * kRows * 16;
    p.off_bias = off; off += (uint32_t)(a.n_roles * n0->n_out_p) * 4;
    off = (off + 15) / 16 * 16;
    p.off_state = off; off += (2u * p.a_dim * kSamp) * 4;
    off = (off + 15) / 16 * 16;
    p.off_rec = off; off += 5u * (uint32_t)p.n_blocks * 16;
    p.off_rrec = off; off += (uint32_t)p.n_blocks * 8;
    off = (off + 1023) / 1024 * 1024;
    p.off_ring = off;
    int max_smem = 0, sms = 0;
    FRL_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, n0->cfg.device));
    FRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, n0->cfg.device));
    int stages = ((int)max_smem - (int)off) / kSlotBytes;
    if (stages > kMaxSlots) stages = kMaxSlots;
    if (const char* cap = getenv("FRL_TC3_STAGES")) stages = std::min(stages, atoi(cap));   // experiments only
    if (stages < 2) return 1;
    p.n_stages = stages;
    const size_t smem_bytes = (size_t)off + (size_t)stages * kSlotBytes;

    const long long tiles_per_role = (a.B + kSamp - 1) / kSamp;
    const long long n_duos = ((tiles_per_role + 1) / 2) * a.n_roles;
    const int grid = 2 * (int)std::min<long long>(n_duos, sms / 2);   // clusters of two CTAs

    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    FRL_CUDA(cudaStreamIsCapturing(stream, &cap));
    if (p.x0_in_a) {
        const long long img_tiles = ((tiles_per_role + 1) / 2) * 2;   // the padding tile of the last duo included
        const size_t img_bytes = (size_t)p.K0p * kRows * 2, need = (size_t)img_tiles * 2 * img_bytes;
        frl_net* nm = const_cast<frl_net*>(n0);
        if (nm->ws_x0s_bytes < need) {
            if (cap != cudaStreamCaptureStatusNone) {
                set_error("the wide-input operand workspace must grow: run one launch of this size before capturing a graph");
                return FRL_ERR_UNSUPPORTED;
            }
            const size_t grow = std::max(need, nm->ws_x0s_bytes * 2);   // geometric: a growing batch retires O(log) buffers
            if (nm->ws_x0s) nm->retired.push_back(nm->ws_x0s);   // kept alive: see frl_net::retired
            nm->ws_x0s = nullptr;
            nm->ws_x0s_bytes = 0;
            FRL_CUDA(cudaMalloc(&nm->ws_x0s, grow));
            nm->ws_x0s_bytes = grow;
        }
        // the images are per launch: a launch on another stream waits for the previous reader
        if (!nm->x0s_event) FRL_CUDA(cudaEventCreateWithFlags(&nm->x0s_event, cudaEventDisableTiming));
        else if (cap == cudaStreamCaptureStatusNone) FRL_CUDA(cudaStreamWaitEvent(stream, nm->x0s_event, 0));
        const long long n_items = img_tiles * (p.K0p / 8) * kRows;
        const unsigned blocks = (unsigned)((n_items + 255) / 256);
        tc3x_state_image_kernel<<<blocks, 256, 0, stream>>>(a.s, static_cast<uint8_t*>(nm->ws_x0s), a.B, img_tiles, p.s_dim, p.a_dim, p.K0p);
        FRL_CUDA(cudaGetLastError());
        p.x0s = static_cast<const uint8_t*>(nm->ws_x0s);
    }
    auto q_launch = [&](const Params& q, size_t smem) {
        return q.x0_in_a ? launch_tc3x<true>(q, smem, grid, n0->cfg.device, stream)
                         : launch_tc3x<false>(q, smem, grid, n0->cfg.device, stream);
    };
    auto run = [&](const Params& q) {
        int rc = q_launch(q, smem_bytes);
        if (rc == FRL_OK && q.x0_in_a && cap == cudaStreamCaptureStatusNone) FRL_CUDA(cudaEventRecord(n0->x0s_event, stream));
        return rc;
    };
    if (const char* tp = getenv("FRL_TC_TRACE")) {
        // timing experiment: stamp the roles of cluster 0's leader CTA during one Euler step and dump them (synchronises!)
        if (smem_bytes + 4 * kTraceCap * 8 > (size_t)max_smem) {
            set_error("FRL_TC_TRACE: no shared memory left for the stamps with this net");
            return FRL_ERR_UNSUPPORTED;
        }
        unsigned long long* d = nullptr;
        FRL_CUDA(cudaMalloc(&d, 8 * kTraceCap * sizeof(unsigned long long)));
        FRL_CUDA(cudaMemset(d, 0, 8 * kTraceCap * sizeof(unsigned long long)));
        p.trace = d;
        p.trace_step = getenv("FRL_TC_TRACE_STEP") ? atoi(getenv("FRL_TC_TRACE_STEP")) : std::min(10, p.n_steps - 1);
        p.off_trace = (uint32_t)smem_bytes;
        p.basis = -1;
        int rc = q_launch(p, smem_bytes + 4 * kTraceCap * 8);
        if (rc != FRL_OK) return rc;
        FRL_CUDA(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(8 * kTraceCap);
        FRL_CUDA(cudaMemcpy(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(d);
        if (FILE* f = fopen(tp, "w")) {
            for (int r = 0; r < 8; ++r)   // 0-3: leader roles; 4: the peer CTA's epilogue warp 0 (its own SM clock)
                for (int i = 0; i < kTraceCap && h[(size_t)r * kTraceCap + i]; ++i)
                    fprintf(f, "%d %llx %llu\n", r, h[(size_t)r * kTraceCap + i] >> 32, h[(size_t)r * kTraceCap + i] & 0xFFFFFFFFull);
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
