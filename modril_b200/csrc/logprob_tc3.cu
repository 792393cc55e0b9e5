// tcgen05 fused log-density integrator for CTA PAIRS (sm_100a, cta_group::2) -- the default tensor-core scoring kernel.
// 64-sample tiles (primal + tangent rows = one M = 128 operand), N = 256, every MMA spanning two SMs.
//
// A round-1 single-CTA experiment showed that one CTA cannot hold two in-flight 64-sample tiles AND a weight ring deep
// enough for the lag between them (profiles/r01f_tc2_experiment.md).  Here two CTAs of a cluster form a pair:
//   * one tcgen05.mma.cta_group::2 (M = 256) multiplies the 128 operand rows of BOTH CTAs with the same weight block, so
//     a thread of the leader CTA issues for two SMs (half the issue cost per pipe cycle);
//   * each CTA loads only its N/2 = 128 rows of a weight block (B is split across the pair): 16 KB per [K = 64] block
//     (+ 4 KB of bias rows in the first block of a layer), four ring slots, and half the B-operand shared-memory
//     bandwidth per SM;
//   * each CTA keeps its own two tiles ("slots"), accumulators (2 x 256 TMEM columns), epilogue warps and producer;
//   * every layer's blocks are streamed once per slot ([layer l for slot 0][layer l for slot 1][layer l + 1 ...]), so a
//     ring slot is refilled as soon as the MMAs of the slot that read it commit (FRL_TC3_SHARED=1: one shared stream).
// Cross-CTA signalling: the peer's epilogue warps arrive on the leader's a_ready barriers through DSMEM (mapa + plain
// mbarrier.arrive.shared::cluster after fence.proxy.async), a relay thread in the peer forwards "my half of the weight
// block has landed", and tcgen05.commit.cta_group::2 ... multicast::cluster releases ring slots / publishes accumulators
// in both CTAs at once.  Reference
// arithmetic: modril/flowril/pretrain.py:192-214 / :112-134.  Measurements and the optimisation log:
// profiles/r01h_logprob_tc3_default.md.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "frl_internal.h"
#include "tc3_common.cuh"

namespace frl {

namespace tc3 {

using namespace tcc;
using namespace tc3h;

#ifndef FRL_TC3_EPI_WARPS
#define FRL_TC3_EPI_WARPS 8
#endif
constexpr int kEpiWarp0 = 6;
constexpr int kEpiWarps = FRL_TC3_EPI_WARPS;   // 8 or 16: four lane quarters x (2 or 4) column slices
constexpr int kColW = 256 / (kEpiWarps / 4);    // accumulator columns per epilogue warp (128 or 64)
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;   // warp 0 producer, warps 1-4 MMA issuers (2 per slot), warp 5 TMEM owner / peer relay, epilogue warps
constexpr int kSlotBytesA = 20480;  // plan A ring slot = this CTA's half of one weight block: [16 bias rows +] K = 64 rows of 128
constexpr int kSlotBytesH = 18432;  // halves plan: [16 bias rows +] K = 128 rows of 64
constexpr int kSlotBytesB = 16384;  // plan B (wider layer-0 inputs leave less room): blocks of at most 64 rows
constexpr int kMaxSlots = 16;
constexpr int kRows = 128;      // rows of the UMMA operand
constexpr int kSamp = 64;       // samples per tile
constexpr int H = 256;
constexpr int kMaxBlocks = tc::kMaxBlocks;

struct Blk {
    uint32_t goff, bytes;
    uint16_t nrows;     // N of the UMMA
    uint16_t a_koff;    // first 8-feature chunk of the A operand
    uint8_t n_k16;      // K = 16 MMAs
    uint8_t src;        // A operand: 0 activations, 1 layer-0 input
    uint8_t first;      // first MMA overwrites the accumulator
    uint8_t commit;     // last block of a layer / accumulator half: bit 0 / bit 1 = commit d_full[0] / d_full[1];
                        // bit 2 = commit ka_done (halves plan: the last block of the layer that reads operand columns < 128)
    uint8_t wait;       // bit0 / bit1: wait a_ready[slot][0 / 1] before issuing
    uint8_t bias;       // the block starts with 16 bias rows: first MMA = ones tile x bias rows, overwrites the accumulator
    uint8_t dhalf;      // accumulator half (TMEM column offset 128 * dhalf) the block accumulates into / commits
    uint8_t pad[1];
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
    int halves;   // trunk layers run as two N = 128 accumulator halves with their own d_full barriers (see make_schedule)
    float delta;
    uint32_t slot_bytes;         // ring slot size of the plan in use
    int debug;                   // FRL_TC_DEBUG (timing experiments): 8 skip epilogue stores, 16 skip the bias add
    uint32_t rank_stride;        // bytes between the weight images of CTA rank 0 and rank 1
    unsigned long long* trace;   // FRL_TC_TRACE: [4 roles][8192] clock stamps of CTA 0 (timing experiments only)
    uint32_t off_a, off_x0, off_ones, off_bias, off_state, off_rec, off_rrec, off_ring;
    uint32_t x0_bytes, state_floats;   // per slot
    int x0_in_a;   // wide layer-0 inputs (K0p > 32): no room for a persistent layer-0 operand tile.  The operand is rebuilt inside
                   // the slot's activation tile every step: one bulk copy of the tile's pre-built image (state columns in
                   // place, everything else zero; x0s, L2-resident) + the a / t / probe entries from the fp32 state
    const uint8_t* x0s;   // [tile][K0p/8][128][8] images of the state columns (x0_in_a only)
    Blk blk[kMaxBlocks];
    // Ring order of one Euler step.  Private streams (default): every layer's blocks are streamed twice, once per tile
    // slot -- [layer l for slot 0][layer l for slot 1][layer l+1 for slot 0] ... -- so the two slots run half a period
    // apart and the epilogue of one overlaps the MMAs of the other.  Shared stream (FRL_TC3_SHARED=1): every block is
    // loaded once and consumed by both slots, which keeps the slots in phase.
    int ring_len, w_consumers;
    uint8_t ring_blk[2 * kMaxBlocks];      // ring position -> block
    uint8_t ring_pos[2][kMaxBlocks];       // [slot][block] -> ring position
    float t_mid[kMaxSteps];
};

// WIDE: the wide-input mode (Params::x0_in_a) is a separate instantiation, so the common narrow-input kernel carries none
// of its code or registers.
template <typename T16, bool WIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) logprob_tc3_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t bar0 = smem_u32(smem);
    auto w_full = [&](int s) { return bar0 + 8u * s; };                    // own half of the block has landed (tx)
    auto w_empty = [&](int s) { return bar0 + 8u * (2 * kMaxSlots + s); }; // both issuers' MMAs on the slot completed
    auto a_ready = [&](int slot, int ch) { return bar0 + 8u * (3 * kMaxSlots + slot * 2 + ch); };   // leader only
    auto d_full = [&](int slot, int half) { return bar0 + 8u * (3 * kMaxSlots + (half ? 12 : 4) + slot); };
    auto hand = [&](int slot, int x) { return bar0 + 8u * (3 * kMaxSlots + 6 + slot * 2 + x); };   // leader only
    auto x0_full = [&](int slot) { return bar0 + 8u * (3 * kMaxSlots + 10 + slot); };   // wide inputs: operand image has landed
    // halves plan: the layer's MMAs no longer read operand columns 0..127, so the epilogue of half 0 may overwrite them (the
    // activation tile is updated in place while the MMAs of half 1 are still running)
    auto ka_done = [&](int slot) { return bar0 + 8u * (3 * kMaxSlots + 14 + slot); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);

    constexpr uint32_t a_tile_bytes = (uint32_t)H * kRows * 2;   // 64 KB per slot
    constexpr uint32_t kChunk = kRows * 16;                       // bytes of one 8-feature chunk of an operand image
    uint8_t* sA = smem + p.off_a;         // [2 slots][H/8][128][8]
    uint8_t* sX0 = smem + p.off_x0;       // [2 slots][K0p/8][128][8]
    uint8_t* sOnes = smem + p.off_ones;   // [2][128][8]: feature 0..2 = 1.0 on the primal rows (bias hi/mid/lo terms)
    float* sBias = reinterpret_cast<float*>(smem + p.off_bias);     // [roles][NOp] head biases (trunk biases ride in the MMA)
    float* sState = reinterpret_cast<float*>(smem + p.off_state);   // per slot: a [a_dim][64], probe [a_dim][64]
    uint4* sRec = reinterpret_cast<uint4*>(smem + p.off_rec);       // [2 slots][n_blocks][2]
    uint2* sRingRec = reinterpret_cast<uint2*>(smem + p.off_rrec);  // [ring_len] {offset, bytes} of the producer's loads
    uint8_t* sRing = smem + p.off_ring;   // n_stages slots of 16 KB: this CTA's N/2 rows of one weight block each

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();          // 0 = leader (issues the MMAs of the pair)
    const long long cluster_id = cluster_id_x();
    const long long n_clusters = gridDim.x / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            // leader: own producer (arrive.expect_tx) + the peer's "my half has landed" relay; peer: own producer only
            mbar_init(w_full(s), cta_rank == 0 ? 2 : 1);
            mbar_init(w_empty(s), p.w_consumers);   // released by the (multicast) commits of the slots that read it
        }
        for (int slot = 0; slot < 2; ++slot) {
            mbar_init(a_ready(slot, 0), kEpiWarps);   // 4 warps of this column half in each of the two CTAs
            mbar_init(a_ready(slot, 1), kEpiWarps);
            mbar_init(d_full(slot, 0), 1);
            mbar_init(d_full(slot, 1), 1);
            mbar_init(ka_done(slot), 1);
            mbar_init(x0_full(slot), 1);
            mbar_init(hand(slot, 0), 1);
            mbar_init(hand(slot, 1), 1);
        }
        fence_barrier_init();
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.NOp * p.n_roles; i += kThreads) sBias[i] = p.bias[i / p.NOp][p.depth * H + i % p.NOp];
    for (int i = threadIdx.x; i < 2 * kRows; i += kThreads) {
        const uint32_t one = Cvt<T16>::one(1.f);   // primal rows are rows 0..7 of every 16-row group
        reinterpret_cast<uint4*>(sOnes)[i] = (i < kRows && (i & 8) == 0) ? make_uint4(one | (one << 16), one, 0, 0)
                                                                           : make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < (int)(2 * p.x0_bytes / 16); i += kThreads)
        reinterpret_cast<uint4*>(sX0)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 2 * p.n_blocks; i += kThreads) {
        const int slot = i / p.n_blocks, b = i % p.n_blocks;
        const Blk& k = p.blk[b];
        const uint32_t src = (k.src == 0 || WIDE) ? smem_u32(sA) + slot * a_tile_bytes : smem_u32(sX0) + slot * p.x0_bytes;
        const uint32_t a_lo = (((src + (uint32_t)k.a_koff * kChunk) >> 4) & 0x3FFF) | ((kChunk >> 4) << 16);
        // M = 256 over the pair, N = nrows (each CTA holds nrows / 2 rows of B)
        const uint32_t idesc = (1u << 4) | (Cvt<T16>::kFmt << 7) | (Cvt<T16>::kFmt << 10) | ((uint32_t)(256 >> 4) << 24) |
                               ((uint32_t)(k.nrows >> 3) << 17);
        sRec[2 * i] = make_uint4(a_lo, idesc,
                                 (uint32_t)k.nrows | ((uint32_t)k.commit << 16) | ((uint32_t)k.wait << 24),
                                 (uint32_t)slot * 256 + (uint32_t)k.dhalf * 128);
        const uint32_t pos = p.ring_pos[slot][b];   // ring position inside a step, as (wraps, stage) of the ring
        sRec[2 * i + 1] = make_uint4((uint32_t)k.n_k16 | ((uint32_t)k.first << 8) | ((uint32_t)k.bias << 16),
                                     ((smem_u32(sOnes) >> 4) & 0x3FFF) | ((kChunk >> 4) << 16),
                                     (pos % (uint32_t)p.n_stages) | ((pos / (uint32_t)p.n_stages) << 8), 0);
    }
    for (int i = threadIdx.x; i < p.ring_len; i += kThreads) {
        const Blk& k = p.blk[p.ring_blk[i]];
        sRingRec[i] = make_uint2(k.goff, k.bytes);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // barrier inits of both CTAs are visible before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // work decomposition: a "quad" is four consecutive 64-sample tiles of one role: 2 per CTA of the pair
    const long long tiles_per_role = (p.B + kSamp - 1) / kSamp;
    const long long n_quads = ((tiles_per_role + 3) / 4) * p.n_roles;
    auto arrive_leader = [&](uint32_t bar) {
        if (cta_rank == 0) mbar_arrive(bar);
        else mbar_arrive_remote(bar, 0);
    };

    if (warp == 0) {
        // ------------------------------- producer: this CTA's half of every weight block --------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            Tracer tr; tr.init((cta_rank == 0 && p.debug == 0) ? p.trace : nullptr, 3, 0);   // role 3: the leader's weight producer
            for (long long quad = cluster_id; quad < n_quads; quad += n_clusters) {
                const uint8_t* wbase = p.wpack[quad % p.n_roles] + (size_t)cta_rank * p.rank_stride;
                for (int step = 0; step < p.n_steps; ++step) {
                    for (int i = 0; i < p.ring_len; ++i) {
                        const uint2 k = sRingRec[i];
                        tr.mark(0x100 + i);
                        mbar_wait(w_empty(stage), phase ^ 1);
                        tr.mark(0x200 + i);
                        mbar_expect_tx(w_full(stage), k.y);
                        bulk_g2s(smem_u32(sRing + (size_t)stage * p.slot_bytes), wbase + k.x, k.y, w_full(stage));
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 5 && cta_rank == 1) {
        // ------------------------------- peer: tell the leader when my half of a block has landed -----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long quad = cluster_id; quad < n_quads; quad += n_clusters)
                for (int step = 0; step < p.n_steps; ++step)
                    for (int i = 0; i < p.ring_len; ++i) {
                        mbar_wait(w_full(stage), phase);
                        mbar_arrive_remote(w_full(stage), 0);
                        if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp >= 1 && warp <= 4 && cta_rank == 0) {
        // ------------------------------- leader: the two MMA issuers of one slot (of both CTAs) -------------------
        // A thread pays ~60 cycles per tcgen05.mma and ~230 per tcgen05.commit.  Two threads per slot alternate blocks:
        // issuer x issues block b (b % 2 == x), hands the baton to its partner, and only then pays for its commits,
        // which run in the shadow of the partner's MMAs.  The baton keeps the MMAs of a slot in block order (the
        // accumulation order stays deterministic); the tensor pipe executes in issue order, so the commit of the
        // last block of a layer also covers the partner's earlier blocks.  With an even number of ring slots a slot
        // of the ring is always consumed by the same issuer, so no thread ever skips a phase of a barrier it waits on.
        if (lane == 0) {
            const int slot = (warp - 1) >> 1, x = (warp - 1) & 1;
            uint32_t par_a0 = 0, par_a1 = 0, par_h = 0;
            long long seq = 0;
            // ring position of the current step's first block, as (stage, wraps); a step advances it by ring_len
            uint32_t base_stage = 0, base_wraps = 0;
            const uint32_t len_stage = (uint32_t)p.ring_len % (uint32_t)p.n_stages, len_wraps = (uint32_t)p.ring_len / (uint32_t)p.n_stages;
            const uint4* rec = sRec + (size_t)slot * p.n_blocks * 2;
            const uint32_t ring_lo = smem_u32(sRing) >> 4;
            constexpr uint32_t kDescHi = (uint32_t)(128 >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
            Tracer tr;   // roles 1 / 2: issuer x = 0 of slots 0 / 1  (FRL_TC_DEBUG=100: both issuers of slot 0)
            if (p.debug == 100) tr.init(slot == 0 ? p.trace : nullptr, 1 + x);
            else tr.init(x == 0 ? p.trace : nullptr, 1 + slot);
            for (long long quad = cluster_id; quad < n_quads; quad += n_clusters) {
                for (int step = 0; step < p.n_steps; ++step) {
                    for (int b = 0; b < p.n_blocks; ++b, ++seq) {
                        const uint4 r0 = rec[2 * b], r1 = rec[2 * b + 1];
                        // every issuer observes every readiness phase of the operand (a skipped phase would alias parities)
                        if (r0.z >> 24) tr.mark(0x500 + b);
                        if (r0.z & (1u << 24)) { mbar_wait_cluster(a_ready(slot, 0), par_a0); par_a0 ^= 1; }
                        if (r0.z & (2u << 24)) { mbar_wait_cluster(a_ready(slot, 1), par_a1); par_a1 ^= 1; }
                        if (r0.z >> 24) tr.mark(0x600 + b);
                        if ((int)(seq & 1) == x) {
                            tr.mark(0x100 + b);
                            if (seq > 0) { mbar_wait(hand(slot, x), par_h); par_h ^= 1; }
                            uint32_t stage = base_stage + (r1.z & 0xFF), wraps = base_wraps + (r1.z >> 8);
                            if (stage >= (uint32_t)p.n_stages) { stage -= p.n_stages; ++wraps; }
                            const uint32_t phase = wraps & 1;
                            mbar_wait(w_full(stage), phase);   // both halves of the block are in the two CTAs' rings
                            tr.mark(0x300 + b);
                            tc_fence_after();
                            const uint32_t nrows = r0.z & 0xFFFF, nk = r1.x & 0xFF, nhalf = nrows >> 1;
                            uint32_t a_lo = r0.x;
                            uint32_t b_lo = ((ring_lo + (uint32_t)stage * (p.slot_bytes >> 4)) & 0x3FFF) | (nhalf << 16);
                            uint32_t acc = ((r1.x >> 8) & 1) ? 0u : 1u;
                            if ((r1.x >> 16) & 1) {   // bias rows first: ones tile x [bias_hi; bias_mid; bias_lo; 0 ...]
                                umma2_f16(tmem + r0.w, ((uint64_t)kDescHi << 32) | r1.y, ((uint64_t)kDescHi << 32) | b_lo, r0.y, 0u);
                                acc = 1u;
                                b_lo += 2 * nhalf;
                            }
                            for (uint32_t u = 0; u < nk; ++u) {
                                umma2_f16(tmem + r0.w, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, r0.y, acc);
                                acc = 1u;
                                a_lo += 2 * (kChunk >> 4);
                                b_lo += 2 * nhalf;
                            }
                            mbar_arrive(hand(slot, x ^ 1));
                            umma2_commit_both(w_empty(stage));
                            const uint32_t cm = (r0.z >> 16) & 0xFF;
                            if (cm & 4) umma2_commit_both(ka_done(slot));
                            if (cm & 1) umma2_commit_both(d_full(slot, 0));
                            if (cm & 2) umma2_commit_both(d_full(slot, 1));
                            tr.mark(0x400 + b);
                        }
                    }
                    base_stage += len_stage;
                    base_wraps += len_wraps;
                    if (base_stage >= (uint32_t)p.n_stages) { base_stage -= p.n_stages; ++base_wraps; }
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ------------------------------------------------ epilogue ------------------------------------------------
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;                 // TMEM lane quarter: operand rows 32q .. 32q+31 = samples 16q .. 16q+15
        const int ch = ew >> 2;                 // column slice (kColW columns)
        const int ah = (ch * kColW) >> 7;       // the 128-column half (a_ready barrier) the slice belongs to
        const int et = ew * 32 + lane;
        const int tq = lane & 3, tr8 = lane >> 2;   // position inside the 16x256b fragment: column pair, row
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const int t_col = p.a_dim + p.s_dim;
        uint32_t par_d[2] = {0, 0}, par_k[2] = {0, 0};
        // halves mode: this warp drains accumulator half `ah` and follows d_full[ah] only (the heads commit both barriers, so
        // no warp ever skips a phase of the barrier it waits on)
        const int dh = p.halves ? ah : 0;
        float log_det[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [slot][16-row group]
        Tracer tr;   // role 0: leader CTA's first epilogue warp, role 3: the peer CTA's
        tr.init((lane == 0 && (ew == 0 || ew == p.debug) && cta_rank == 0) ? p.trace : nullptr, ew == 0 ? 0 : 3, 0);

        // wide inputs: start the bulk copy of a tile's state-column image into the slot's activation tile (one thread), and
        // -- once it has landed -- write the a / t (primal rows) and probe (tangent rows) entries of a sample from the fp32 state
        uint32_t par_x0[2] = {0, 0};
        const uint32_t x0_img_bytes = (uint32_t)p.K0p * kRows * 2;
        auto x0_fetch = [&](int slot, long long tile) {
            mbar_expect_tx(x0_full(slot), x0_img_bytes);
            const uint8_t* src = p.x0s + (size_t)tile * x0_img_bytes;
            for (uint32_t off = 0; off < x0_img_bytes; off += 16384u)
                bulk_g2s(smem_u32(sA) + slot * a_tile_bytes + off, src + off, min(16384u, x0_img_bytes - off), x0_full(slot));
        };
        auto x0_sample = [&](uint8_t* x0, const float* sA_state, const float* sP_state, int srow, int j0, int jstep, bool write_t, float tval) {
            const int rowp = row_primal(srow);
            for (int j = j0; j < p.a_dim; j += jstep) {
                *reinterpret_cast<uint16_t*>(x0 + (j >> 3) * kChunk + rowp * 16 + (j & 7) * 2) = Cvt<T16>::one(sA_state[j * kSamp + srow]);
                *reinterpret_cast<uint16_t*>(x0 + (j >> 3) * kChunk + (rowp + 8) * 16 + (j & 7) * 2) = Cvt<T16>::one(sP_state[j * kSamp + srow]);
            }
            if (write_t) *reinterpret_cast<uint16_t*>(x0 + (t_col >> 3) * kChunk + rowp * 16 + (t_col & 7) * 2) = Cvt<T16>::one(tval);
        };

        for (long long quad = cluster_id; quad < n_quads; quad += n_clusters) {
            const int role = (int)(quad % p.n_roles);
            const long long tile0 = 4 * (quad / p.n_roles) + 2 * cta_rank;   // this CTA's two tiles (padding tiles run masked)
            constexpr int n_slots = 2;
            const float sign = p.sign[role];
            const float* bh = sBias + role * p.NOp;

            // ---- tile init: layer-0 operand rows + fp32 state (X0 was zeroed once; padding / tangent s,t columns stay 0) ----
            for (int slot = 0; slot < n_slots; ++slot) {
                const long long row0 = (tile0 + slot) * kSamp;
                uint8_t* x0 = sX0 + slot * p.x0_bytes;
                float* sA_state = sState + slot * p.state_floats;
                float* sP_state = sA_state + p.a_dim * kSamp;
                if (WIDE) {   // fp32 state first, then (after the barrier below) the operand from it
                    for (int idx = et; idx < kSamp * p.a_dim; idx += kEpiWarps * 32) {
                        const int r = idx / p.a_dim, c = idx % p.a_dim;
                        const long long g = row0 + r;
                        float pr = 0.f;
                        if (g < p.B) pr = p.basis >= 0 ? (c == p.basis ? 1.f : 0.f) : p.probe[g * p.a_dim + c];
                        sA_state[c * kSamp + r] = g < p.B ? p.a[g * p.a_dim + c] : 0.f;
                        sP_state[c * kSamp + r] = pr;
                    }
                    log_det[slot][0] = log_det[slot][1] = 0.f;
                    continue;
                }
                for (int idx = et; idx < kSamp * p.s_dim; idx += kEpiWarps * 32) {
                    const int r = idx / p.s_dim, c = idx % p.s_dim;
                    const long long g = row0 + r;
                    const float v = g < p.B ? p.s[g * p.s_dim + c] : 0.f;
                    const int kk = p.a_dim + c;
                    *reinterpret_cast<uint16_t*>(x0 + (kk >> 3) * kChunk + row_primal(r) * 16 + (kk & 7) * 2) = Cvt<T16>::one(v);
                }
                for (int idx = et; idx < kSamp * p.a_dim; idx += kEpiWarps * 32) {
                    const int r = idx / p.a_dim, c = idx % p.a_dim;
                    const long long g = row0 + r;
                    const float v = g < p.B ? p.a[g * p.a_dim + c] : 0.f;
                    float pr = 0.f;
                    if (g < p.B) pr = p.basis >= 0 ? (c == p.basis ? 1.f : 0.f) : p.probe[g * p.a_dim + c];
                    sA_state[c * kSamp + r] = v;
                    sP_state[c * kSamp + r] = pr;
                    *reinterpret_cast<uint16_t*>(x0 + (c >> 3) * kChunk + row_primal(r) * 16 + (c & 7) * 2) = Cvt<T16>::one(v);
                    *reinterpret_cast<uint16_t*>(x0 + (c >> 3) * kChunk + (row_primal(r) + 8) * 16 + (c & 7) * 2) = Cvt<T16>::one(pr);
                }
                if (et < kSamp)
                    *reinterpret_cast<uint16_t*>(x0 + (t_col >> 3) * kChunk + row_primal(et) * 16 + (t_col & 7) * 2) =
                        Cvt<T16>::one(p.t_mid[0]);
                log_det[slot][0] = log_det[slot][1] = 0.f;
            }
            if (WIDE) {
                named_sync(1, kEpiWarps * 32);   // the fp32 state of both tiles is complete
                if (et == 0)
                    for (int slot = 0; slot < n_slots; ++slot) x0_fetch(slot, tile0 + slot);
                for (int slot = 0; slot < n_slots; ++slot) {
                    mbar_wait(x0_full(slot), par_x0[slot]);
                    par_x0[slot] ^= 1;
                    const float* sA_state = sState + slot * p.state_floats;
                    if (et < kSamp) x0_sample(sA + slot * a_tile_bytes, sA_state, sA_state + p.a_dim * kSamp, et, 0, 1, true, p.t_mid[0]);
                }
            }
            fence_async_smem();
            named_sync(1, kEpiWarps * 32);
            if (lane == 0)
                for (int slot = 0; slot < n_slots; ++slot) arrive_leader(a_ready(slot, ah));

            for (int step = 0; step < p.n_steps; ++step) {
                // ---- trunk layers: accumulator -> next layer's operand ----
                for (int l = 0; l < p.depth; ++l) {
                    for (int slot = 0; slot < n_slots; ++slot) {
                        tr.mark(0x100 + l * 2 + slot);
                        mbar_wait(d_full(slot, dh), par_d[slot]);
                        par_d[slot] ^= 1;
                        tc_fence_after();
                        tr.mark(0x200 + l * 2 + slot);
                        // Both 16-row groups of this warp's lane quarter, 128 columns each, in two software-pipelined
                        // rounds: the TMEM loads of group 1 are in flight while group 0 is converted and stored (a
                        // tcgen05.wait::ld covers every outstanding load, so the rounds are the pipelining granularity).
                        constexpr int kHalves = kColW / 64;
                        uint32_t z[2][kHalves][32];
                        auto load_grp = [&](int grp) {
                            const uint32_t tb = lane_addr + ((uint32_t)(grp * 16) << 16) + (uint32_t)slot * 256 + (uint32_t)ch * kColW;
#pragma unroll
                            for (int half = 0; half < kHalves; ++half) tmem_ld_16x256b_x8(tb + 64 * half, z[grp][half]);
                        };
                        auto store_grp = [&](int grp) {
                            // lane i addresses row i%8 of matrix i/8: matrices 0/1 = primal/tangent rows of chunk k,
                            // matrices 2/3 = the same of chunk k+1
                            const uint32_t rowa = (uint32_t)(q * 32 + grp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8);
                            uint32_t addr = smem_u32(sA) + slot * a_tile_bytes + (uint32_t)(ch * (kColW / 8) + (lane >> 4)) * kChunk + rowa * 16;
#pragma unroll
                            for (int half = 0; half < kHalves; ++half) {
#pragma unroll
                                for (int k = 0; k < 8; k += 2) {
                                    uint32_t w[4];
#pragma unroll
                                    for (int e = 0; e < 2; ++e) {
                                        // z[4k..4k+3] = primal (c, c+1), tangent (c, c+1) with c = half*64 + 8k + 2 tq; the bias is
                                        // already in the accumulator.  primal: relu + round + pack in one F2FP; tangent:
                                        // masked by the sign bits of the primal pre-activation
                                        const uint32_t lo = z[grp][half][4 * (k + e)], hi = z[grp][half][4 * (k + e) + 1];
                                        w[2 * e] = Cvt<T16>::pack2_relu(lo, hi);
                                        w[2 * e + 1] = Cvt<T16>::pack2_sat(z[grp][half][4 * (k + e) + 2], z[grp][half][4 * (k + e) + 3]) &
                                                       ~neg_mask16x2(lo, hi);
                                    }
                                    stmatrix_x4(addr, w[0], w[1], w[2], w[3]);
                                    addr += 2 * kChunk;
                                }
                            }
                        };
                        load_grp(0);
#pragma unroll
                        for (int half = 0; half < kHalves; ++half) tmem_ld_wait_dep(z[0][half]);
                        tr.mark(0x500 + l * 2 + slot);
                        load_grp(1);
                        if (p.halves && ah == 0) {   // in-place update: wait until half 1's MMAs are past operand columns 0..127
                            mbar_wait(ka_done(slot), par_k[slot]);
                            par_k[slot] ^= 1;
                        }
                        store_grp(0);
                        tr.mark(0x600 + l * 2 + slot);
#pragma unroll
                        for (int half = 0; half < kHalves; ++half) tmem_ld_wait_dep(z[1][half]);
                        tr.mark(0x700 + l * 2 + slot);
                        store_grp(1);
                        tr.mark(0x800 + l * 2 + slot);
                        tc_fence_before();
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) arrive_leader(a_ready(slot, ah));
                        tr.mark(0x300 + l * 2 + slot);
                    }
                }
                // ---- heads: tanh, divergence, Euler update (pretrain.py:205-211) ----
                for (int slot = 0; slot < n_slots; ++slot) {
                    tr.mark(0x1F0 + slot);
                    mbar_wait(d_full(slot, dh), par_d[slot]);
                    par_d[slot] ^= 1;
                    tc_fence_after();
                    tr.mark(0x2F0 + slot);
                    // wide inputs: the activation tile is free (its last reader, the heads MMA, is done): fetch the next step's
                    // layer-0 operand image while the heads are evaluated
                    if (WIDE && step + 1 < p.n_steps && et == 0) x0_fetch(slot, tile0 + slot);
                    {
                        // the heads have only NOp <= 64 columns: the two warps of a lane quarter split its two 16-row groups
                        uint8_t* x0 = WIDE ? sA + slot * a_tile_bytes : sX0 + slot * p.x0_bytes;
                        float* sA_state = sState + slot * p.state_floats;
                        float* sP_state = sA_state + p.a_dim * kSamp;
                        if (ch % (kEpiWarps / 8) == 0) {   // one warp per 16-row group works on the heads
                            const int grp = ch / (kEpiWarps / 8);
                            const int srow = q * 16 + grp * 8 + tr8;             // sample inside the tile
                            const int rowp = q * 32 + grp * 16 + tr8;
                            const long long g = (tile0 + slot) * kSamp + srow;
                            const uint32_t hb = lane_addr + ((uint32_t)(grp * 16) << 16) + (uint32_t)slot * 256;
                            float div = 0.f;
                            for (int g16 = 0; g16 < p.NOp; g16 += 16) {
                                uint32_t vv[8];
                                tmem_ld_16x256b_x2(hb + (uint32_t)g16, vv);
                                tmem_ld_wait_dep8(vv);
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    const int c = g16 + 8 * k + 2 * tq;   // this thread's column pair (c, c+1): primal vv[4k], vv[4k+1]; tangent vv[4k+2], vv[4k+3]
                                    if (p.n_heads == 2) {
                                        const int j = c >> 1;             // columns (2j, 2j+1) = (v_c, rho) of action dim j
                                        if (j < p.a_dim) {
                                            const float th = tanhf(__uint_as_float(vv[4 * k + 1]) + bh[c + 1]);
                                            const float v = __uint_as_float(vv[4 * k]) + bh[c] + sign * th;
                                            const float dv = __uint_as_float(vv[4 * k + 2]) +
                                                             sign * (1.f - th * th) * __uint_as_float(vv[4 * k + 3]);
                                            div = fmaf(dv, sP_state[j * kSamp + srow], div);
                                            const float an = sA_state[j * kSamp + srow] + v * p.delta;
                                            sA_state[j * kSamp + srow] = an;
                                            if (!WIDE)
                                                *reinterpret_cast<uint16_t*>(x0 + (j >> 3) * kChunk + rowp * 16 + (j & 7) * 2) = Cvt<T16>::one(an);
                                        }
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 2; ++e) {
                                            const int j = c + e;
                                            if (j < p.a_dim) {
                                                const float v = __uint_as_float(vv[4 * k + e]) + bh[j];
                                                div = fmaf(__uint_as_float(vv[4 * k + 2 + e]), sP_state[j * kSamp + srow], div);
                                                const float an = sA_state[j * kSamp + srow] + v * p.delta;
                                                sA_state[j * kSamp + srow] = an;
                                                if (!WIDE)
                                                    *reinterpret_cast<uint16_t*>(x0 + (j >> 3) * kChunk + rowp * 16 + (j & 7) * 2) = Cvt<T16>::one(an);
                                            }
                                        }
                                    }
                                }
                            }
                            // the four threads of a sample hold different action dims: sum their divergence terms
                            div += __shfl_xor_sync(0xffffffffu, div, 1);
                            div += __shfl_xor_sync(0xffffffffu, div, 2);
                            log_det[slot][grp] -= p.delta * div;
                            __syncwarp();   // the state of this sample (written by its four threads) is complete
                            if (step + 1 < p.n_steps) {
                                if (tq == 0 && !WIDE)
                                    *reinterpret_cast<uint16_t*>(x0 + (t_col >> 3) * kChunk + rowp * 16 + (t_col & 7) * 2) =
                                        Cvt<T16>::one(p.t_mid[step + 1]);
                                if (p.probe_step_stride != 0) {
                                    for (int j = tq; j < p.a_dim; j += 4) {
                                        const float pr = g < p.B ? p.probe[(step + 1) * p.probe_step_stride + g * p.a_dim + j] : 0.f;
                                        sP_state[j * kSamp + srow] = pr;
                                        if (!WIDE)
                                            *reinterpret_cast<uint16_t*>(x0 + (j >> 3) * kChunk + (rowp + 8) * 16 + (j & 7) * 2) =
                                                Cvt<T16>::one(pr);
                                    }
                                }
                            } else if (tq == 0 && g < p.B) {
                                if (p.accumulate) {
                                    p.out[role][g] += log_det[slot][grp];
                                } else {
                                    float prior = 0.f;
                                    for (int j = 0; j < p.a_dim; ++j) {
                                        const float x = sA_state[j * kSamp + srow];
                                        prior += -0.5f * x * x - 0.91893853320467274178f;
                                    }
                                    p.out[role][g] = prior + log_det[slot][grp];
                                }
                            }
                        }
                    }
                    if (step + 1 < p.n_steps) {
                        if (WIDE) {   // the image has landed: this thread's entries of its sample (a, probe; t by the first of the four)
                            mbar_wait(x0_full(slot), par_x0[slot]);
                            par_x0[slot] ^= 1;
                            if (ch % (kEpiWarps / 8) == 0) {
                                const float* sA_state = sState + slot * p.state_floats;
                                x0_sample(sA + slot * a_tile_bytes, sA_state, sA_state + p.a_dim * kSamp,
                                          q * 16 + (ch / (kEpiWarps / 8)) * 8 + tr8, tq, 4, tq == 0, p.t_mid[step + 1]);
                            }
                        }
                        tc_fence_before();
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) arrive_leader(a_ready(slot, ah));
                    }
                    tr.mark(0x3F0 + slot);
                }
            }
            // the next pair's init rewrites X0 / state: every epilogue warp must be done with this pair first
            tc_fence_before();
            named_sync(1, kEpiWarps * 32);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the leader's last MMAs read the peer's shared memory: nobody leaves or frees TMEM early
    if (warp == 5) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// Wide inputs: operand images of the state columns, one per 64-sample tile, in the layout of the layer-0 operand
// ([K0p/8][128 rows][8]; primal rows carry s in columns [a_dim, a_dim + s_dim), everything else is zero).  Built once per
// launch; the integrator copies a tile's image into shared memory at every Euler step (L2 hits after the first).
template <typename T16>
__global__ void tc3_state_image_kernel(const float* __restrict__ s, uint8_t* __restrict__ img, long long B, long long n_tiles,
                                       int s_dim, int a_dim, int K0p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int chunks = K0p / 8;
    if (i >= n_tiles * chunks * kRows) return;
    const int rowi = (int)(i % kRows);
    const int chunk = (int)((i / kRows) % chunks);
    const long long tile = i / ((long long)kRows * chunks);
    uint32_t w[4] = {0, 0, 0, 0};
    if ((rowi & 8) == 0) {
        const long long g = tile * kSamp + (((rowi >> 4) << 3) | (rowi & 7));
        if (g < B) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = chunk * 8 + e - a_dim;
                const uint32_t h = (c >= 0 && c < s_dim) ? Cvt<T16>::one(s[g * s_dim + c]) : 0u;
                w[e >> 1] |= h << (16 * (e & 1));
            }
        }
    }
    *reinterpret_cast<uint4*>(img + (size_t)i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Block stream of one Euler step (shared by the packer and the kernel).  One kernel block = one ring slot of each CTA =
// one tcgen05.commit: [K = 64][N = 256] for the trunk (each CTA of the pair holds 128 of the 256 rows: 16 KB), the whole
// (<= 32 column) layer-0 input, or half of the heads' K.  Blk::goff / Blk::bytes describe ONE CTA's half.
struct Schedule {
    std::vector<Blk> blk;
    std::vector<tc::PackBlk> pack[2];   // per CTA rank
    uint32_t image_bytes = 0;           // of one rank's image
};

static bool supported(const frl_net* net) {
    return net->cfg.hidden == H && net->d_in_p <= 192 && net->n_out_p <= 64;
}
// Shared-memory layout of a net: the weight-stream plan and whether the layer-0 operand shares the activation tile
// (Params::x0_in_a: needed when it is wider than 32 columns, and chosen when dropping the separate tiles is what makes the
// halves plan fit).  Plans:
//   halves: four 18 KB slots, trunk layer = 4 blocks of N = 128  [h0: bias16 + K128 | K128] [h1: bias16 + K128 | K128]
//   A:      four 20 KB slots, trunk layer = 4 blocks of N = 256  [bias16 + K64] [K64] [K64] [K64]
//   B:      four 16 KB slots, trunk layer = 5 blocks of N = 256  [bias16 + K48] [K64] [K64] [K64] [K16]
// The choice depends only on the net's dimensions, so the packer and every launch agree.
struct Layout {
    int plan;    // 2 halves, 0 A, 1 B, -1 none (the caller falls back to the single-CTA kernel)
    bool wide;
};
static int block_count(const frl_net* net, int plan, bool wide) {
    const int depth = net->cfg.depth, K0p = net->d_in_p;
    if (plan == 2) return 2 * ((K0p + 127) / 128) + 4 * (depth - 1) + 2;
    const int first_k = plan == 0 ? 64 : 48;
    const int l0 = wide ? 1 + (std::max(0, K0p - first_k) + 63) / 64 : 1;
    return l0 + (plan == 0 ? 4 : 5) * (depth - 1) + 2;
}
static uint32_t front_bytes(const frl_net* net, int n_roles, int plan, bool wide) {
    uint32_t off = 1024;
    off += 2u * H * kRows * 2;
    if (!wide) off += 2u * (uint32_t)net->d_in_p * kRows * 2;
    off += 2u * kRows * 16;
    off += (uint32_t)(n_roles * net->n_out_p) * 4;
    off = (off + 15) / 16 * 16;
    off += 2u * (2u * net->cfg.a_dim * kSamp) * 4;
    off = (off + 15) / 16 * 16;
    off += (uint32_t)block_count(net, plan, wide) * (4u * 16 + 8);   // issue records (2 slots x 2 uint4) + producer records
    return (off + 1023) / 1024 * 1024;
}
static Layout choose_layout(const frl_net* net) {
    const uint32_t max_smem = 232448;   // sm_100a opt-in maximum of dynamic shared memory per CTA
    const bool must_wide = net->d_in_p > 32;
    if (const char* force = getenv("FRL_TC3_PLAN")) return Layout{atoi(force), must_wide};   // experiments only
    if (!must_wide && front_bytes(net, 2, 2, false) + 4u * kSlotBytesH <= max_smem) return Layout{2, false};
    if (front_bytes(net, 2, 2, true) + 4u * kSlotBytesH <= max_smem) return Layout{2, true};
    if (front_bytes(net, 2, 0, must_wide) + 4u * kSlotBytesA <= max_smem) return Layout{0, must_wide};
    if (front_bytes(net, 2, 1, must_wide) + 4u * kSlotBytesB <= max_smem) return Layout{1, must_wide};
    return Layout{-1, must_wide};
}

static Schedule make_schedule(const frl_net* net) {
    Schedule sc;
    const int depth = net->cfg.depth, K0p = net->d_in_p, NOp = net->n_out_p;
    const Layout lay = choose_layout(net);
    const int plan = lay.plan;
    uint32_t goff = 0;
    // n_base: first output feature of the block (both CTAs together hold features n_base .. n_base + nrows - 1, rank r the
    // r-th half of them); dhalf: accumulator half
    auto add = [&](int layer, int nrows, int k0, int klen, int a_koff, int src, int bias, int first, int commit, int wait,
                   int n_base = 0, int dhalf = 0) {
        Blk b{};
        b.goff = goff; b.nrows = (uint16_t)nrows; b.a_koff = (uint16_t)a_koff;
        b.n_k16 = (uint8_t)(klen / 16); b.src = (uint8_t)src; b.first = (uint8_t)first; b.commit = (uint8_t)commit;
        b.wait = (uint8_t)wait; b.bias = (uint8_t)bias; b.dhalf = (uint8_t)dhalf;
        if (bias) {
            for (int r = 0; r < 2; ++r) sc.pack[r].push_back(tc::PackBlk{goff, nrows / 2, 16, n_base + r * (nrows / 2), -1, layer});
            goff += (uint32_t)(nrows / 2) * 16 * 2;
        }
        for (int r = 0; r < 2; ++r) sc.pack[r].push_back(tc::PackBlk{goff, nrows / 2, klen, n_base + r * (nrows / 2), k0, layer});
        goff += (uint32_t)(nrows / 2) * klen * 2;
        b.bytes = goff - b.goff;
        sc.blk.push_back(b);
    };
    // the accumulator-overwriting first block of every layer waits a_ready[slot][0] and [1]: both column halves of the
    // previous epilogue (in both CTAs) have drained the accumulator and written the whole operand
    if (plan == 2) {
        // Halves: every trunk layer runs as two N = 128 accumulator halves (output features 0..127 / 128..255), each with
        // its own d_full barrier, in K = 128 blocks: [h0: K 0..127 | K 128..255] [h1: K 0..127 | K 128..255].  The epilogue
        // warps of half 0 drain it while the MMAs of half 1 run, and the next layer's first block (h0, K 0..127: it needs
        // only the activations and the accumulator columns of half 0) starts while half 1 is still being drained.
        for (int h = 0; h < 2; ++h)   // layer 0: the whole (narrow) input, or K = 128 pieces of a wide one
            for (int k0 = 0; k0 < K0p; k0 += 128) {
                const int klen = std::min(128, K0p - k0);
                add(0, 128, k0, klen, k0 / 8, 1, k0 == 0, k0 == 0, (k0 + klen >= K0p ? 1 << h : 0) | ((h == 1 && k0 == 0) ? 4 : 0),
                    (h == 0 && k0 == 0) ? 3 : 0, 128 * h, h);
            }
        for (int l = 1; l < depth; ++l) {
            for (int h = 0; h < 2; ++h)
                for (int k0 = 0; k0 < H; k0 += 128)
                    add(l, 128, k0, 128, k0 / 8, 0, k0 == 0, k0 == 0, (k0 == 128 ? 1 << h : 0) | ((h == 1 && k0 == 0) ? 4 : 0),
                        h == 0 ? (k0 == 0 ? 1 : 2) : 0, 128 * h, h);
        }
        add(depth, NOp, 0, H / 2, 0, 0, 0, 1, 0, 1);
        add(depth, NOp, H / 2, H / 2, (H / 2) / 8, 0, 0, 0, 3, 2);   // both barriers: see the epilogue's d_full note
        sc.image_bytes = goff;
        return sc;
    }
    if (!lay.wide) {
        add(0, H, 0, K0p, 0, 1, 1, 1, 1, 3);
    } else {   // layer 0 in ring-slot sized pieces: [bias16 + K64 | K48] [K64] ...
        const int first_k = plan == 0 ? 64 : 48;
        for (int k0 = 0; k0 < K0p;) {
            const int klen = std::min(k0 == 0 ? first_k : 64, K0p - k0);
            add(0, H, k0, klen, k0 / 8, 1, k0 == 0, k0 == 0, k0 + klen >= K0p, k0 == 0 ? 3 : 0);
            k0 += klen;
        }
    }
    for (int l = 1; l < depth; ++l) {
        if (plan == 0) {
            for (int k0 = 0; k0 < H; k0 += 64) add(l, H, k0, 64, k0 / 8, 0, k0 == 0, k0 == 0, k0 + 64 >= H, k0 == 0 ? 3 : 0);
        } else {
            add(l, H, 0, 48, 0, 0, 1, 1, 0, 3);
            for (int k0 = 48; k0 < 240; k0 += 64) add(l, H, k0, 64, k0 / 8, 0, 0, 0, 0, 0);
            add(l, H, 240, 16, 240 / 8, 0, 0, 0, 1, 0);
        }
    }
    add(depth, NOp, 0, H / 2, 0, 0, 0, 1, 0, 3);
    add(depth, NOp, H / 2, H / 2, (H / 2) / 8, 0, 0, 0, 1, 0);
    sc.image_bytes = goff;
    return sc;
}

static size_t rank_image_stride(const frl_net* net, uint32_t image_bytes) {
    const size_t img = ((size_t)image_bytes + 255) / 256 * 256;
    const size_t nb = (size_t)net->cfg.depth * net->cfg.hidden + net->n_out_p;
    return (2 * img + nb * sizeof(float) + 255) / 256 * 256;
}

}  // namespace tc3

int tc3_pack(frl_net* net, const float* params_dev, cudaStream_t stream) {
    using namespace tc3;
    if (!supported(net) || choose_layout(net).plan < 0) return FRL_OK;
    Schedule sc = make_schedule(net);
    if ((int)sc.blk.size() > kMaxBlocks) return FRL_OK;
    // [rank 0: fp16 image | bf16 image | bias] [rank 1: ...]
    const size_t stride = rank_image_stride(net, sc.image_bytes);
    if (!net->pack_tc3) FRL_CUDA(cudaMalloc(&net->pack_tc3, 2 * stride));
    for (int r = 0; r < 2; ++r) {
        void* dst = static_cast<uint8_t*>(net->pack_tc3) + r * stride;
        int rc = tc_pack_images(net, sc.pack[r], sc.image_bytes, &dst, params_dev, stream);
        if (rc != FRL_OK) return rc;
    }
    return FRL_OK;
}

template <typename T16, bool WIDE>
static int launch_tc3(const tc3::Params& p, size_t smem_bytes, int grid, int device, cudaStream_t stream) {
    auto kern = tc3::logprob_tc3_kernel<T16, WIDE>;
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
    kern<<<grid, tc3::kThreads, smem_bytes, stream>>>(p);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

int launch_logprob_tc3(const ScoreArgs& a, int precision, cudaStream_t stream) {
    using namespace tc3;
    const frl_net* n0 = a.net[0];
    Schedule sc = make_schedule(n0);
    const size_t img = ((size_t)sc.image_bytes + 255) / 256 * 256;
    const bool f16 = precision == FRL_PREC_FP16;

    Params p{};
    for (int r = 0; r < a.n_roles; ++r) {
        const uint8_t* base = static_cast<const uint8_t*>(a.net[r]->pack_tc3);
        if (!base) {
            set_error("tensor-core weight image missing for one of the nets");
            return FRL_ERR_UNSUPPORTED;
        }
        p.wpack[r] = f16 ? base : base + img;
        p.bias[r] = reinterpret_cast<const float*>(base + 2 * img);
        p.sign[r] = a.sign[r];
        p.out[r] = a.out_logp[r];
    }
    p.rank_stride = (uint32_t)rank_image_stride(n0, sc.image_bytes);
    if (const char* dbg = getenv("FRL_TC_DEBUG")) p.debug = atoi(dbg);
    p.s = a.s; p.a = a.a; p.probe = a.probe;
    p.B = a.B; p.n_roles = a.n_roles; p.n_heads = n0->cfg.n_heads; p.depth = n0->cfg.depth;
    p.s_dim = n0->cfg.s_dim; p.a_dim = n0->cfg.a_dim; p.K0p = n0->d_in_p; p.NOp = n0->n_out_p;
    p.n_steps = a.n_steps;
    p.delta = (float)(1.0 / a.n_steps);
    for (int k = 0; k < a.n_steps; ++k) p.t_mid[k] = a.t_mid_host[k];
    p.probe_step_stride = a.probe_mode == FRL_PROBE_PER_STEP ? a.B * (long long)p.a_dim : 0;
    p.n_blocks = (int)sc.blk.size();
    for (int i = 0; i < p.n_blocks; ++i) p.blk[i] = sc.blk[i];
    const char* shared_env = getenv("FRL_TC3_SHARED");   // experiments / A-B tests only
    const bool shared_stream = shared_env && shared_env[0] == '1';
    if (shared_stream) {
        p.ring_len = p.n_blocks;
        p.w_consumers = 2;
        for (int i = 0; i < p.n_blocks; ++i) p.ring_blk[i] = p.ring_pos[0][i] = p.ring_pos[1][i] = (uint8_t)i;
    } else {
        p.ring_len = 0;
        p.w_consumers = 1;
        for (int b0 = 0; b0 < p.n_blocks;) {
            int b1 = b0;
            while (!(p.blk[b1].commit & 3)) ++b1;   // a phase ends with a block that commits an accumulator (half)
            for (int slot = 0; slot < 2; ++slot)
                for (int b = b0; b <= b1; ++b) {
                    p.ring_pos[slot][b] = (uint8_t)p.ring_len;
                    p.ring_blk[p.ring_len++] = (uint8_t)b;
                }
            b0 = b1 + 1;
        }
    }

    uint32_t off = 1024;   // barriers + TMEM slot
    p.off_a = off; off += 2u * H * kRows * 2;
    const Layout lay = choose_layout(n0);
    p.x0_in_a = lay.wide ? 1 : 0;
    p.x0_bytes = p.x0_in_a ? 0u : (uint32_t)p.K0p * kRows * 2;
    p.off_x0 = off; off += 2u * p.x0_bytes;
    p.off_ones = off; off += 2u * kRows * 16;
    p.off_bias = off; off += (uint32_t)(a.n_roles * n0->n_out_p) * 4;
    off = (off + 15) / 16 * 16;
    p.state_floats = 2u * p.a_dim * kSamp;
    p.off_state = off; off += 2u * p.state_floats * 4;
    off = (off + 15) / 16 * 16;
    p.off_rec = off; off += 4u * (uint32_t)p.n_blocks * 16;
    p.off_rrec = off; off += (uint32_t)p.ring_len * 8;
    off = (off + 1023) / 1024 * 1024;
    p.off_ring = off;
    int max_smem = 0, sms = 0;
    FRL_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, n0->cfg.device));
    FRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, n0->cfg.device));
    const int plan = lay.plan;
    p.slot_bytes = plan == 2 ? kSlotBytesH : plan == 0 ? kSlotBytesA : kSlotBytesB;
    p.halves = plan == 2 ? 1 : 0;
    int stages = ((int)max_smem - (int)off) / (int)p.slot_bytes;
    if (stages > kMaxSlots) stages = kMaxSlots;
    if (const char* cap = getenv("FRL_TC3_STAGES")) stages = std::min(stages, atoi(cap));   // experiments only
    if (stages < 4) return 1;   // not enough shared memory for the ring: the caller falls back to the 128-sample kernel
    p.n_stages = stages;
    const size_t smem_bytes = (size_t)off + (size_t)stages * p.slot_bytes;

    const long long tiles_per_role = (a.B + kSamp - 1) / kSamp;
    const long long n_quads = ((tiles_per_role + 3) / 4) * a.n_roles;
    const int grid = 2 * (int)std::min<long long>(n_quads, sms / 2);   // clusters of two CTAs

    if (p.x0_in_a) {
        const long long img_tiles = ((tiles_per_role + 3) / 4) * 4;   // padding tiles of the last quad included
        const size_t img_bytes = (size_t)p.K0p * kRows * 2, need = (size_t)img_tiles * img_bytes;
        frl_net* nm = const_cast<frl_net*>(n0);
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        FRL_CUDA(cudaStreamIsCapturing(stream, &cap));
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
        // the images are per launch: a launch on another stream (frl_score_host alternates two) waits for the previous reader
        if (!nm->x0s_event) FRL_CUDA(cudaEventCreateWithFlags(&nm->x0s_event, cudaEventDisableTiming));
        else if (cap == cudaStreamCaptureStatusNone) FRL_CUDA(cudaStreamWaitEvent(stream, nm->x0s_event, 0));
        const long long n_items = img_tiles * (p.K0p / 8) * kRows;
        const unsigned blocks = (unsigned)((n_items + 255) / 256);
        if (f16) tc3_state_image_kernel<__half><<<blocks, 256, 0, stream>>>(a.s, static_cast<uint8_t*>(nm->ws_x0s), a.B, img_tiles, p.s_dim, p.a_dim, p.K0p);
        else tc3_state_image_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(a.s, static_cast<uint8_t*>(nm->ws_x0s), a.B, img_tiles, p.s_dim, p.a_dim, p.K0p);
        FRL_CUDA(cudaGetLastError());
        p.x0s = static_cast<const uint8_t*>(nm->ws_x0s);
    }
    auto run = [&](const Params& q) {
        int rc;
        if (q.x0_in_a)
            rc = f16 ? launch_tc3<__half, true>(q, smem_bytes, grid, n0->cfg.device, stream)
                     : launch_tc3<__nv_bfloat16, true>(q, smem_bytes, grid, n0->cfg.device, stream);
        else
            rc = f16 ? launch_tc3<__half, false>(q, smem_bytes, grid, n0->cfg.device, stream)
                     : launch_tc3<__nv_bfloat16, false>(q, smem_bytes, grid, n0->cfg.device, stream);
        if (rc == FRL_OK && q.x0_in_a) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            FRL_CUDA(cudaStreamIsCapturing(stream, &cap));
            if (cap == cudaStreamCaptureStatusNone) FRL_CUDA(cudaEventRecord(n0->x0s_event, stream));
        }
        return rc;
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
