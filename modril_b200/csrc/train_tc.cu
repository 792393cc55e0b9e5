// tcgen05 tensor-core path of the conditional-flow-matching loss + gradient (sm_100a).
// Same contract as cfm_loss_grad_f32 (train_f32.cu): replaces  fm_loss(...) ; loss.backward()  of the reference
// (modril/flowril/pretrain.py:179-190, :94-109; modril/flowril/flowril.py:317-322), with fp16 operands and fp32
// accumulation in TMEM.
//
// Data layout.  Every activation / gradient matrix [B rows][F features] lives in HBM as 128-row tiles, each tile
// already the shared-memory image of a UMMA operand:  [tile][F/8][128 rows][8 x fp16]  (no swizzle; 8-row x 16-byte
// core matrices contiguous).  One such image serves
//   * as the K-major A operand of the forward / dgrad GEMMs (contraction over the features), and
//   * as the MN-major operand of the wgrad GEMMs (contraction over the 128 rows)
// without any transposition (tests/umma_probe_mn.cu), and a tile or a 64-feature slice of it is ONE contiguous
// cp.async.bulk copy.  Epilogues write tiles straight from registers: for a fixed 8-feature chunk the 32 rows of a
// warp are 512 contiguous bytes.
//
// Kernels
//   tc_prepare_kernel        X0 = [(1-t) a0 + t noise | s | t | 0] as fp16 tiles
//   tc_layer_kernel<EPI>     persistent GEMM over row tiles: D[128][N] = A_tile[128][K] * W^T, weights resident in
//                            shared memory, A streamed through a bulk-copy ring, two TMEM accumulator stages so the
//                            epilogue of tile i overlaps the MMAs of tile i+1.
//                              EPI_RELU  H_l  = relu(D + b)                         (forward)
//                              EPI_MASK  dZ_l = D * (H_l > 0)                       (dgrad)
//                              EPI_LOSS  V = D + b_heads -> loss term, dV           (heads forward + CFM loss)
//   tc_wgrad_kernel          dW = dZ^T X over a range of row tiles (both operands MN-major), all output features in
//                            one CTA (two M = 128 halves in TMEM); the otherwise idle epilogue warps form the bias
//                            gradient (column sums of dZ) from the same shared-memory tiles; partial sums per split
//   tc_reduce_kernel(s)      fixed-order reduction of the partials into the flat fp32 gradient (deterministic)
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "frl_internal.h"
#include "tc_common.cuh"

namespace frl {

namespace ttc {

using namespace tcc;

constexpr int kRows = 128;
constexpr int kChunkBytes = 16384;  // 64 features x 128 rows x 2 B
#ifndef FRL_TRAIN_EPI_WARPS
#define FRL_TRAIN_EPI_WARPS 16
#endif
constexpr int kLayerEpiWarps = FRL_TRAIN_EPI_WARPS;   // 8 or 16: four TMEM lane quarters x (2 or 4) column slices
// warp 0 producer, warp 1 MMA issuer + TMEM owner, then the epilogue warps, and one publisher warp (fused kernel: it releases
// the finished tiles to the next stage, so that the gpu-scope fence is off the epilogue warps' critical path)
constexpr int kLayerThreads = (3 + kLayerEpiWarps) * 32;
constexpr int kPubWarp = 2 + kLayerEpiWarps;
constexpr int kWgradThreads = 192;  // warp 0 producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue
constexpr int kMaxStages = 8;
enum { EPI_RELU = 0, EPI_MASK = 1, EPI_LOSS = 2 };

// One tile stream between two stages: R slots of `slot_bytes` (+ optionally `bits_words` mask words per slot).  With
// ready == done == nullptr it is a plain per-tile array (R = number of tiles) written by an earlier kernel.
struct Edge {
    uint8_t* tiles;            // [R][slot_bytes]
    uint32_t* bits;            // [R][bits_words] ReLU masks travelling with the tiles (may be null)
    unsigned* ready;           // [R] cumulative arrivals of the producer's epilogue warps (null: always ready)
    unsigned* done;            // [R] cumulative releases by the consumers                 (null: slots are never reused)
    int R;
    int bits_R;                // slots of the mask-bit array (a plain per-tile array in the fused kernel: the masks are read
                               // by the far end of the pipeline, and 4 KB per tile is cheap to keep for all tiles)
    unsigned ready_per_tile;   // arrivals that make one tile ready
    unsigned done_per_tile;    // releases that free one slot (sum over all consumers)
    uint32_t slot_bytes, bits_words;
    __device__ __forceinline__ int slot(int tile) const { return tile % R; }
    __device__ __forceinline__ unsigned use(int tile) const { return (unsigned)(tile / R); }
};

struct LayerParams {
    Edge in;                    // A tiles [K/8][128][8]
    Edge out;                   // out tiles [N/8][128][8] (+ mask bits written by EPI_RELU when out.bits != null)
    Edge mask;                  // EPI_MASK: the stream whose bits are applied ([N/32][128] words per tile: bits j / 16 + j of
                                // word (chunk, row) = activation (row, 32*chunk + 2j / 2j + 1) > 0)
    const uint8_t* w_img;       // [K/8][N][8]
    const float* bias;          // [N] (EPI_RELU); EPI_LOSS: bias of head 0 [a_dim] ...
    const float* bias2;         // ... and of head 1 [a_dim] (null for one head), straight from the flat parameters
    unsigned* err;              // time-out counter of the global waits (may be null)
    unsigned long long* trace;  // FRL_FUSED_TRACE: [CTA][8] cycle counters (null otherwise)
    int K, N, n_tiles, n_stages;
    int reverse;                // (unfused launches) walk the tiles from the last to the first: consecutive kernels
                                // alternate, so a kernel starts with the tiles its predecessor wrote last (still in L2)
    // EPI_LOSS: up to two terms (roles) stacked by tiles: term 1 starts at tile `tile_split` (== n_tiles if there is one)
    const float* a0[2];
    const float* noise[2];
    const float* t[2];
    float sign[2];
    float rel[2];           // dV multiplier of the term (its grad_scale / mean divisor relative to the final scale)
    long long B[2];
    int tile_split;
    int a_dim, n_heads;
    double* loss_partial;   // [n_tiles * 4]
    float* dvsum_partial;   // [n_tiles * 4][N]
};

// ---- global flags ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// slot give-back by a consumer: its copy of the tile has already landed in its shared memory (observed through the
// mbarrier), so later writes to the slot cannot reach it -- no fence needed (a release here would stall the MMA thread
// for a gpu-scope membar per tile)
__device__ __forceinline__ void red_relaxed_gpu(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {   // L2 (ring slots are rewritten: never a stale L1 line)
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Programmatic dependent launch: a kernel launched with the attribute may start while its predecessor in the stream is
// still running; pdl_wait() returns when the predecessor has completed (no-op for ordinary launches), pdl_launch() lets the
// successor's CTAs be scheduled as soon as SMs free up.  Everything before pdl_wait() must not depend on the predecessor.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
constexpr long long kSpinLimit = 1000ll * 1000 * 1000;   // ~0.5 s of SM clocks (a healthy update takes ~0.2 ms in total)
// wait until *p >= target (wrap-safe).  Gives up after kSpinLimit cycles and counts the time-out; once any wait of the
// launch has timed out, all the others return at once, so a broken hand-off costs one limit, not one per tile.
__device__ __noinline__ void wait_ge(const unsigned* p, unsigned target, unsigned* err, unsigned err0) {
    if ((int)(ld_acquire_gpu(p) - target) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu(p) - target) < 0) {
        __nanosleep(64);
        if (err && ld_acquire_gpu(err) != err0) return;
        if (clock64() - t0 > kSpinLimit) {
            if (err) atomicAdd(err, 1u);
            return;
        }
    }
}
__device__ __forceinline__ void wait_ready(const Edge& e, int tile, unsigned* err, unsigned err0) {
    if (e.ready) wait_ge(e.ready + e.slot(tile), (e.use(tile) + 1) * e.ready_per_tile, err, err0);
}
// The same with the counter value read one tile EARLIER (relaxed, so that the load thread does not sit out an L2 round trip
// per tile and stream): if the old value already says "published", an acquire fence orders the copies behind it.
__device__ __forceinline__ unsigned probe_ready(const Edge& e, int tile, int n_tiles) {
    return (e.ready && tile < n_tiles) ? ld_relaxed_gpu(e.ready + e.slot(tile)) : 0u;
}
// Relaxed polling (FRL_FUSED_POLL=1 builds use it; see the note at kRelaxedPoll).
__device__ __noinline__ void wait_ge_relaxed(const unsigned* p, unsigned target, unsigned* err, unsigned err0) {
    if ((int)(ld_relaxed_gpu(p) - target) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_relaxed_gpu(p) - target) < 0) {
        __nanosleep(32);
        if (err && ld_relaxed_gpu(err) != err0) return;
        if (clock64() - t0 > kSpinLimit) {
            if (err) atomicAdd(err, 1u);
            return;
        }
    }
}
// An acquire load (or an acquire fence) in the load thread costs 1.5 - 3 k cycles per tile here -- the membar it implies
// waits behind the 64 KB of epilogue stores the same SM has in flight -- which is as long as a whole tile takes.  With
// kRelaxedPoll the load thread polls the counter with relaxed gpu-scope loads (served by the L2) and issues the bulk copies
// behind the control dependency: the producer's publisher fenced and RELEASED the tile before it bumped the counter, so the
// data are in the L2 when the new counter value is; the copies read the L2 (never an L1 line), and fence.proxy.async
// orders them behind the poll in this thread.
#ifndef FRL_FUSED_RELAXED_POLL
#define FRL_FUSED_RELAXED_POLL 1
#endif
constexpr bool kRelaxedPoll = FRL_FUSED_RELAXED_POLL != 0;
__device__ __forceinline__ void wait_ready_probed(const Edge& e, int tile, unsigned probe, unsigned* err, unsigned err0) {
    if (!e.ready) return;
    const unsigned target = (e.use(tile) + 1) * e.ready_per_tile;
    if (kRelaxedPoll && e.done == nullptr) {   // (measured: helps the chains of mode 2, hurts the ring-reusing kernel of mode 1)
        if ((int)(probe - target) >= 0) return;
        wait_ge_relaxed(e.ready + e.slot(tile), target, err, err0);
    } else {
        wait_ge(e.ready + e.slot(tile), target, err, err0);
    }
}
// Slot reuse: the consumers bump `done` after their copy of the tile has LANDED in their shared memory, so nothing the
// producer writes afterwards can reach them: a relaxed poll is enough (the stores depend on it through control flow).
// `probe` is the counter value read one tile earlier (its latency is hidden behind that tile's epilogue).
__device__ __forceinline__ void wait_free(const Edge& e, int tile, unsigned probe, unsigned* err, unsigned err0) {
    if (!e.done) return;
    const unsigned target = e.use(tile) * e.done_per_tile;
    if ((int)(probe - target) >= 0) return;
    const unsigned* p = e.done + e.slot(tile);
    const long long t0 = clock64();
    while ((int)(ld_relaxed_gpu(p) - target) < 0) {
        __nanosleep(32);
        if (err && ld_relaxed_gpu(err) != err0) return;
        if (clock64() - t0 > kSpinLimit) {
            if (err) atomicAdd(err, 1u);
            return;
        }
    }
}

// X0 tiles: [tile][K0p/8][128][8];  one thread per (tile, 8-feature chunk, row).  Tiles [0, tile_split) hold term 0, the
// rest term 1 (each term starts on a tile boundary).
struct PrepTerm {
    const float *s, *a0, *t, *noise;
    long long B;
};
__global__ void tc_prepare_kernel(PrepTerm term0, PrepTerm term1, int tile_split, uint8_t* __restrict__ X0t, int n_tiles,
                                  int s_dim, int a_dim, int K0p) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int chunks = K0p / 8;
    if (i >= (long long)n_tiles * chunks * kRows) return;
    const int r = (int)(i % kRows);
    const int kc = (int)((i / kRows) % chunks);
    const long long tile = i / ((long long)kRows * chunks);
    const bool second = tile >= tile_split;
    const PrepTerm& tm = second ? term1 : term0;
    const float* __restrict__ s = tm.s;
    const float* __restrict__ a0 = tm.a0;
    const float* __restrict__ t = tm.t;
    const float* __restrict__ noise = tm.noise;
    const long long B = tm.B;
    const long long row = (tile - (second ? tile_split : 0)) * kRows + r;
    __half h[8];
    float tt = 0.f;
    if (row < B) tt = t[row];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kc * 8 + e;
        float v = 0.f;
        if (row < B) {
            if (k < a_dim)
                v = __fadd_rn(__fmul_rn(1.f - tt, a0[row * a_dim + k]), __fmul_rn(tt, noise[row * a_dim + k]));
            else if (k < a_dim + s_dim)
                v = s[row * s_dim + (k - a_dim)];
            else if (k == a_dim + s_dim)
                v = tt;
        }
        h[e] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
    *reinterpret_cast<uint4*>(X0t + (size_t)i * 16) = *reinterpret_cast<uint4*>(h);
}

// One GEMM stage over the row tiles rank, rank + pool, ... (rank = blockIdx.x, pool = gridDim.x for the stand-alone kernel)
template <int EPI>
__device__ __forceinline__ void layer_body(const LayerParams& p, const int rank, const int pool) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t bar0 = smem_u32(smem);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t w_full = bar0 + 8u * (2 * kMaxStages);
    // Accumulator stages: two of 256 columns, every epilogue warp works on every tile.  The heads stage (EPI_LOSS) has a
    // narrow accumulator and a long, latency-bound epilogue per row (loads of t / a0 / noise, powf, tanhf): FOUR stages of
    // 64 columns, and the four column-slice groups of epilogue warps each take every fourth tile, so four tiles' epilogues
    // run side by side.
    constexpr int kAcc = EPI == EPI_LOSS ? 4 : 2;
    constexpr uint32_t kAccCols = EPI == EPI_LOSS ? 64 : 256;
    constexpr int kArr = EPI == EPI_LOSS ? 4 : kLayerEpiWarps;   // epilogue warps per tile
    auto tmem_full = [&](int i) { return bar0 + 8u * (2 * kMaxStages + 1 + i); };
    auto tmem_empty = [&](int i) { return bar0 + 8u * (2 * kMaxStages + 5 + i); };
    auto stored = [&](int i) { return bar0 + 8u * (2 * kMaxStages + 9 + i); };       // epilogue warps -> publisher: tile written
    auto published = [&](int i) { return bar0 + 8u * (2 * kMaxStages + 13 + i); };   // publisher -> epilogue warps: barrier reusable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    float* sBias = reinterpret_cast<float*>(smem + 1024);   // [N <= 256]
    uint8_t* sW = smem + 2048;
    const uint32_t w_bytes = (uint32_t)p.K * p.N * 2;
    uint8_t* sRing = smem + 2048 + ((w_bytes + 1023) / 1024) * 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (p.K + 63) / 64;
    const unsigned err0 = p.err ? p.err[1] : 0u;   // the time-out count when this update was enqueued (copied by the host side memset chain)

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(w_full, 1);
        for (int i = 0; i < kAcc; ++i) {
            mbar_init(tmem_full(i), 1);
            mbar_init(tmem_empty(i), kArr);
            mbar_init(stored(i), kArr);
            mbar_init(published(i), 1);
        }
        fence_barrier_init();
        // the layer's weights do not depend on the previous kernel of the update: start their copy at once (this thread is
        // the load thread), before the programmatic-dependency wait below
        mbar_expect_tx(w_full, w_bytes);
        for (uint32_t off = 0; off < w_bytes; off += kChunkBytes)
            bulk_g2s(smem_u32(sW + off), p.w_img + off, min((uint32_t)kChunkBytes, w_bytes - off), w_full);
    }
    if (warp == 1) tmem_alloc512(smem_u32(tmem_slot));
    const bool publish = p.out.ready != nullptr || (EPI == EPI_MASK && p.mask.done != nullptr);
    if (EPI == EPI_RELU)
        for (int i = threadIdx.x; i < p.N; i += kLayerThreads) sBias[i] = p.bias[i];
    if (EPI == EPI_LOSS)   // head columns: [head 0 (a_dim) | head 1 (a_dim) | padding]
        for (int i = threadIdx.x; i < p.N; i += kLayerThreads)
            sBias[i] = i < p.a_dim ? p.bias[i] : (i < 2 * p.a_dim && p.bias2 ? p.bias2[i - p.a_dim] : 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch();
    pdl_wait();   // from here on: tiles written by the previous kernel, and tile buffers it may still be reading

    if (warp == 0) {
        // ------------------------------------------------ producer ------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            long long t_wait = 0, t_ring = 0;
            const long long t_begin = clock64();
            // (flagged streams are walked forwards: reverse == 0)
            unsigned pr_in = probe_ready(p.in, rank, p.n_tiles), pr_mask = EPI == EPI_MASK ? probe_ready(p.mask, rank, p.n_tiles) : 0u;
            for (int it = rank; it < p.n_tiles; it += pool) {
                const int tile = p.reverse ? p.n_tiles - 1 - it : it;
                // the tile (and, for the dgrad stages, the masks that its epilogue will read) has been published
                const long long tw = clock64();
                wait_ready_probed(p.in, tile, pr_in, p.err, err0);
                if (EPI == EPI_MASK) wait_ready_probed(p.mask, tile, pr_mask, p.err, err0);
                pr_in = probe_ready(p.in, it + pool, p.n_tiles);                         // next tile's counters: in flight while
                if (EPI == EPI_MASK) pr_mask = probe_ready(p.mask, it + pool, p.n_tiles);   // this tile's chunks are issued
                t_wait += clock64() - tw;
                if (p.in.ready) fence_proxy_async_all();   // other SMs' generic-proxy stores -> this bulk (async-proxy) read
                const uint8_t* src = p.in.tiles + (size_t)p.in.slot(tile) * p.in.slot_bytes;
                for (int c = 0; c < n_chunks; ++c) {
                    const uint32_t bytes = (uint32_t)min(64, p.K - 64 * c) * (kRows * 2);
                    const long long tr = clock64();
                    mbar_wait(empty(stage), phase ^ 1);
                    t_ring += clock64() - tr;
                    mbar_expect_tx(full(stage), bytes);
                    bulk_g2s(smem_u32(sRing + (size_t)stage * kChunkBytes), src + (size_t)c * kChunkBytes, bytes, full(stage));
                    if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                }
            }
            if (p.trace) {
                unsigned long long* tr = p.trace + (size_t)blockIdx.x * 8;
                tr[0] = (unsigned long long)(clock64() - t_begin);   // load thread: total, waiting for input tiles, waiting for a free smem slot
                tr[1] = (unsigned long long)t_wait;
                tr[2] = (unsigned long long)t_ring;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ----------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16((uint32_t)p.N, false, false);
            const uint32_t w_addr = smem_u32(sW), ring_addr = smem_u32(sRing);
            int stage = 0, n = 0;
            uint32_t phase = 0;
            mbar_wait(w_full, 0);
            long long t_acc = 0, t_full = 0;
            for (int it = rank; it < p.n_tiles; it += pool, ++n) {
                const int tile = p.reverse ? p.n_tiles - 1 - it : it;
                const int as = n % kAcc;
                const uint32_t aphase = (uint32_t)(n / kAcc) & 1u;
                const long long ta = clock64();
                mbar_wait(tmem_empty(as), aphase ^ 1);
                t_acc += clock64() - ta;
                tc_fence_after();
                for (int c = 0; c < n_chunks; ++c) {
                    const int nk = min(64, p.K - 64 * c) / 16;
                    const long long tf = clock64();
                    mbar_wait(full(stage), phase);
                    t_full += clock64() - tf;
                    // the whole input tile is in shared memory: its ring slot may be rewritten
                    if (c == n_chunks - 1 && p.in.done) red_relaxed_gpu(p.in.done + p.in.slot(tile), 1u);
                    tc_fence_after();
                    for (int j = 0; j < nk; ++j) {
                        const uint64_t da = make_desc(ring_addr + stage * kChunkBytes + j * (2 * kRows * 16), kRows * 16, 128);
                        const uint64_t db = make_desc(w_addr + (uint32_t)(c * 8 + 2 * j) * (uint32_t)p.N * 16, (uint32_t)p.N * 16, 128);
                        umma_f16(tmem + (uint32_t)as * kAccCols, da, db, idesc, (c | j) != 0);
                    }
                    umma_commit(empty(stage));
                    if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full(as));
            }
            if (p.trace) {
                unsigned long long* tr = p.trace + (size_t)blockIdx.x * 8;
                tr[3] = (unsigned long long)t_acc;    // MMA thread: waiting for a free accumulator (epilogue behind), for operand chunks
                tr[4] = (unsigned long long)t_full;
            }
        }
    } else if (warp == kPubWarp) {
        // ------------------------------------------------ publisher -----------------------------------------------
        // One thread hands every finished tile to the next stage: its release covers the 16 epilogue warps' stores (they
        // arrived on `stored`, which it acquired), and its gpu-scope fence costs the epilogue warps nothing.
        if (lane == 0 && publish) {
            int n = 0;
            for (int it = rank; it < p.n_tiles; it += pool, ++n) {
                const int tile = p.reverse ? p.n_tiles - 1 - it : it;
                const int as = n % kAcc;
                mbar_wait(stored(as), (uint32_t)(n / kAcc) & 1u);
                if (p.out.ready) red_release_gpu(p.out.ready + p.out.slot(tile), (unsigned)kLayerEpiWarps);
                if (EPI == EPI_MASK && p.mask.done) red_release_gpu(p.mask.done + p.mask.slot(tile), (unsigned)kLayerEpiWarps);
                mbar_arrive(published(as));
            }
        }
    } else {
        // ------------------------------------------------ epilogue ------------------------------------------------
        const int ew = warp - 2;
        const int q = warp & 3;      // TMEM lane quarter this warp may access
        const int ch = ew >> 2;      // column slice
        const int row = q * 32 + lane;
        long long t_free = 0, t_tm = 0, t_pub = 0;
        // local tile n of this CTA is row tile rank + n * pool; EPI_LOSS: warp group `ch` takes n = ch, ch + 4, ...
        const int n0 = EPI == EPI_LOSS ? ch : 0, dn = EPI == EPI_LOSS ? 4 : 1;
        unsigned probe = 0;   // lane 0: `done` counter of the NEXT tile's output slot, read one tile ahead
        if (p.out.done && lane == 0 && rank + n0 * pool < p.n_tiles) probe = ld_relaxed_gpu(p.out.done + p.out.slot(rank + n0 * pool));
        for (int n = n0, it = rank + n0 * pool; it < p.n_tiles; n += dn, it += dn * pool) {
            const int tile = p.reverse ? p.n_tiles - 1 - it : it;
            const int oslot = p.out.slot(tile);
            const int as = n % kAcc;
            const uint32_t aphase = (uint32_t)(n / kAcc) & 1u;
            const long long t0e = clock64();
            if (p.out.done) {   // every consumer of the tile that used this output slot before has copied it out
                if (lane == 0) {
                    wait_free(p.out, tile, probe, p.err, err0);
                    if (it + dn * pool < p.n_tiles) probe = ld_relaxed_gpu(p.out.done + p.out.slot(it + dn * pool));
                }
                __syncwarp();
            }
            const long long t1e = clock64();
            mbar_wait(tmem_full(as), aphase);
            t_free += t1e - t0e;
            t_tm += clock64() - t1e;
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)as * kAccCols;
            uint8_t* const out_tile = p.out.tiles + (size_t)oslot * p.out.slot_bytes;
            if (EPI == EPI_RELU || EPI == EPI_MASK) {
                const int half = p.N / (kLayerEpiWarps / 4);   // columns per slice
                uint8_t* out = out_tile + (size_t)row * 16;
                // ReLU masks travel as one bit per activation (4 KB per tile instead of re-reading the 64 KB tile)
                uint32_t* const bits_out = p.out.bits ? p.out.bits + (size_t)(tile % p.out.bits_R) * p.out.bits_words + row : nullptr;
                const uint32_t* const bits_in =
                    EPI == EPI_MASK ? p.mask.bits + (size_t)(tile % p.mask.bits_R) * p.mask.bits_words + row : nullptr;
                auto finish = [&](int c0, const uint32_t (&z)[32], uint32_t bits) {
                    uint32_t o[16];
                    if (EPI == EPI_RELU) {
                        const float4* b4 = reinterpret_cast<const float4*>(sBias + c0);   // broadcast vector loads
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 b = b4[j4];
                            o[2 * j4] = pack2_f16_relu(__uint_as_float(z[4 * j4]) + b.x, __uint_as_float(z[4 * j4 + 1]) + b.y);
                            o[2 * j4 + 1] = pack2_f16_relu(__uint_as_float(z[4 * j4 + 2]) + b.z, __uint_as_float(z[4 * j4 + 3]) + b.w);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            // H >= 0 after ReLU, so H > 0  <=>  (H + 0x7FFF) has bit 15 set (no carry between the halves):
                            // bit j = column 2j, bit 16 + j = column 2j + 1 of the 32-column chunk
                            bits |= ((o[j] + 0x7FFF7FFFu) >> (15 - j)) & (0x00010001u << j);
                        if (bits_out) bits_out[(size_t)(c0 / 32) * kRows] = bits;
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t keep = ((bits >> j) & 0x00010001u) * 0xFFFFu;   // 0x0001 / 0x10000 -> 0xFFFF masks
                            o[j] = pack2_f16(__uint_as_float(z[2 * j]), __uint_as_float(z[2 * j + 1])) & keep;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(out + (size_t)(c0 / 8 + i) * (kRows * 16)) =
                            make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                };
                // two 32-column chunks per round (one when a slice is only 32 columns wide: H = 128 with 16 epilogue warps):
                // both TMEM loads are in flight together
                const bool two = half >= 64;
                for (int c0 = ch * half; c0 < (ch + 1) * half; c0 += two ? 64 : 32) {
                    uint32_t z0[32], z1[32];
                    tmem_ld32(taddr + (uint32_t)c0, z0);
                    if (two) tmem_ld32(taddr + (uint32_t)c0 + 32, z1);
                    uint32_t bits0 = 0, bits1 = 0;
                    if (EPI == EPI_MASK) {
                        bits0 = ld_cg_u32(bits_in + (size_t)(c0 / 32) * kRows);
                        if (two) bits1 = ld_cg_u32(bits_in + (size_t)(c0 / 32 + 1) * kRows);
                    }
                    tmem_ld_wait_dep(z0);
                    if (two) tmem_ld_wait_dep(z1);
                    if (c0 + (two ? 64 : 32) >= (ch + 1) * half) {   // the accumulator stage is drained: the MMAs of tile i + 2 may start
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tmem_empty(as));
                    }
                    finish(c0, z0, bits0);
                    if (two) finish(c0 + 32, z1, bits1);
                }
            } else {
                // heads: V = D + b;  loss term and dL/dV (scaled by 2 w only: 1/B and the caller's grad_scale are applied
                // in the final reduction so that fp16 keeps the values at O(1))  -- pretrain.py:184-189
                float v[64];
                for (int g16 = 0; g16 < p.N; g16 += 16) {
                    uint32_t r16[16];
                    tmem_ld16(taddr + (uint32_t)g16, r16);
                    tmem_ld_wait_dep16(r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[g16 + i] = __uint_as_float(r16[i]) + sBias[g16 + i];
                }
                const int term_i = tile >= p.tile_split ? 1 : 0;
                const long long g = (long long)(tile - (term_i ? p.tile_split : 0)) * kRows + row;
                const float sign = p.sign[term_i];
                const int A = p.a_dim;
                double term = 0.0;
                if (g < p.B[term_i]) {
                    const float* noise = p.noise[term_i];
                    const float* a0 = p.a0[term_i];
                    const float w = powf(1.f - p.t[term_i][g], 1.5f);
                    const float w2 = 2.f * w * p.rel[term_i];
                    float acc = 0.f;
                    for (int j = 0; j < A; ++j) {
                        float pred = v[j], th = 0.f;
                        if (p.n_heads == 2) {
                            th = tanhf(v[A + j]);
                            pred += sign * th;
                        }
                        const float diff = pred - (noise[g * A + j] - a0[g * A + j]);
                        acc += diff * diff;
                        const float d = w2 * diff;
                        v[j] = d;
                        if (p.n_heads == 2) v[A + j] = d * sign * (1.f - th * th);
                    }
                    for (int j = p.n_heads * A; j < p.N; ++j) v[j] = 0.f;
                    term = (double)(w * acc);
                } else {
                    for (int j = 0; j < p.N; ++j) v[j] = 0.f;
                }
                uint8_t* out = out_tile + (size_t)row * 16;
                for (int c8 = 0; c8 < p.N; c8 += 8) {
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = pack2_f16(v[c8 + 2 * i], v[c8 + 2 * i + 1]);
                    *reinterpret_cast<uint4*>(out + (size_t)(c8 / 8) * (kRows * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
                    // bias gradient of the heads: column sums of the (fp16-rounded) dV over the warp's 32 rows
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __half2 hv = *reinterpret_cast<const __half2*>(&o[i]);
                        float lo = __low2float(hv), hi = __high2float(hv);
                        for (int s = 16; s > 0; s >>= 1) {
                            lo += __shfl_xor_sync(0xffffffffu, lo, s);
                            hi += __shfl_xor_sync(0xffffffffu, hi, s);
                        }
                        if (lane == 0) {
                            p.dvsum_partial[((size_t)tile * 4 + q) * p.N + c8 + 2 * i] = lo;
                            p.dvsum_partial[((size_t)tile * 4 + q) * p.N + c8 + 2 * i + 1] = hi;
                        }
                    }
                }
                for (int s = 16; s > 0; s >>= 1) term += __shfl_xor_sync(0xffffffffu, term, s);
                if (lane == 0) p.loss_partial[(size_t)tile * 4 + q] = term;
            }
            if (EPI == EPI_LOSS) {   // (the other epilogues release the accumulator stage as soon as it is in registers)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty(as));
            }
            if (publish) {
                // this warp's part of the output tile is written (and its mask words are read): tell the publisher
                const long long tp = clock64();
                __syncwarp();
                if (lane == 0) {
                    if (n >= kAcc) mbar_wait(published(as), (uint32_t)((n / kAcc - 1) & 1));
                    mbar_arrive(stored(as));
                }
                t_pub += clock64() - tp;
            }
        }
        if (p.trace && warp == 2 && lane == 0) {
            unsigned long long* tr = p.trace + (size_t)blockIdx.x * 8;
            tr[5] = (unsigned long long)t_free;   // epilogue warp 0: waiting for a free output slot, for the accumulator, in the fence
            tr[6] = (unsigned long long)t_tm;
            tr[7] = (unsigned long long)t_pub;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc512(tmem);
    }
}

template <int EPI>
__global__ void __launch_bounds__(kLayerThreads, 1) tc_layer_kernel(const __grid_constant__ LayerParams p) {
    layer_body<EPI>(p, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad:  partial[split][m][0..FQ) = sum_{tiles of the split} sum_b P[b][m] Q[b][.]   for all FP features m of P,
//         partial[split][m][FQ]    = sum_b P[b][m]                                      (bias gradient)
// One CTA owns all output features: two M = 128 halves accumulate in TMEM columns [0, FQ) and [256, 256 + FQ), so a
// row tile costs one P tile + one Q tile of traffic.  P travels in 128-feature halves through a 2-slot ring, Q through
// a 2-slot ring of whole tiles.  The four epilogue warps are idle until the end, so they form the bias column sums
// from the P halves in shared memory while the tensor pipe works on them.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxJobs = kMaxDepth + 1;
struct WgradJob {
    Edge P;                   // tiles [FP/8][128][8]
    Edge Q;                   // tiles [FQ/8][128][8]
    float* partial;           // [n_splits][FP][FQ + 16]
    unsigned* err;
    unsigned long long* trace;
    int FP, FQ;
};
struct WgradParams {
    WgradJob job[kMaxJobs];
    int first[kMaxJobs + 1];   // job j owns the CTAs [first[j], first[j + 1]): its splits, sized by the bytes it moves per tile
    int n_jobs, n_tiles;
};
constexpr uint32_t kWgPSlot = 32768, kWgQSlot = 65536;
constexpr uint32_t kWgPOff = 1024, kWgQOff = kWgPOff + 2 * kWgPSlot;
constexpr uint32_t kWgSmemBytes = kWgQOff + 2 * kWgQSlot;

// Split `rank` of `pool` accumulates the row tiles rank, rank + pool, ... (all of them resident in its TMEM) and writes
// partial[rank].  Uses the first kWgradThreads threads of the CTA.
__device__ __forceinline__ void wgrad_body(const WgradJob& jb, const int n_tiles, const int rank, const int pool) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t bar0 = smem_u32(smem);
    auto p_full = [&](int s) { return bar0 + 8u * s; };
    auto p_empty = [&](int s) { return bar0 + 8u * (2 + s); };
    auto q_full = [&](int s) { return bar0 + 8u * (4 + s); };
    auto q_empty = [&](int s) { return bar0 + 8u * (6 + s); };
    const uint32_t acc_full = bar0 + 8u * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    uint8_t* sP = smem + kWgPOff;
    uint8_t* sQ = smem + kWgQOff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = rank;
    const int t0 = rank, t1 = n_tiles, dt = pool;
    const int n_halves = jb.FP / kRows;
    const uint32_t q_tile_bytes = jb.Q.slot_bytes;
    const unsigned err0 = jb.err ? jb.err[1] : 0u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(p_full(s), 1);
            mbar_init(p_empty(s), 1 + 4);   // MMA commit + the four column-sum warps
            mbar_init(q_full(s), 1);
            mbar_init(q_empty(s), 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int ps = 0, qs = 0;
            uint32_t pph = 0, qph = 0;
            auto load_p = [&](int tile, int mh) {
                mbar_wait(p_empty(ps), pph ^ 1);
                mbar_expect_tx(p_full(ps), kWgPSlot);
                const uint8_t* src = jb.P.tiles + (size_t)jb.P.slot(tile) * jb.P.slot_bytes + (size_t)mh * kWgPSlot;
                bulk_g2s(smem_u32(sP + ps * kWgPSlot), src, 16384u, p_full(ps));
                bulk_g2s(smem_u32(sP + ps * kWgPSlot + 16384), src + 16384, 16384u, p_full(ps));
                if (++ps == 2) { ps = 0; pph ^= 1; }
            };
            long long t_wait = 0;
            const long long t_begin = clock64();
            unsigned pr_p = probe_ready(jb.P, t0, t1), pr_q = probe_ready(jb.Q, t0, t1);
            for (int tile = t0; tile < t1; tile += dt) {
                const long long tw = clock64();
                wait_ready_probed(jb.P, tile, pr_p, jb.err, err0);
                wait_ready_probed(jb.Q, tile, pr_q, jb.err, err0);
                pr_p = probe_ready(jb.P, tile + dt, t1);
                pr_q = probe_ready(jb.Q, tile + dt, t1);
                t_wait += clock64() - tw;
                if (jb.P.ready || jb.Q.ready) fence_proxy_async_all();
                load_p(tile, 0);
                mbar_wait(q_empty(qs), qph ^ 1);
                mbar_expect_tx(q_full(qs), q_tile_bytes);
                const uint8_t* qsrc = jb.Q.tiles + (size_t)jb.Q.slot(tile) * q_tile_bytes;
                for (uint32_t off = 0; off < q_tile_bytes; off += kChunkBytes)
                    bulk_g2s(smem_u32(sQ + qs * kWgQSlot + off), qsrc + off, min((uint32_t)kChunkBytes, q_tile_bytes - off),
                             q_full(qs));
                if (++qs == 2) { qs = 0; qph ^= 1; }
                for (int mh = 1; mh < n_halves; ++mh) load_p(tile, mh);
            }
            if (jb.trace) {
                unsigned long long* tr = jb.trace + (size_t)blockIdx.x * 8;
                tr[0] = (unsigned long long)(clock64() - t_begin);
                tr[1] = (unsigned long long)t_wait;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16((uint32_t)jb.FQ, true, true);
            const uint32_t p_addr = smem_u32(sP), q_addr = smem_u32(sQ);
            int ps = 0, qs = 0;
            uint32_t pph = 0, qph = 0;
            for (int tile = t0; tile < t1; tile += dt) {
                for (int mh = 0; mh < n_halves; ++mh) {
                    mbar_wait(p_full(ps), pph);
                    if (mh == 0) {
                        mbar_wait(q_full(qs), qph);
                        if (jb.Q.done) red_relaxed_gpu(jb.Q.done + jb.Q.slot(tile), 1u);   // copied out: the slot may be reused
                    }
                    if (mh == n_halves - 1 && jb.P.done) red_relaxed_gpu(jb.P.done + jb.P.slot(tile), 1u);
                    tc_fence_after();
                    for (int kk = 0; kk < kRows / 16; ++kk) {
                        // MN-major: 8-row groups are 128 B apart (LBO), 8-feature chunks 128*16 B apart (SBO)
                        const uint64_t da = make_desc(p_addr + ps * kWgPSlot + kk * 256, 128, kRows * 16);
                        const uint64_t db = make_desc(q_addr + qs * kWgQSlot + kk * 256, 128, kRows * 16);
                        umma_f16(tmem + (uint32_t)mh * 256, da, db, idesc, (tile != t0 || kk != 0));
                    }
                    umma_commit(p_empty(ps));
                    if (++ps == 2) { ps = 0; pph ^= 1; }
                }
                umma_commit(q_empty(qs));
                if (++qs == 2) { qs = 0; qph ^= 1; }
            }
            umma_commit(acc_full);
        }
    } else if (warp < kWgradThreads / 32) {
        const int q = warp & 3;
        const int et = (warp - 2) * 32 + lane;     // 0..127
        // column sums of P: thread (chunk = et / 8, g = et % 8) adds rows g, g + 8, ... of one 8-feature chunk
        const int chunk = et >> 3, g = et & 7;
        float bsum[2][8];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 8; ++e) bsum[h][e] = 0.f;
        int ps = 0;
        uint32_t pph = 0;
        for (int tile = t0; tile < t1; tile += dt) {
#pragma unroll
            for (int mh = 0; mh < 2; ++mh) {
                if (mh < n_halves) {
                    mbar_wait(p_full(ps), pph);
                    const uint8_t* src = sP + ps * kWgPSlot + chunk * (kRows * 16) + g * 16;
#pragma unroll 4
                    for (int r = 0; r < kRows / 8; ++r) {
                        const uint4 v = *reinterpret_cast<const uint4*>(src + r * 128);
                        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(hv[e]);
                            bsum[mh][2 * e] += f.x;
                            bsum[mh][2 * e + 1] += f.y;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(p_empty(ps));
                    if (++ps == 2) { ps = 0; pph ^= 1; }
                }
            }
        }
        const int ldp = jb.FQ + 16;
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = bsum[mh][e];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (g == 0 && mh < n_halves)
                    jb.partial[((size_t)split * jb.FP + mh * kRows + chunk * 8 + e) * ldp + jb.FQ] = v;
            }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int m = q * 32 + lane;
        for (int mh = 0; mh < n_halves; ++mh) {
            float* dst = jb.partial + ((size_t)split * jb.FP + mh * kRows + m) * ldp;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)mh * 256;
            for (int c0 = 0; c0 < jb.FQ; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                tmem_ld_wait_dep16(r);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(dst + c0 + 4 * i) =
                        make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                    __uint_as_float(r[4 * i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc512(tmem);
    }
}

__global__ void __launch_bounds__(kWgradThreads, 1) tc_wgrad_kernel(const __grid_constant__ WgradParams p) {
    int j = 0;
    while (j + 1 < p.n_jobs && (int)blockIdx.x >= p.first[j + 1]) ++j;
    wgrad_body(p.job[j], p.n_tiles, (int)blockIdx.x - p.first[j], p.first[j + 1] - p.first[j]);
}

// ---------------------------------------------------------------------------------------------------------------
// The fused update: every stage of the chain at once, one CTA per SM, tiles handed from stage to stage through L2 rings
// (see the header comment).  Stage s owns the CTAs [first[s], first[s + 1]).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxLayerStages = 2 * kMaxDepth + 2;   // forward 0..depth-1, heads + loss, dgrad heads, dgrad depth-1..1
struct FusedParams {
    LayerParams layer[kMaxLayerStages];
    WgradJob job[kMaxJobs];
    int epi[kMaxLayerStages];
    int first[kMaxLayerStages + kMaxJobs + 1];
    int n_layer_stages, n_jobs, n_tiles;
};

__global__ void __launch_bounds__(kLayerThreads, 1) tc_fused_kernel(const __grid_constant__ FusedParams fp) {
    const int n_stages = fp.n_layer_stages + fp.n_jobs;
    int s = 0;
    while (s + 1 < n_stages && (int)blockIdx.x >= fp.first[s + 1]) ++s;
    const int rank = (int)blockIdx.x - fp.first[s], pool = fp.first[s + 1] - fp.first[s];
    if (rank >= fp.n_tiles) return;   // fewer tiles than CTAs in this pool: nothing to do (its partial is not read)
    if (s < fp.n_layer_stages) {
        const int epi = fp.epi[s];
        if (epi == EPI_RELU) layer_body<EPI_RELU>(fp.layer[s], rank, pool);
        else if (epi == EPI_MASK) layer_body<EPI_MASK>(fp.layer[s], rank, pool);
        else layer_body<EPI_LOSS>(fp.layer[s], rank, pool);
    } else {
        wgrad_body(fp.job[s - fp.n_layer_stages], fp.n_tiles, rank, pool);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// reductions of the partials into the flat gradient (fixed order: deterministic)
// ---------------------------------------------------------------------------------------------------------------
struct RedParams {
    const float* part[kMaxJobs];   // per wgrad job
    int ldp[kMaxJobs];
    long long off_w[kMaxDepth], off_b[kMaxDepth], off_wh[2], off_bh[2];
    int depth, H, d_in, a_dim, n_heads;
    int n_splits[kMaxJobs];   // partial sums per wgrad job
    long long n_trunk;   // number of flat entries before the heads
    long long n_params;
    float scale;
    int accumulate;
};

__device__ __forceinline__ void reduce_main_body(const RedParams& rp, float* __restrict__ grad, const int block) {
    const long long e = block * (long long)blockDim.x + threadIdx.x;
    if (e >= rp.n_params) return;
    int job, m, col;
    if (e < rp.n_trunk) {
        int l = 0;
        while (l + 1 < rp.depth && rp.off_w[l + 1] <= e) ++l;
        const int K = l == 0 ? rp.d_in : rp.H;
        int n;
        if (e < rp.off_b[l]) {
            const long long i = e - rp.off_w[l];
            n = (int)(i / K);
            col = (int)(i % K);
        } else {
            n = (int)(e - rp.off_b[l]);
            col = rp.ldp[l] - 16;   // the column-sum column
        }
        job = l;
        m = n;
    } else {
        int h = (rp.n_heads == 2 && e >= rp.off_wh[1]) ? 1 : 0;
        if (e >= rp.off_bh[h]) return;   // head biases: reduce_small_body
        const long long i = e - rp.off_wh[h];
        const int j = (int)(i / rp.H), k = (int)(i % rp.H);
        job = rp.depth;
        m = k;
        col = h * rp.a_dim + j;
    }
    const float* P = rp.part[job];
    const int ldp = rp.ldp[job];
    float acc = 0.f;
    const int n_splits = rp.n_splits[job];
    for (int s = 0; s < n_splits; ++s) acc += P[((size_t)s * rp.H + m) * ldp + col];
    acc *= rp.scale;
    grad[e] = rp.accumulate ? grad[e] + acc : acc;
}

// block 0: loss;  block 1 + o: bias gradient of head output o
// (the loss partials of tiles [0, n_part0) belong to term 0, the rest to term 1)
struct SmallParams {
    const double* lpart;
    const float* dvsum;
    int n_part, n_part0, NOp;
    double inv_div0, inv_div1;
    float *loss0, *loss1;
};
// block 0: the two losses;  block 1 + o: bias gradient of head output o
__device__ __forceinline__ void reduce_small_body(const SmallParams& sp, const RedParams& rp, float* __restrict__ grad, const int block) {
    const double* __restrict__ lpart = sp.lpart;
    const float* __restrict__ dvsum = sp.dvsum;
    const int n_part = sp.n_part, n_part0 = sp.n_part0, NOp = sp.NOp;
    const double inv_div0 = sp.inv_div0, inv_div1 = sp.inv_div1;
    float* loss0 = sp.loss0;
    float* loss1 = sp.loss1;
    __shared__ double sh[256];
    double acc = 0.0;
    if (block == 0) {
        double acc1 = 0.0;
        for (int i = threadIdx.x; i < n_part0; i += 256) acc += lpart[i];
        for (int i = n_part0 + threadIdx.x; i < n_part; i += 256) acc1 += lpart[i];
        sh[threadIdx.x] = acc1;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0 && loss1) loss1[0] = (float)(sh[0] * inv_div1);
        __syncthreads();
    } else {
        const int o = block - 1;
        for (int i = threadIdx.x; i < n_part; i += 256) acc += (double)dvsum[(size_t)i * NOp + o];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    if (block == 0) {
        if (loss0) loss0[0] = (float)(sh[0] * inv_div0);
    } else if (grad) {
        const int o = block - 1;
        const int h = o / rp.a_dim, j = o % rp.a_dim;
        if (h < rp.n_heads) {
            const float v = (float)sh[0] * rp.scale;
            float* g = grad + rp.off_bh[h] + j;
            *g = rp.accumulate ? *g + v : v;
        }
    }
}

// ONE launch for the whole reduction: blocks [0, main_blocks) the weight / trunk-bias gradients, the rest the losses and
// the head-bias gradients
__global__ void __launch_bounds__(256) tc_reduce_kernel(const __grid_constant__ RedParams rp, const __grid_constant__ SmallParams sp,
                                                        float* __restrict__ grad, int main_blocks) {
    pdl_wait();
    if ((int)blockIdx.x < main_blocks) reduce_main_body(rp, grad, (int)blockIdx.x);
    else reduce_small_body(sp, rp, grad, (int)blockIdx.x - main_blocks);
}

// ---------------------------------------------------------------------------------------------------------------
// weight images (fp16) for the GEMMs above, rebuilt by frl_net_set_params
// ---------------------------------------------------------------------------------------------------------------
struct ImgJob {
    long long src_off;   // into the flat fp32 parameters
    int rs, cs;          // source strides of (row, feature)
    int R, F;            // valid rows / features of this job
    int row0, feat0;     // placement inside the image
    long long img_off;   // byte offset of the image
    int img_rows;        // rows of the image (N of the GEMM)
    long long first;     // prefix sum of R * F
};
struct ImgTable {
    ImgJob job[4 * kMaxDepth + 8];
    int n_jobs;
    long long total;
};
__global__ void tc_image_kernel(const float* __restrict__ flat, uint8_t* __restrict__ img, ImgTable tab) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tab.total; i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < tab.n_jobs && tab.job[j + 1].first <= i) ++j;
        const ImgJob& jb = tab.job[j];
        const long long e = i - jb.first;
        const int r = (int)(e / jb.F), f = (int)(e % jb.F);
        const float v = flat[jb.src_off + (long long)r * jb.rs + (long long)f * jb.cs];
        const int ri = jb.row0 + r, fi = jb.feat0 + f;
        __half* dst = reinterpret_cast<__half*>(img + jb.img_off) + ((size_t)(fi / 8) * jb.img_rows + ri) * 8 + (fi % 8);
        *dst = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
}

struct TrainPack {   // byte offsets into net->pack_train
    size_t wf[kMaxDepth];   // forward  image of layer l: [Kp_l/8][H][8]      (N = H outputs,  K = Kp_l inputs)
    size_t wd[kMaxDepth];   // dgrad    image of layer l: [H/8][H][8]         (N = H inputs,   K = H outputs), l >= 1
    size_t whf, whd;        // heads forward [H/8][NOp][8], heads dgrad [NOp/8][H][8]
    size_t total;
};
static TrainPack train_pack_layout(const frl_net* net) {
    TrainPack tp{};
    const int H = net->cfg.hidden, depth = net->cfg.depth;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
    for (int l = 0; l < depth; ++l) tp.wf[l] = take((size_t)(l == 0 ? net->d_in_p : H) * H * 2);
    for (int l = 1; l < depth; ++l) tp.wd[l] = take((size_t)H * H * 2);
    tp.whf = take((size_t)H * net->n_out_p * 2);
    tp.whd = take((size_t)net->n_out_p * H * 2);
    tp.total = off;
    return tp;
}

}  // namespace ttc

int tc_train_pack(frl_net* net, const float* params_dev, cudaStream_t st) {
    using namespace ttc;
    const TrainPack tp = train_pack_layout(net);
    if (!net->pack_train) {   // the padding of the images is written once: re-packs only rewrite the weights
        FRL_CUDA(cudaMalloc(&net->pack_train, tp.total));
        FRL_CUDA(cudaMemsetAsync(net->pack_train, 0, tp.total, st));
    }
    const int H = net->cfg.hidden, depth = net->cfg.depth, A = net->cfg.a_dim;
    ImgTable tab{};
    long long total = 0;
    auto add = [&](long long src_off, int rs, int cs, int R, int F, int row0, int feat0, size_t img_off, int img_rows) {
        ImgJob& j = tab.job[tab.n_jobs++];
        j = ImgJob{src_off, rs, cs, R, F, row0, feat0, (long long)img_off, img_rows, total};
        total += (long long)R * F;
    };
    for (int l = 0; l < depth; ++l) {
        const int K = l == 0 ? net->d_in : H;
        add(net->off_w[l], K, 1, H, K, 0, 0, tp.wf[l], H);              // rows = outputs n, features = inputs k
        if (l >= 1) add(net->off_w[l], 1, K, H, H, 0, 0, tp.wd[l], H);  // rows = inputs k, features = outputs n
    }
    for (int h = 0; h < net->cfg.n_heads; ++h) {
        add(net->off_wh[h], H, 1, A, H, h * A, 0, tp.whf, net->n_out_p);   // rows = outputs o, features = k
        add(net->off_wh[h], 1, H, H, A, 0, h * A, tp.whd, H);              // rows = k, features = outputs o
    }
    tab.total = total;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 1184);
    tc_image_kernel<<<blocks, 256, 0, st>>>(params_dev, static_cast<uint8_t*>(net->pack_train), tab);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the opt-in maximum
static cudaError_t set_smem_attr_impl(const void* kern, int device) {
    struct Seen { const void* k; int dev; };
    static Seen seen[64];
    static int n_seen = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_seen; ++i)
        if (seen[i].k == kern && seen[i].dev == device) return cudaSuccess;
    int max_smem = 0;
    cudaError_t e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e == cudaSuccess && n_seen < 64) seen[n_seen++] = Seen{kern, device};
    return e;
}
template <typename K>
static cudaError_t set_smem_attr(K kern, int device, int /*bytes*/) {
    return set_smem_attr_impl(reinterpret_cast<const void*>(kern), device);
}

struct DevInfo { int sms, max_smem; };
static int dev_info(int device, DevInfo* out) {
    static DevInfo cache[64];
    static bool have[64] = {};
    if (device >= 0 && device < 64 && have[device]) { *out = cache[device]; return FRL_OK; }
    DevInfo d{};
    FRL_CUDA(cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    FRL_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device));
    if (device >= 0 && device < 64) { cache[device] = d; have[device] = true; }
    *out = d;
    return FRL_OK;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait): the kernel may be scheduled while its
// predecessor in the stream is finishing; inside a stream capture this becomes a programmatic dependency edge of the graph.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// (fused path) how many CTAs each stage gets: proportional to its estimated cycles per tile, at least one each
static void size_pools(const std::vector<double>& cost, int budget, std::vector<int>* pool) {
    const int n = (int)cost.size();
    pool->assign(n, 1);
    double total = 0;
    for (double c : cost) total += c;
    int used = n;
    for (int i = 0; i < n; ++i) {
        const int extra = std::max(0, (int)(cost[i] / total * budget) - 1);
        (*pool)[i] += extra;
        used += extra;
    }
    while (used < budget) {   // leftovers to the stages with the most work per CTA
        int best = 0;
        for (int i = 1; i < n; ++i)
            if (cost[i] / (*pool)[i] > cost[best] / (*pool)[best]) best = i;
        ++(*pool)[best];
        ++used;
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int cfm_loss_grad_tc(frl_net* net, const CfmTerm* terms, int n_terms, int accumulate, float* grad, cudaStream_t st) {
    using namespace ttc;
    const int H = net->cfg.hidden, A = net->cfg.a_dim, depth = net->cfg.depth;
    const int K0p = net->d_in_p, NOp = net->n_out_p;
    FRL_REQUIRE(n_terms == 1 || n_terms == 2, "one or two loss terms");
    long long B_total = 0;
    for (int i = 0; i < n_terms; ++i) {
        FRL_REQUIRE(terms[i].B > 0 && terms[i].B < (1ll << 30), "batch size out of range");
        B_total += terms[i].B;
    }
    FRL_REQUIRE(B_total < (1ll << 30), "batch size out of range");
    {
        int rcp = ensure_pack(net, PACK_TRAIN, st);   // fp16 GEMM images
        if (rcp != FRL_OK) return rcp;
    }
    FRL_REQUIRE(net->pack_train != nullptr, "tensor-core training weights missing");
    FRL_REQUIRE(NOp <= 64, "a_dim too large for the tensor-core training path");
    // every term starts on a tile boundary; the final reduction applies the scale of the first term with a non-zero
    // scale, the loss epilogue multiplies the other term's dV by the ratio
    int nt_of[2] = {0, 0};
    double scale_of[2] = {0.0, 0.0}, inv_div[2] = {0.0, 0.0};
    for (int i = 0; i < n_terms; ++i) {
        nt_of[i] = (int)((terms[i].B + kRows - 1) / kRows);
        const double div = (double)(terms[i].mean_divisor > 0 ? terms[i].mean_divisor : terms[i].B);
        inv_div[i] = 1.0 / div;
        scale_of[i] = (double)terms[i].grad_scale / div;
    }
    const int nt = nt_of[0] + nt_of[1];
    const double final_scale = scale_of[0] != 0.0 ? scale_of[0] : (scale_of[1] != 0.0 ? scale_of[1] : 1.0);
    const int device = net->cfg.device;
    const TrainPack tp = train_pack_layout(net);
    const uint8_t* pk = static_cast<const uint8_t*>(net->pack_train);

    DevInfo di{};
    int rc = dev_info(device, &di);
    if (rc != FRL_OK) return rc;
    const int sms = di.sms;
    const int n_jobs = depth + 1;                 // trunk layers + heads
    const int n_layer_stages = 2 * depth + 1;     // forward 0..depth-1, heads + loss, dgrad heads, dgrad depth-1..1
    const int n_stages_total = n_layer_stages + n_jobs;
    // How the GEMM part is launched (FRL_TRAIN_FUSED):
    //   0  one kernel per layer GEMM (11 launches), every tile through HBM between them;
    //   1  ONE dataflow kernel for everything (tiles in L2 rings; profiles/r02_fused_train.md: 210 B of HBM per row, but ten
    //      stages deep with 7-12 CTAs per stage it fills, drains and balances badly: 0.42 ms against 0.33 ms);
    //   2  two SHORT dataflow kernels -- the forward chain [F_0 .. F_{depth-1}, heads + loss] and the dgrad chain
    //      [dgrad heads, D_{depth-1} .. D_1] -- each on all SMs, then the stand-alone wgrad kernel.  Inside a chain a tile
    //      goes to the next stage through the L2 (ready flags, no slot reuse: every tile is also written out for wgrad and
    //      the masks), so the chains read almost nothing from HBM and run L2-fed; only wgrad streams the tiles back in.
    const int want_mode = env_int("FRL_TRAIN_FUSED", 0);
    const int mode = (grad != nullptr && sms >= 2 * n_stages_total) ? want_mode : 0;
    // Per-layer launches chained by programmatic dependent launch when the update is small (measured on a B200, maze: the
    // reference's own batch of 2 x 128 rows 0.094 -> 0.083 ms per update, but 2 x 65536 rows 0.334 -> 0.348 ms: with ~7
    // tiles per CTA the launch gaps are already hidden and the early weight copies only compete with the predecessor's tail).
    // FRL_TRAIN_PDL=0 / 1 forces it off / on.
    const bool use_pdl = env_int("FRL_TRAIN_PDL", nt <= 64 ? 1 : 0) != 0;
    bool pdl_reduce = false;
    const bool fused = mode == 1;        // everything in one kernel, ring slots reused
    const bool chains = mode == 2;       // two chain kernels + wgrad kernel, plain per-tile arrays with ready flags

    // ---- stage pools (dataflow kernels) / split count (per-layer launches) ----
    // stage order: [F_0 .. F_{depth-1}] [LOSS] [DH] [D_{depth-1} .. D_1] [W_0 .. W_{depth-1}] [W_heads]
    std::vector<int> pool(n_stages_total, 1);
    if (fused || chains) {
        std::vector<double> cost(n_stages_total, 0.0);
        // cycles-per-tile estimates (relative weights): measured per-tile times of the stages on a B200
        const double heavy = env_int("FRL_FUSED_COST_GEMM", 2850), wg = env_int("FRL_FUSED_COST_WGRAD", 2500);
        const double c_l0 = env_int("FRL_FUSED_COST_L0", 2500), c_loss = env_int("FRL_FUSED_COST_LOSS", 2400);
        const double c_dh = env_int("FRL_FUSED_COST_DH", 1450), c_w0 = env_int("FRL_FUSED_COST_W0", 2050);
        const double c_wh = env_int("FRL_FUSED_COST_WH", 1950);
        for (int l = 0; l < depth; ++l) cost[l] = l == 0 ? c_l0 + heavy * K0p / H : heavy;
        cost[depth] = c_loss;
        cost[depth + 1] = c_dh + heavy * NOp / H;
        for (int l = depth - 1; l >= 1; --l) cost[depth + 2 + (depth - 1 - l)] = heavy;
        for (int l = 0; l < depth; ++l) cost[n_layer_stages + l] = l == 0 ? c_w0 + wg * K0p / H : wg;
        cost[n_layer_stages + depth] = c_wh + wg * NOp / H;
        const int budget = std::min(sms, env_int("FRL_FUSED_CTAS", sms));
        if (fused) {
            size_pools(cost, budget, &pool);
        } else {   // every chain gets all the SMs
            std::vector<int> part;
            const std::vector<double> fwd(cost.begin(), cost.begin() + depth + 1), bwd(cost.begin() + depth + 1, cost.begin() + n_layer_stages);
            size_pools(fwd, budget, &part);
            std::copy(part.begin(), part.end(), pool.begin());
            size_pools(bwd, budget, &part);
            std::copy(part.begin(), part.end(), pool.begin() + depth + 1);
        }
    }
    // per-layer launches: the wgrad kernel is HBM-bound, so a job gets CTAs in proportion to the bytes it reads per tile (a
    // trunk layer: two 64 KB tiles; layer 0 and the heads: one 64 KB tile and a narrow one)
    int n_splits[kMaxJobs];
    if (fused) {
        for (int j = 0; j < n_jobs; ++j) n_splits[j] = std::min(pool[n_layer_stages + j], nt);
    } else {
        std::vector<double> bytes(n_jobs);   // (also mode 2: its wgrad is the stand-alone kernel)
        for (int j = 0; j < n_jobs; ++j) bytes[j] = (double)H + (j == 0 ? K0p : (j == depth ? NOp : H));
        std::vector<int> share;
        size_pools(bytes, std::max(n_jobs, sms), &share);
        for (int j = 0; j < n_jobs; ++j) n_splits[j] = std::max(1, std::min(share[j], nt));
    }

    // ---- rings (fused): slots per stream ~ (stages between the producer and the last consumer + 2) x tiles in flight per stage
    int maxpool = 1;
    for (int v : pool) maxpool = std::max(maxpool, v);
    const int inflight = std::max(4, (maxpool + 4) * env_int("FRL_FUSED_RING_PCT", 100) / 100);
    int R_h[kMaxDepth], R_dz[kMaxDepth];
    for (int l = 0; l < depth; ++l) {
        R_h[l] = fused ? std::min(nt, (2 * (depth - l) + 2) * inflight) : nt;
        R_dz[l] = fused ? std::min(nt, 3 * inflight) : nt;
    }

    // ---- workspace (bytes) ----
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
    const size_t tileH = (size_t)H * kRows * 2, tileX = (size_t)K0p * kRows * 2, tileV = (size_t)NOp * kRows * 2;
    const size_t bitsW = (size_t)(H / 32) * kRows;   // mask words per tile
    const size_t o_err = take(2 * sizeof(unsigned));   // [time-out count][its value when this update was enqueued]: first, so it stays put
    const size_t o_x0 = take(tileX * nt);
    size_t o_h[kMaxDepth], o_dz[kMaxDepth], o_bits[kMaxDepth];
    for (int l = 0; l < depth; ++l) o_h[l] = take(tileH * R_h[l]);
    for (int l = 0; l < depth; ++l) o_dz[l] = take(tileH * R_dz[l]);
    for (int l = 0; l < depth; ++l) o_bits[l] = take((size_t)nt * bitsW * sizeof(uint32_t));
    const size_t o_dv = take(tileV * nt);
    const size_t o_lp = take((size_t)nt * 4 * sizeof(double));
    const size_t o_dvs = take((size_t)nt * 4 * NOp * sizeof(float));
    size_t o_part[kMaxJobs];
    int ldp[kMaxJobs];
    for (int j = 0; j < n_jobs; ++j) {
        ldp[j] = (j == 0 ? K0p : (j == depth ? NOp : H)) + 16;
        o_part[j] = take((size_t)n_splits[j] * H * ldp[j] * sizeof(float));
    }
    // flags: ready / done counters of every stream
    size_t n_flags = 0;
    size_t f_h[kMaxDepth], f_dz[kMaxDepth], f_dv = 0;
    if (fused || chains) {
        for (int l = 0; l < depth; ++l) { f_h[l] = n_flags; n_flags += 2 * (size_t)R_h[l]; }
        for (int l = 0; l < depth; ++l) { f_dz[l] = n_flags; n_flags += 2 * (size_t)R_dz[l]; }
        f_dv = n_flags; n_flags += (size_t)nt;
    }
    const size_t o_flags = take(n_flags * sizeof(unsigned));
    if (net->ws_tc_bytes < off) {
        if (net->ws_tc) net->retired.push_back(net->ws_tc);   // kept alive: see frl_net::retired
        net->ws_tc = nullptr;
        net->ws_tc_bytes = 0;
        FRL_CUDA(cudaMalloc(&net->ws_tc, off));
        net->ws_tc_bytes = off;
        net->train_flags = nullptr;
    }
    uint8_t* ws = static_cast<uint8_t*>(net->ws_tc);
    double* lpart = reinterpret_cast<double*>(ws + o_lp);
    float* dvsum = reinterpret_cast<float*>(ws + o_dvs);
    unsigned* flags = reinterpret_cast<unsigned*>(ws + o_flags);
    unsigned* err = reinterpret_cast<unsigned*>(ws + o_err);
    if (net->train_flags != err) {   // new workspace: the time-out count starts at 0
        FRL_CUDA(cudaMemsetAsync(err, 0, 2 * sizeof(unsigned), st));
        net->train_flags = err;
    }
    if (fused || chains) {
        FRL_CUDA(cudaMemsetAsync(flags, 0, n_flags * sizeof(unsigned), st));
        FRL_CUDA(cudaMemcpyAsync(err + 1, err, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
    }

    // FRL_FUSED_TRACE=file: per-CTA cycle counters of one launch (synchronises; debugging only)
    const char* trace_path = (fused || chains) ? getenv("FRL_FUSED_TRACE") : nullptr;
    unsigned long long* trace_dev = nullptr;
    if (trace_path && *trace_path) {
        FRL_CUDA(cudaMalloc(&trace_dev, (size_t)sms * 8 * sizeof(unsigned long long)));
        FRL_CUDA(cudaMemsetAsync(trace_dev, 0, (size_t)sms * 8 * sizeof(unsigned long long), st));
    }

    // ---- the streams ----
    constexpr unsigned kW = (unsigned)kLayerEpiWarps;   // arrivals of a layer stage per tile
    auto plain = [&](size_t o_tiles, size_t slot_bytes, int R) {
        Edge e{};
        e.tiles = ws + o_tiles; e.R = R; e.slot_bytes = (uint32_t)slot_bytes; e.ready_per_tile = kW; e.done_per_tile = 1;
        return e;
    };
    const Edge e_x0 = plain(o_x0, tileX, nt);
    Edge e_h[kMaxDepth], e_dz[kMaxDepth], e_dv = plain(o_dv, tileV, nt);
    for (int l = 0; l < depth; ++l) {
        e_h[l] = plain(o_h[l], tileH, R_h[l]);
        e_h[l].bits = grad ? reinterpret_cast<uint32_t*>(ws + o_bits[l]) : nullptr;
        e_h[l].bits_words = (uint32_t)bitsW;
        e_h[l].bits_R = nt;
        e_dz[l] = plain(o_dz[l], tileH, R_dz[l]);
        if (fused || chains) {
            e_h[l].ready = flags + f_h[l];
            e_dz[l].ready = flags + f_dz[l];
        }
        if (fused) {   // ring slots are reused: the consumers give them back
            e_h[l].done = flags + f_h[l] + R_h[l];
            e_h[l].done_per_tile = 2;   // next forward stage (or heads) and wgrad of the next layer (or heads)
            e_dz[l].done = flags + f_dz[l] + R_dz[l];
            e_dz[l].done_per_tile = l >= 1 ? 2 : 1;   // dgrad through layer l (l >= 1) and wgrad of layer l
        }
    }
    if (fused || chains) e_dv.ready = flags + f_dv;   // one slot per tile: never reused

    // ---- stage parameters ----
    LayerParams LP[kMaxLayerStages] = {};
    int epi_of[kMaxLayerStages];
    int n_launch = 0;   // (per-layer launches) the kernels alternate their tile order (see LayerParams::reverse)
    auto finish_layer = [&](LayerParams& p, int epi, int idx) -> int {
        p.n_tiles = nt;
        p.err = (fused || chains) ? err : nullptr;
        p.trace = trace_dev;
        p.reverse = (fused || chains) ? 0 : ((n_launch++) & 1);
        const size_t base = 2048 + ((size_t)p.K * p.N * 2 + 1023) / 1024 * 1024;
        int stages = (int)(((size_t)di.max_smem - base) / kChunkBytes);
        stages = std::min(stages, kMaxStages);
        if (fused || chains) stages = std::min(stages, std::max(2, env_int("FRL_FUSED_SMEM_STAGES", kMaxStages)));
        if (stages < 2) {
            set_error("tensor-core training: layer does not fit in shared memory");
            return FRL_ERR_UNSUPPORTED;
        }
        p.n_stages = stages;
        epi_of[idx] = epi;
        return FRL_OK;
    };
    for (int l = 0; l < depth; ++l) {   // forward
        LayerParams& p = LP[l];
        p.in = l == 0 ? e_x0 : e_h[l - 1];
        p.out = e_h[l];
        p.w_img = pk + tp.wf[l];
        p.bias = net->params_dev + net->off_b[l];   // biases are read from the flat parameters (no fp32 pack needed)
        p.K = l == 0 ? K0p : H;
        p.N = H;
        if ((rc = finish_layer(p, EPI_RELU, l)) != FRL_OK) return rc;
    }
    {   // heads + loss
        LayerParams& p = LP[depth];
        p.in = e_h[depth - 1];
        p.out = e_dv;
        p.w_img = pk + tp.whf;
        p.bias = net->params_dev + net->off_bh[0];
        p.bias2 = net->cfg.n_heads == 2 ? net->params_dev + net->off_bh[1] : nullptr;
        p.K = H;
        p.N = NOp;
        for (int i = 0; i < 2; ++i) {
            const CfmTerm& tm = terms[i < n_terms ? i : 0];
            p.a0[i] = tm.a0; p.noise[i] = tm.noise; p.t[i] = tm.t; p.sign[i] = tm.sign; p.B[i] = i < n_terms ? tm.B : 0;
            p.rel[i] = (float)(scale_of[i] / final_scale);
        }
        p.tile_split = nt_of[0];
        p.a_dim = A; p.n_heads = net->cfg.n_heads;
        p.loss_partial = lpart;
        p.dvsum_partial = dvsum;
        if ((rc = finish_layer(p, EPI_LOSS, depth)) != FRL_OK) return rc;
    }
    {   // dgrad through the heads: dZ_{depth-1} = (dV Wh) * mask_{depth-1}
        LayerParams& p = LP[depth + 1];
        p.in = e_dv;
        p.out = e_dz[depth - 1];
        p.mask = e_h[depth - 1];
        p.mask.done = nullptr;   // (the masks live in a per-tile array: nothing to give back)
        p.w_img = pk + tp.whd;
        p.K = NOp;
        p.N = H;
        if ((rc = finish_layer(p, EPI_MASK, depth + 1)) != FRL_OK) return rc;
    }
    for (int l = depth - 1; l >= 1; --l) {   // dgrad through layer l: dZ_{l-1} = (dZ_l W_l) * mask_{l-1}
        const int idx = depth + 2 + (depth - 1 - l);
        LayerParams& p = LP[idx];
        p.in = e_dz[l];
        p.out = e_dz[l - 1];
        p.mask = e_h[l - 1];
        p.mask.done = nullptr;
        p.w_img = pk + tp.wd[l];
        p.K = H;
        p.N = H;
        if ((rc = finish_layer(p, EPI_MASK, idx)) != FRL_OK) return rc;
    }
    WgradJob WJ[kMaxJobs] = {};
    for (int l = 0; l < n_jobs; ++l) {
        WgradJob& jb = WJ[l];
        jb.P = l == depth ? e_h[depth - 1] : e_dz[l];
        jb.FP = H;
        if (l == 0) { jb.Q = e_x0; jb.FQ = K0p; }
        else if (l == depth) { jb.Q = e_dv; jb.FQ = NOp; }
        else { jb.Q = e_h[l - 1]; jb.FQ = H; }
        jb.partial = reinterpret_cast<float*>(ws + o_part[l]);
        jb.err = (fused || chains) ? err : nullptr;
        jb.trace = trace_dev;
    }

    RedParams rp{};
    rp.depth = depth; rp.H = H; rp.d_in = net->d_in; rp.a_dim = A; rp.n_heads = net->cfg.n_heads;
    for (int j = 0; j < n_jobs; ++j) rp.n_splits[j] = n_splits[j];
    for (int l = 0; l < depth; ++l) { rp.off_w[l] = net->off_w[l]; rp.off_b[l] = net->off_b[l]; }
    for (int h = 0; h < net->cfg.n_heads; ++h) { rp.off_wh[h] = net->off_wh[h]; rp.off_bh[h] = net->off_bh[h]; }
    rp.n_trunk = net->off_wh[0];
    rp.n_params = net->n_params;
    rp.scale = (float)final_scale;
    rp.accumulate = accumulate;
    for (int j = 0; j < n_jobs; ++j) { rp.part[j] = reinterpret_cast<const float*>(ws + o_part[j]); rp.ldp[j] = ldp[j]; }
    const SmallParams small{lpart, dvsum, nt * 4, nt_of[0] * 4, NOp, inv_div[0], inv_div[1], terms[0].loss,
                            n_terms > 1 ? terms[1].loss : nullptr};

    // ---- X0 tiles ----
    {
        const CfmTerm& t1 = terms[n_terms - 1];
        const PrepTerm p0{terms[0].s, terms[0].a0, terms[0].t, terms[0].noise, terms[0].B};
        const PrepTerm p1{t1.s, t1.a0, t1.t, t1.noise, t1.B};
        const long long n = (long long)nt * (K0p / 8) * kRows;
        tc_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p0, p1, nt_of[0], ws + o_x0, nt, net->cfg.s_dim, A, K0p);
        FRL_CUDA(cudaGetLastError());
    }

    auto stage_name = [&](int i, char* name, size_t n) {
        if (i < depth) snprintf(name, n, "F%d", i);
        else if (i == depth) snprintf(name, n, "LOSS");
        else if (i == depth + 1) snprintf(name, n, "DH");
        else if (i < n_layer_stages) snprintf(name, n, "D%d", depth - 1 - (i - depth - 2));
        else snprintf(name, n, "W%d", i - n_layer_stages);
    };
    // FRL_FUSED_TRACE: cycle counters of the stages [s0, s1) of the dataflow kernel just launched (synchronises)
    auto dump_trace = [&](int s0, int s1, const int* first, const char* fmode) -> int {
        if (!trace_dev) return FRL_OK;
        std::vector<unsigned long long> h((size_t)sms * 8);
        FRL_CUDA(cudaStreamSynchronize(st));
        FRL_CUDA(cudaMemcpy(h.data(), trace_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        FRL_CUDA(cudaMemset(trace_dev, 0, h.size() * sizeof(unsigned long long)));
        if (FILE* f = fopen(trace_path, fmode)) {
            fprintf(f, "# tiles %d; stage: F<l> forward, LOSS, DH dgrad heads, D<l> dgrad through layer l, W<l> wgrad (W%d = heads)\n", nt, depth);
            fprintf(f, "# cta stage rank pool | load thread: total wait_in wait_smem_slot | mma: wait_acc wait_chunks | epilogue warp 0: wait_out_slot wait_acc publish\n");
            for (int i = s0; i < s1; ++i) {
                char name[16];
                stage_name(i, name, sizeof name);
                for (int r = 0; r < pool[i]; ++r) {
                    const unsigned long long* t = &h[(size_t)(first[i - s0] + r) * 8];
                    fprintf(f, "%3d %-4s %2d %2d | %8llu %8llu %8llu | %8llu %8llu | %8llu %8llu %8llu\n", first[i - s0] + r, name, r, pool[i],
                            t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
                }
            }
            fclose(f);
        }
        return FRL_OK;
    };
    if (fused) {
        FusedParams fp{};
        fp.n_layer_stages = n_layer_stages; fp.n_jobs = n_jobs; fp.n_tiles = nt;
        for (int i = 0; i < n_layer_stages; ++i) { fp.layer[i] = LP[i]; fp.epi[i] = epi_of[i]; }
        for (int j = 0; j < n_jobs; ++j) fp.job[j] = WJ[j];
        int first = 0;
        for (int i = 0; i < n_stages_total; ++i) { fp.first[i] = first; first += pool[i]; }
        fp.first[n_stages_total] = first;
        FRL_CUDA(set_smem_attr(tc_fused_kernel, device, di.max_smem));
        tc_fused_kernel<<<first, kLayerThreads, (size_t)di.max_smem, st>>>(fp);
        FRL_CUDA(cudaGetLastError());
        if ((rc = dump_trace(0, n_stages_total, fp.first, "w")) != FRL_OK) return rc;
        if (trace_dev) FRL_CUDA(cudaFree(trace_dev));
    } else if (chains) {
        auto launch_chain = [&](int s0, int s1) -> int {   // layer stages [s0, s1) as one dataflow kernel on all SMs
            FusedParams fp{};
            fp.n_layer_stages = s1 - s0; fp.n_jobs = 0; fp.n_tiles = nt;
            int first = 0;
            for (int i = s0; i < s1; ++i) {
                fp.layer[i - s0] = LP[i];
                fp.epi[i - s0] = epi_of[i];
                fp.first[i - s0] = first;
                first += pool[i];
            }
            fp.first[s1 - s0] = first;
            FRL_CUDA(set_smem_attr(tc_fused_kernel, device, di.max_smem));
            tc_fused_kernel<<<first, kLayerThreads, (size_t)di.max_smem, st>>>(fp);
            FRL_CUDA(cudaGetLastError());
            return dump_trace(s0, s1, fp.first, s0 == 0 ? "w" : "a");
        };
        if ((rc = launch_chain(0, depth + 1)) != FRL_OK) return rc;                  // forward chain + heads / loss
        if ((rc = launch_chain(depth + 1, n_layer_stages)) != FRL_OK) return rc;     // dgrad chain
        WgradParams wp{};
        wp.n_jobs = n_jobs; wp.n_tiles = nt;
        int first = 0;
        for (int l = 0; l < n_jobs; ++l) { wp.job[l] = WJ[l]; wp.first[l] = first; first += n_splits[l]; }
        wp.first[n_jobs] = first;
        FRL_CUDA(set_smem_attr(tc_wgrad_kernel, device, (int)kWgSmemBytes));
        tc_wgrad_kernel<<<first, kWgradThreads, kWgSmemBytes, st>>>(wp);
        FRL_CUDA(cudaGetLastError());
        if (trace_dev) FRL_CUDA(cudaFree(trace_dev));
    } else {
        auto launch = [&](int idx) -> int {
            const LayerParams& p = LP[idx];
            // > half of the SM's shared memory, so exactly one CTA (and its 512 TMEM columns) lives on an SM
            const size_t base = 2048 + ((size_t)p.K * p.N * 2 + 1023) / 1024 * 1024;
            const size_t smem_bytes = std::max(base + (size_t)p.n_stages * kChunkBytes, (size_t)120 * 1024);
            const int grid = std::min(p.n_tiles, sms);
            // every layer kernel after the first follows a <= one-CTA-per-SM persistent kernel: its CTAs may move in as
            // those leave (the first one follows the many small blocks of the prepare kernel and is launched normally)
            const bool pdl = use_pdl && idx >= 1;
            if (epi_of[idx] == EPI_RELU) {
                FRL_CUDA(set_smem_attr(tc_layer_kernel<EPI_RELU>, device, (int)smem_bytes));
                FRL_CUDA(launch_pdl(tc_layer_kernel<EPI_RELU>, dim3(grid), dim3(kLayerThreads), smem_bytes, st, pdl, p));
            } else if (epi_of[idx] == EPI_MASK) {
                FRL_CUDA(set_smem_attr(tc_layer_kernel<EPI_MASK>, device, (int)smem_bytes));
                FRL_CUDA(launch_pdl(tc_layer_kernel<EPI_MASK>, dim3(grid), dim3(kLayerThreads), smem_bytes, st, pdl, p));
            } else {
                FRL_CUDA(set_smem_attr(tc_layer_kernel<EPI_LOSS>, device, (int)smem_bytes));
                FRL_CUDA(launch_pdl(tc_layer_kernel<EPI_LOSS>, dim3(grid), dim3(kLayerThreads), smem_bytes, st, pdl, p));
            }
            FRL_CUDA(cudaGetLastError());
            return FRL_OK;
        };
        for (int idx = 0; idx <= depth; ++idx)
            if ((rc = launch(idx)) != FRL_OK) return rc;
        if (!grad) {
            tc_reduce_kernel<<<1, 256, 0, st>>>(rp, small, nullptr, 0);   // the losses only
            FRL_CUDA(cudaGetLastError());
            return FRL_OK;
        }
        for (int idx = depth + 1; idx < n_layer_stages; ++idx)
            if ((rc = launch(idx)) != FRL_OK) return rc;
        WgradParams wp{};
        wp.n_jobs = n_jobs; wp.n_tiles = nt;
        int first = 0;
        for (int l = 0; l < n_jobs; ++l) { wp.job[l] = WJ[l]; wp.first[l] = first; first += n_splits[l]; }
        wp.first[n_jobs] = first;
        const size_t smem_bytes = kWgSmemBytes;
        FRL_CUDA(set_smem_attr(tc_wgrad_kernel, device, (int)smem_bytes));
        FRL_CUDA(launch_pdl(tc_wgrad_kernel, dim3(first), dim3(kWgradThreads), smem_bytes, st, use_pdl, wp));
        FRL_CUDA(cudaGetLastError());
        pdl_reduce = use_pdl;
    }
    const int main_blocks = (int)((net->n_params + 255) / 256);
    FRL_CUDA(launch_pdl(tc_reduce_kernel, dim3(main_blocks + 1 + net->cfg.n_heads * A), dim3(256), 0, st, pdl_reduce, rp, small, grad,
                        main_blocks));
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

// number of global waits of the fused training kernel that gave up (0 in a healthy run); synchronises the device
long long train_timeouts(const frl_net* net) {
    if (!net->train_flags) return 0;
    unsigned v = 0;
    if (cudaSetDevice(net->cfg.device) != cudaSuccess) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(&v, net->train_flags, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)v;
}

}  // namespace frl
