// One-shot all-reduce (sum) of the data-parallel gradient over NVLink peer memory, as an ordinary kernel on the caller's
// stream: graph-capturable, no host synchronisation, no NCCL launch between the gradient kernels and clip + Adam.
// Replaces what torch.distributed / DistributedDataParallel would do for the reference's update
// (modril/flowril/flowril.py:311-330 under data parallelism).
//
// Every rank owns an exchange buffer  [flags: one 128-byte line per peer][data slot 0][data slot 1]  allocated with
// cudaMalloc and exported with cudaIpcGetMemHandle; frl_comm_connect maps the buffers of all peers (NVSwitch: every peer
// at full bandwidth).  One all-reduce of epoch e:
//   publish  copy the local vector(s) into the local data slot e & 1, __threadfence_system, and -- by the last block to
//            finish -- store e into flag[my rank] of EVERY peer's buffer;
//   reduce   every block waits until the local flags of all ranks read >= e (written remotely by the peers), then sums
//            its slice over the ranks IN RANK ORDER from the peers' data slots (so every rank computes bit-identical
//            sums) and writes the result back into the caller's vector(s).
// Two data slots make a second barrier unnecessary: a rank can be at most one epoch ahead of a peer (epoch e + 1 needs
// every peer's flag e + 1, which a peer raises only after its reduce of epoch e has been enqueued behind its publish
// e + 1 ... i.e. after it finished reading slot e & 1), so slot e & 1 is never overwritten while someone reads it.
// The vector is < 1 MB (the net has 0.2 M parameters): 8 ranks read 7 x 0.8 MB each, ~10 us on NVLink 5.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "frl_internal.h"

struct frl_comm {
    int rank = 0, world = 1, device = 0;
    long long max_floats = 0;
    uint8_t* local = nullptr;                 // exchange buffer of this rank
    std::vector<uint8_t*> peer;               // mapped exchange buffers, [world] (peer[rank] == local)
    std::vector<bool> opened;                 // mapped with cudaIpcOpenMemHandle (to be closed)
    unsigned long long* state = nullptr;      // device: [0] epoch, [1] publish block counter, [2] reduce block counter, [3] time-outs, [4] scatter block counter
    uint8_t** peer_dev = nullptr;             // device copy of `peer`
};

namespace frl {

constexpr int kFlagStride = 128;     // bytes: one line per rank
constexpr int kCommBlocks = 128;
constexpr int kCommThreads = 256;
constexpr int kMaxBufs = 8;

struct CommBufs {
    float* ptr[kMaxBufs];
    long long len[kMaxBufs];
    long long first[kMaxBufs + 1];   // position of buffer k in the data slot: prefix sums of the lengths rounded up to 4 floats
    int n;
};

__device__ __forceinline__ float4 ld_peer_v4(const float* p) {   // volatile: never served from a stale L1 line
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// exchange buffer: [flag set A: one line per rank][flag set B][vector slot 0][vector slot 1][result slot 0][result slot 1]
__host__ __device__ inline size_t comm_data_offset(int world, long long max_floats, int slot) {
    const size_t flags = ((size_t)2 * world * kFlagStride + 255) / 256 * 256;
    return flags + (size_t)slot * (((size_t)max_floats * 4 + 255) / 256 * 256);
}
constexpr int kCommAreas = 4;
// LL ("low latency") region behind the areas: every 16-byte cell holds two {fp32 value, 32-bit epoch} pairs, written with ONE
// 16-byte store whose 8-byte halves are atomic, so a value is valid exactly when its flag reads the current epoch: no
// fence and no separate flag round trip.   [recv1: world x per cells-of-4-floats (32 B)] [recv2: quads x 32 B]
__host__ __device__ inline long long comm_ll_per(int world, long long quads) { return (quads + world - 1) / world; }
__host__ __device__ inline size_t comm_ll_bytes(int world, long long max_floats) {
    const long long quads = (max_floats + 3) / 4 + 8;
    return (size_t)(world * (comm_ll_per(world, quads) + 1) + quads) * 32 + 512;
}

// Element i of the packed vector (all buffers back to back, each padded to a multiple of 4 floats) <-> buffer k, offset j.
// The buffer loop is unrolled with compile-time indices: CommBufs stays in the constant bank (a dynamic index into kernel
// parameters makes every thread copy the struct to local memory first).
template <typename F>
__device__ __forceinline__ void for_each_quad(const CommBufs& b, F&& body) {
#pragma unroll
    for (int k = 0; k < kMaxBufs; ++k) {
        if (k < b.n) {
            const long long quads = (b.len[k] + 3) / 4;
            const bool vec = (reinterpret_cast<uintptr_t>(b.ptr[k]) & 15) == 0;
            for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x)
                body(b.ptr[k] + 4 * q, b.first[k] / 4 + q, (int)min(4ll, b.len[k] - 4 * q), vec);
        }
    }
}

// PHASES: 1 = publish, 2 = wait + reduce, 3 = both in one launch (every block of the grid is resident: grid <= SMs)
template <int PHASES>
__global__ void __launch_bounds__(kCommThreads) comm_kernel(const __grid_constant__ CommBufs b, uint8_t* const* __restrict__ peer, int rank,
                                                            int world, long long max_floats, unsigned long long* state,
                                                            long long spin_limit) {
    const unsigned long long epoch = state[0] + 1;
    const size_t off = comm_data_offset(world, max_floats, (int)(epoch & 1));
    if (PHASES & 1) {
        float4* slot = reinterpret_cast<float4*>(peer[rank] + off);
        for_each_quad(b, [&](float* src, long long q, int valid, bool vec) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec && valid == 4) {
                v = *reinterpret_cast<const float4*>(src);
            } else {
                if (valid > 0) v.x = src[0];
                if (valid > 1) v.y = src[1];
                if (valid > 2) v.z = src[2];
                if (valid > 3) v.w = src[3];
            }
            slot[q] = v;
        });
        // the block's part of the slot is visible to the peers before the flag below: ONE system-scope fence per block
        // (bar.sync orders the other threads' stores before thread 0's fence; fences are cumulative)
        __syncthreads();
        __shared__ bool last;
        if (threadIdx.x == 0) {
            __threadfence_system();
            last = atomicAdd(&state[1], 1ull) == gridDim.x - 1;
        }
        __syncthreads();
        if (last) {               // the whole vector of this rank is in its slot: raise my flag in every peer's buffer
            __threadfence_system();
            for (int r = threadIdx.x; r < world; r += blockDim.x)
                *reinterpret_cast<volatile unsigned long long*>(peer[r] + (size_t)rank * kFlagStride) = epoch;
            if (threadIdx.x == 0) state[1] = 0;
        }
    }
    if (PHASES & 2) {
        if (threadIdx.x < world) {
            volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(peer[rank] + (size_t)threadIdx.x * kFlagStride);
            const long long t0 = clock64();
            while (*flag < epoch) {
                if (clock64() - t0 > spin_limit) {   // a peer never arrived (dead process): give up instead of hanging the GPU
                    atomicAdd(&state[3], 1ull);
                    break;
                }
            }
        }
        if (threadIdx.x < world) __threadfence_system();   // acquire side of the flags read above
        __syncthreads();
        for_each_quad(b, [&](float* dst, long long q, int valid, bool vec) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r0 = 0; r0 < world; r0 += 8) {   // up to eight peer loads in flight; summed in rank order (bit-identical on every rank)
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < world) v[u] = ld_peer_v4(reinterpret_cast<const float*>(peer[r0 + u] + off) + 4 * q);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < world) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
            }
            if (vec && valid == 4) {
                *reinterpret_cast<float4*>(dst) = acc;
            } else {
                if (valid > 0) dst[0] = acc.x;
                if (valid > 1) dst[1] = acc.y;
                if (valid > 2) dst[2] = acc.z;
                if (valid > 3) dst[3] = acc.w;
            }
        });
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&state[2], 1ull) == gridDim.x - 1) {
            state[2] = 0;
            state[0] = epoch;   // the next all-reduce uses the other slot
        }
    }
}

__device__ __forceinline__ void st_peer_v4(float* p, float4 v) {
    asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void wait_flags(uint8_t* base, int set, int world, unsigned long long epoch, long long spin_limit,
                                           unsigned long long* state) {
    if (threadIdx.x < world) {
        volatile unsigned long long* flag =
            reinterpret_cast<volatile unsigned long long*>(base + ((size_t)set * world + threadIdx.x) * kFlagStride);
        const long long t0 = clock64();
        while (*flag < epoch) {
            if (clock64() - t0 > spin_limit) {   // a peer never arrived (dead process): give up instead of hanging the GPU
                atomicAdd(&state[3], 1ull);
                break;
            }
        }
        __threadfence_system();   // acquire side of the flags
    }
    __syncthreads();
}
// every block has finished its part (bar.sync + ONE system fence per block): the last one raises this rank's flag of the
// given set in every peer's buffer
__device__ __forceinline__ void raise_flags(uint8_t* const* peer, int set, int rank, int world, unsigned long long epoch,
                                            unsigned long long* counter) {
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(counter, 1ull) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        for (int r = threadIdx.x; r < world; r += blockDim.x)
            *reinterpret_cast<volatile unsigned long long*>(peer[r] + ((size_t)set * world + rank) * kFlagStride) = epoch;
        if (threadIdx.x == 0) *counter = 0;
    }
}

// Two-shot all-reduce (world >= 3): remote READS are round trips, remote WRITES are posted, so instead of every rank
// reading the whole vector of every peer (one-shot: (W-1) x n floats of reads per rank) each rank reduces ONE slice
// (n / W floats from each peer, summed in rank order) and WRITES the reduced slice into every peer's result area:
// (W-1) x n / W floats read + as many written per rank, two flag rounds.  Every slice is reduced by exactly one rank, so
// all ranks end up with bit-identical vectors.
//   A  copy the local vector(s) into the local vector slot; flag set A
//   B  wait for everyone's A; reduce my slice from the peers' vector slots; scatter it into all result slots; flag set B
//   C  wait for everyone's B; copy the local result slot back into the caller's vector(s)
__global__ void __launch_bounds__(kCommThreads) comm_two_shot_kernel(const __grid_constant__ CommBufs b, uint8_t* const* __restrict__ peer,
                                                                     int rank, int world, long long max_floats,
                                                                     unsigned long long* state, long long spin_limit) {
    const unsigned long long epoch = state[0] + 1;
    const int par = (int)(epoch & 1);
    const size_t off_vec = comm_data_offset(world, max_floats, par), off_res = comm_data_offset(world, max_floats, 2 + par);
    {
        float4* slot = reinterpret_cast<float4*>(peer[rank] + off_vec);
        for_each_quad(b, [&](float* src, long long q, int valid, bool vec) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec && valid == 4) {
                v = *reinterpret_cast<const float4*>(src);
            } else {
                if (valid > 0) v.x = src[0];
                if (valid > 1) v.y = src[1];
                if (valid > 2) v.z = src[2];
                if (valid > 3) v.w = src[3];
            }
            slot[q] = v;
        });
    }
    raise_flags(peer, 0, rank, world, epoch, &state[1]);
    wait_flags(peer[rank], 0, world, epoch, spin_limit, state);
    {
        const long long quads = b.first[b.n] / 4;
        const long long q0 = quads * rank / world, q1 = quads * (rank + 1) / world;
        for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; q < q1; q += (long long)gridDim.x * blockDim.x) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r0 = 0; r0 < world; r0 += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < world) v[u] = ld_peer_v4(reinterpret_cast<const float*>(peer[r0 + u] + off_vec) + 4 * q);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < world) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
            }
            for (int r = 0; r < world; ++r) st_peer_v4(reinterpret_cast<float*>(peer[r] + off_res) + 4 * q, acc);
        }
    }
    raise_flags(peer, 1, rank, world, epoch, &state[4]);
    wait_flags(peer[rank], 1, world, epoch, spin_limit, state);
    {
        const float* res = reinterpret_cast<const float*>(peer[rank] + off_res);
        for_each_quad(b, [&](float* dst, long long q, int valid, bool vec) {
            const float4 acc = ld_peer_v4(res + 4 * q);
            if (vec && valid == 4) {
                *reinterpret_cast<float4*>(dst) = acc;
            } else {
                if (valid > 0) dst[0] = acc.x;
                if (valid > 1) dst[1] = acc.y;
                if (valid > 2) dst[2] = acc.z;
                if (valid > 3) dst[3] = acc.w;
            }
        });
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&state[2], 1ull) == gridDim.x - 1) {
        state[2] = 0;
        state[0] = epoch;
    }
}

struct LLQuad { uint4 a, b; };   // {v0, flag, v1, flag} {v2, flag, v3, flag}
__device__ __forceinline__ void st_ll(uint8_t* cell, float4 v, unsigned flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"(__float_as_uint(v.x)), "r"(flag),
                 "r"(__float_as_uint(v.y)), "r"(flag) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell + 16), "r"(__float_as_uint(v.z)), "r"(flag),
                 "r"(__float_as_uint(v.w)), "r"(flag) : "memory");
}
// spin until all four flags of the cell read `flag`; false after spin_limit cycles (a peer never arrived)
__device__ __forceinline__ bool ld_ll(const uint8_t* cell, unsigned flag, long long spin_limit, float4* out) {
    uint4 a, b;
    long long t0 = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(cell) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(cell + 16) : "memory");
        if (a.y == flag && a.w == flag && b.y == flag && b.w == flag) break;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > spin_limit) return false;
    }
    *out = make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
    return true;
}
__device__ __forceinline__ float4 ld_quad(const float* src, int valid, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && valid == 4) return *reinterpret_cast<const float4*>(src);
    if (valid > 0) v.x = src[0];
    if (valid > 1) v.y = src[1];
    if (valid > 2) v.z = src[2];
    if (valid > 3) v.w = src[3];
    return v;
}
__device__ __forceinline__ void st_quad(float* dst, float4 v, int valid, bool vec) {
    if (vec && valid == 4) { *reinterpret_cast<float4*>(dst) = v; return; }
    if (valid > 0) dst[0] = v.x;
    if (valid > 1) dst[1] = v.y;
    if (valid > 2) dst[2] = v.z;
    if (valid > 3) dst[3] = v.w;
}

// The default all-reduce: two-shot over LL cells, no fences, no barriers.  Quad q of the packed vector is OWNED by rank
// q / per.  Every thread walks the same quads in the same order on every rank, in three passes:
//   1  quads owned by a peer: write my values as LL cells into the owner's recv1[my rank]           (posted NVLink writes)
//   2  quads I own: sum my own values and the peers' cells (polled in LOCAL memory) in rank order, store the result in
//      the caller's vector and as LL cells into every peer's recv2
//   3  quads owned by a peer: poll my recv2 cell and store the result in the caller's vector
// Pass 1 never waits, pass 2 waits for pass-1 writes only, pass 3 for pass-2 writes only: no deadlock.  A quad is read
// (pass 1) and overwritten (pass 3) by the same thread, in program order.  Each slice is reduced by one rank, so all
// ranks hold bit-identical results.  A cell is rewritten by the next all-reduce only after its reader has consumed it
// (the writer of all-reduce e + 1 has received every reduced slice of e, which the peers send after reading recv1; and
// phase-2 cells of e + 1 are written after the peers' phase-1 cells of e + 1 arrived, i.e. after they finished e).
__global__ void __launch_bounds__(kCommThreads) comm_ll_kernel(const __grid_constant__ CommBufs b, uint8_t* const* __restrict__ peer, int rank,
                                                               int world, long long max_floats, unsigned long long* state,
                                                               long long spin_limit) {
    const unsigned flag = (unsigned)(state[0] + 1);
    const long long quads = b.first[b.n] / 4;
    const long long per = comm_ll_per(world, quads);
    const size_t off1 = comm_data_offset(world, max_floats, kCommAreas);
    const size_t off2 = off1 + (size_t)world * (comm_ll_per(world, (max_floats + 3) / 4 + 8) + 1) * 32;
    bool timed_out = false;
    for_each_quad(b, [&](float* src, long long q, int valid, bool vec) {
        const int owner = (int)(q / per);
        if (owner != rank) st_ll(peer[owner] + off1 + ((size_t)rank * per + (size_t)(q - owner * per)) * 32, ld_quad(src, valid, vec), flag);
    });
    for_each_quad(b, [&](float* src, long long q, int valid, bool vec) {
        if ((int)(q / per) != rank) return;
        const float4 mine = ld_quad(src, valid, vec);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < world; r0 += 8) {   // up to eight peers' cells polled together (their loads overlap), summed in rank order
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u;
                v[u] = mine;
                if (r < world && r != rank) {   // first probe of every cell without waiting on the others
                    const uint8_t* cell = peer[rank] + off1 + ((size_t)r * per + (size_t)(q - rank * per)) * 32;
                    uint4 a, b;
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(cell) : "memory");
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(cell + 16) : "memory");
                    if (a.y == flag && a.w == flag && b.y == flag && b.w == flag)
                        v[u] = make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
                    else
                        v[u].x = __int_as_float(0x7fc00001);   // marker: not there yet (re-polled below; a real value is overwritten anyway)
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u;
                if (r >= world) continue;
                if (r != rank && __float_as_int(v[u].x) == 0x7fc00001) {
                    // (a genuine payload with exactly this NaN bit pattern is simply re-read once: ld_ll returns it as it is)
                    if (!ld_ll(peer[rank] + off1 + ((size_t)r * per + (size_t)(q - rank * per)) * 32, flag, spin_limit, &v[u])) timed_out = true;
                }
                acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
        }
        st_quad(src, acc, valid, vec);
        for (int r = 0; r < world; ++r)
            if (r != rank) st_ll(peer[r] + off2 + (size_t)q * 32, acc, flag);
    });
    for_each_quad(b, [&](float* dst, long long q, int valid, bool vec) {
        if ((int)(q / per) == rank) return;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!ld_ll(peer[rank] + off2 + (size_t)q * 32, flag, spin_limit, &v)) timed_out = true;
        st_quad(dst, v, valid, vec);
    });
    if (timed_out) atomicAdd(&state[3], 1ull);
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&state[2], 1ull) == gridDim.x - 1) {
        state[2] = 0;
        state[0] = state[0] + 1;
    }
}

static int make_bufs(const frl_comm* c, float* const* bufs, const long long* ns, int n_bufs, CommBufs* out) {
    FRL_REQUIRE(c != nullptr, "null comm");
    FRL_REQUIRE(n_bufs >= 1 && n_bufs <= kMaxBufs, "1..8 buffers");
    CommBufs b{};
    long long total = 0;
    for (int i = 0; i < n_bufs; ++i) {
        FRL_REQUIRE(bufs[i] != nullptr && ns[i] >= 0, "null buffer");
        b.ptr[i] = bufs[i];
        b.len[i] = ns[i];
        b.first[i] = total;
        total += (ns[i] + 3) / 4 * 4;
    }
    b.first[n_bufs] = total;
    b.n = n_bufs;
    FRL_REQUIRE(total <= c->max_floats, "all-reduce larger than the exchange buffer");
    *out = b;
    return FRL_OK;
}

}  // namespace frl

using namespace frl;

extern "C" int frl_comm_create(int rank, int world, int device, long long max_floats, frl_comm** out) {
    FRL_REQUIRE(out != nullptr, "null out");
    FRL_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "bad rank / world");
    FRL_REQUIRE(max_floats >= 1, "max_floats >= 1");
    FRL_CUDA(cudaSetDevice(device));
    frl_comm* c = new (std::nothrow) frl_comm();
    if (!c) { set_error("out of host memory"); return FRL_ERR_NOMEM; }
    c->rank = rank; c->world = world; c->device = device; c->max_floats = max_floats;
    const size_t bytes = comm_data_offset(world, max_floats, kCommAreas) + comm_ll_bytes(world, max_floats);
    cudaError_t e = cudaMalloc(&c->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->state, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->state, 0, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&c->peer_dev, (size_t)world * sizeof(uint8_t*));
    if (e != cudaSuccess) {
        if (c->local) cudaFree(c->local);
        if (c->state) cudaFree(c->state);
        delete c;
        return cuda_fail(e, "frl_comm_create allocation", __FILE__, __LINE__);
    }
    c->peer.assign(world, nullptr);
    c->opened.assign(world, false);
    c->peer[rank] = c->local;
    FRL_CUDA(cudaDeviceSynchronize());
    *out = c;
    return FRL_OK;
}

extern "C" int frl_comm_handle(const frl_comm* c, void* out64) {
    FRL_REQUIRE(c && out64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    FRL_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    FRL_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(out64, &h, 64);
    return FRL_OK;
}

static int finish_connect(frl_comm* c) {
    for (int r = 0; r < c->world; ++r) FRL_REQUIRE(c->peer[r] != nullptr, "missing peer buffer");
    FRL_CUDA(cudaMemcpy(c->peer_dev, c->peer.data(), (size_t)c->world * sizeof(uint8_t*), cudaMemcpyHostToDevice));
    return FRL_OK;
}

extern "C" int frl_comm_connect(frl_comm* c, const void* handles) {
    FRL_REQUIRE(c && handles, "null argument");
    FRL_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * 64, 64);
        void* p = nullptr;
        FRL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<uint8_t*>(p);
        c->opened[r] = true;
    }
    return finish_connect(c);
}

extern "C" int frl_comm_connect_local(frl_comm* c, frl_comm* const* all) {
    FRL_REQUIRE(c && all, "null argument");
    FRL_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        FRL_REQUIRE(all[r] != nullptr && all[r]->world == c->world && all[r]->rank == r && all[r]->max_floats == c->max_floats,
                    "frl_comm_connect_local: rank r of `all` must be the comm of rank r with the same geometry");
        c->peer[r] = all[r]->local;
        if (all[r]->device != c->device) {   // same process: plain device pointers, peer access switched on once per device pair
            int can = 0;
            FRL_CUDA(cudaDeviceCanAccessPeer(&can, c->device, all[r]->device));
            FRL_REQUIRE(can != 0, "frl_comm_connect_local: no peer access between the devices");
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        }
    }
    return finish_connect(c);
}

template <int PHASES>
static int comm_launch(frl_comm* c, float* const* bufs, const long long* ns, int n_bufs, void* stream) {
    CommBufs b;
    int rc = make_bufs(c, bufs, ns, n_bufs, &b);
    if (rc != FRL_OK) return rc;
    FRL_CUDA(cudaSetDevice(c->device));
    const long long spin_limit = 20ll * 1000 * 1000 * 1000;   // ~10 s of SM clocks
    comm_kernel<PHASES><<<kCommBlocks, kCommThreads, 0, (cudaStream_t)stream>>>(b, c->peer_dev, c->rank, c->world, c->max_floats,
                                                                                c->state, spin_limit);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_comm_publish(frl_comm* c, float* const* bufs, const long long* ns, int n_bufs, void* stream) {
    return comm_launch<1>(c, bufs, ns, n_bufs, stream);
}

extern "C" int frl_comm_reduce(frl_comm* c, float* const* bufs, const long long* ns, int n_bufs, void* stream) {
    return comm_launch<2>(c, bufs, ns, n_bufs, stream);
}

extern "C" int frl_comm_allreduce(frl_comm* c, float* const* bufs, const long long* ns, int n_bufs, void* stream) {
    // one launch: 128 blocks, all resident, so the in-kernel waits cannot deadlock
    FRL_REQUIRE(c != nullptr, "null comm");
    // default: the fence-free LL two-shot kernel; FRL_COMM_ALGO = one_shot | two_shot selects the flag-based kernels (A/B tests)
    const char* algo = getenv("FRL_COMM_ALGO");
    const std::string a = algo ? algo : "";
    if (a == "one_shot") return comm_launch<3>(c, bufs, ns, n_bufs, stream);
    CommBufs b;
    int rc = make_bufs(c, bufs, ns, n_bufs, &b);
    if (rc != FRL_OK) return rc;
    FRL_CUDA(cudaSetDevice(c->device));
    const long long spin_limit = 20ll * 1000 * 1000 * 1000;
    if (a != "two_shot") {
        // no barrier inside: any grid size is safe; enough threads for one quad per thread and pass (capped at two CTAs per SM)
        const long long quads = b.first[b.n] / 4;
        const int blocks = (int)std::max<long long>(1, std::min<long long>((quads + kCommThreads - 1) / kCommThreads, 296));
        comm_ll_kernel<<<blocks, kCommThreads, 0, (cudaStream_t)stream>>>(b, c->peer_dev, c->rank, c->world, c->max_floats, c->state,
                                                                         spin_limit);
        FRL_CUDA(cudaGetLastError());
        return FRL_OK;
    }
    comm_two_shot_kernel<<<kCommBlocks, kCommThreads, 0, (cudaStream_t)stream>>>(b, c->peer_dev, c->rank, c->world, c->max_floats,
                                                                                 c->state, spin_limit);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" long long frl_comm_timeouts(frl_comm* c) {
    if (!c) return -1;
    unsigned long long v = 0;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, c->state + 3, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)v;
}

extern "C" int frl_comm_destroy(frl_comm* c) {
    if (!c) return FRL_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
    if (c->state) cudaFree(c->state);
    if (c->peer_dev) cudaFree(c->peer_dev);
    delete c;
    return FRL_OK;
}
