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
#include <cstring>
#include <vector>

#include "frl_internal.h"

struct frl_comm {
    int rank = 0, world = 1, device = 0;
    long long max_floats = 0;
    uint8_t* local = nullptr;                 // exchange buffer of this rank
    std::vector<uint8_t*> peer;               // mapped exchange buffers, [world] (peer[rank] == local)
    std::vector<bool> opened;                 // mapped with cudaIpcOpenMemHandle (to be closed)
    unsigned long long* state = nullptr;      // device: [0] epoch, [1] publish block counter, [2] reduce block counter, [3] time-outs
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

__host__ __device__ inline size_t comm_data_offset(int world, long long max_floats, int slot) {
    const size_t flags = ((size_t)world * kFlagStride + 255) / 256 * 256;
    return flags + (size_t)slot * (((size_t)max_floats * 4 + 255) / 256 * 256);
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
    const size_t bytes = comm_data_offset(world, max_floats, 2);
    cudaError_t e = cudaMalloc(&c->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->state, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->state, 0, 4 * sizeof(unsigned long long));
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
    return comm_launch<3>(c, bufs, ns, n_bufs, stream);   // one launch: 128 blocks, all resident, so the in-kernel wait cannot deadlock
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
