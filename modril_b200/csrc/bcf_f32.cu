// BCF flow-matching policy (scope row f3): VectorNetwork forward, the Euler action sampler and the behaviour-cloning
// flow-matching loss + gradient, FP32 on the SGEMM building block of sgemm.cuh.
// Replaces  VectorNetwork.forward  (deps/rl-toolkit/rlf/rl/model.py:422-432),  FlowPolicy.get_action
// (rlf/policies/flow_policy.py:49-63)  and the loss / backward of  BehaviorCloneFlowMatching._bc_step
// (rlf/algos/il/bcf.py:108-128).
//
//   s_feat = relu(Ws s + bs), t_feat = relu(Wt2 relu(wt1 t + bt1) + bt2), x_feat = relu(Wa x + ba)
//   v = W2 relu(W1 [s_feat | t_feat | x_feat] + b1) + b2
// The three embeddings are written straight into the column slices of one [R][3H] buffer (ldc = 3H), which is also the
// ReLU mask of the backward pass.  The loss runs the net on 2B stacked rows (rows 0..B-1: (x_t, t), rows B..2B-1:
// (a, 1)), so forward and backward are one pass each.  Batch reductions are split and reduced in a fixed order
// (deterministic).  The sampler is ONE kernel over all Euler steps (see vnet_sample_kernel).
#include <algorithm>
#include <cmath>
#include <new>

#include "frl_internal.h"
#include "sgemm.cuh"

struct frl_vnet {
    frl_vnet_config cfg;
    int Sp, Ap, Aop;                 // padded input / output widths (multiples of 16)
    long long off[12];               // flat offsets in state_dict order
    long long n_params;
    float* pack = nullptr;           // kernel-side layouts
    size_t o_wt1, o_bt1, o_Wt2t, o_Wt2, o_bt2, o_Wst, o_bs, o_Wat, o_ba, o_W1t, o_W1, o_b1, o_W2t, o_W2, o_b2;
    size_t pack_floats = 0;
    const float* params_dev = nullptr;
    unsigned long long version = 0, packed = 0;
    float* ws = nullptr;
    size_t ws_floats = 0;
    std::vector<float*> retired;
};

namespace frl {

struct VPackJob {
    long long src_off, dst_off;
    int rows, cols, dst_ld, transpose;
    long long first;
};
struct VPackTable {
    VPackJob job[20];
    int n_jobs;
    long long total;
};
__global__ void vnet_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, VPackTable tab) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tab.total; i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < tab.n_jobs && tab.job[j + 1].first <= i) ++j;
        const VPackJob& jb = tab.job[j];
        const long long e = i - jb.first;
        const int r = (int)(e / jb.cols), c = (int)(e % jb.cols);
        const float v = src[jb.src_off + e];
        if (jb.transpose) dst[jb.dst_off + (long long)c * jb.dst_ld + r] = v;
        else dst[jb.dst_off + (long long)r * jb.dst_ld + c] = v;
    }
}

static int vnet_ensure_pack(frl_vnet* n, cudaStream_t st) {
    if (n->packed == n->version) return FRL_OK;
    const int H = n->cfg.hidden, T = n->cfg.time_dim, S = n->cfg.s_dim, A = n->cfg.a_dim;
    FRL_CUDA(cudaMemsetAsync(n->pack, 0, n->pack_floats * sizeof(float), st));
    VPackTable tab{};
    long long total = 0;
    auto add = [&](int idx, size_t dst, int rows, int cols, int dst_ld, int transpose) {
        VPackJob& j = tab.job[tab.n_jobs++];
        j = VPackJob{n->off[idx], (long long)dst, rows, cols, dst_ld, transpose, total};
        total += (long long)rows * cols;
    };
    add(0, n->o_wt1, 1, T, T, 0);          // time_embed.0.weight [T,1] -> vector
    add(1, n->o_bt1, 1, T, T, 0);
    add(2, n->o_Wt2t, H, T, H, 1);         // [T][H]
    add(2, n->o_Wt2, H, T, T, 0);          // [H][T]
    add(3, n->o_bt2, 1, H, H, 0);
    add(4, n->o_Wst, H, S, H, 1);          // [Sp][H]
    add(5, n->o_bs, 1, H, H, 0);
    add(6, n->o_Wat, H, A, H, 1);          // [Ap][H]
    add(7, n->o_ba, 1, H, H, 0);
    add(8, n->o_W1t, H, 3 * H, H, 1);      // [3H][H]
    add(8, n->o_W1, H, 3 * H, 3 * H, 0);   // [H][3H]
    add(9, n->o_b1, 1, H, H, 0);
    add(10, n->o_W2t, A, H, n->Aop, 1);    // [H][Aop]
    add(10, n->o_W2, A, H, H, 0);          // [Aop][H]
    add(11, n->o_b2, 1, A, n->Aop, 0);
    tab.total = total;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 592);
    vnet_pack_kernel<<<blocks, 256, 0, st>>>(n->params_dev, n->pack, tab);
    FRL_CUDA(cudaGetLastError());
    n->packed = n->version;
    return FRL_OK;
}

static int vnet_ws(frl_vnet* n, size_t floats) {
    if (n->ws_floats >= floats) return FRL_OK;
    if (n->ws) n->retired.push_back(n->ws);   // kept alive until destroy (captured graphs / in-flight launches)
    n->ws = nullptr;
    n->ws_floats = 0;
    FRL_CUDA(cudaMalloc(&n->ws, floats * sizeof(float)));
    n->ws_floats = floats;
    return FRL_OK;
}

// padded inputs + first time-embedding layer.  Row r of the R stacked rows takes (x, s, t) from source row r % B:
//   mode 0: x = x_src, t = t_src                               (plain forward / sampler)
//   mode 1: rows < B: x = (1-t) x0 + t a, t;  rows >= B: x = a, t = 1     (bcf.py:110-121)
__global__ void vnet_prepare_kernel(const float* __restrict__ s, const float* __restrict__ x_src,
                                    const float* __restrict__ x0, const float* __restrict__ t_src, float t_const,
                                    const float* __restrict__ wt1, const float* __restrict__ bt1, float* __restrict__ Sp_in,
                                    float* __restrict__ Xp_in, float* __restrict__ T1, float* __restrict__ Tv, long long B,
                                    long long R, int S, int A, int Sp, int Ap, int T, int mode, int write_s) {
    const int W = Sp + Ap + T;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= R * W) return;
    const long long r = i / W, b = r % B;
    const int c = (int)(i % W);
    float t = t_src ? t_src[b] : t_const;
    if (mode == 1 && r >= B) t = 1.f;
    if (c < Sp) {
        if (write_s) Sp_in[r * Sp + c] = c < S ? s[b * S + c] : 0.f;
    } else if (c < Sp + Ap) {
        const int k = c - Sp;
        float v = 0.f;
        if (k < A) {
            if (mode == 0) v = x_src[b * A + k];
            else if (r < B) v = __fadd_rn(__fmul_rn(1.f - t, x0[b * A + k]), __fmul_rn(t, x_src[b * A + k]));
            else v = x_src[b * A + k];
        }
        Xp_in[r * Ap + k] = v;
    } else {
        const int k = c - Sp - Ap;
        const float z = fmaf(t, wt1[k], bt1[k]);
        T1[r * T + k] = z > 0.f ? z : 0.f;
        if (k == 0 && Tv) Tv[r] = t;
    }
}

// losses and dL/dV in place (bcf.py:112-128): rows < B flow term, rows >= B the t = 1 behaviour-cloning term
__global__ void __launch_bounds__(256) bcf_loss_kernel(float* __restrict__ V, int ldv, const float* __restrict__ a,
                                                       const float* __restrict__ x0, long long B, int A, float dscale,
                                                       float coef, double* __restrict__ partial) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double flow = 0.0, act = 0.0;
    if (r < 2 * B) {
        const long long b = r % B;
        float* v = V + r * ldv;
        float acc = 0.f;
        for (int j = 0; j < A; ++j) {
            const float d = r < B ? v[j] - (a[b * A + j] - x0[b * A + j]) : x0[b * A + j] + v[j] - a[b * A + j];
            acc += d * d;
            v[j] = (r < B ? dscale : dscale * coef) * d;
        }
        for (int j = A; j < ldv; ++j) v[j] = 0.f;
        if (r < B) flow = acc; else act = acc;
    }
    __shared__ double sh[2][8];
    for (int o = 16; o > 0; o >>= 1) {
        flow += __shfl_xor_sync(0xffffffffu, flow, o);
        act += __shfl_xor_sync(0xffffffffu, act, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = flow; sh[1][threadIdx.x >> 5] = act; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double f = 0.0, c = 0.0;
        for (int i = 0; i < 8; ++i) { f += sh[0][i]; c += sh[1][i]; }
        partial[2 * blockIdx.x] = f;
        partial[2 * blockIdx.x + 1] = c;
    }
}
__global__ void bcf_loss_final_kernel(const double* __restrict__ partial, int n, double inv, float coef, float* __restrict__ out) {
    __shared__ double sh[2][256];
    double f = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { f += partial[2 * i]; c += partial[2 * i + 1]; }
    sh[0][threadIdx.x] = f; sh[1][threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float flow = (float)(sh[0][0] * inv), act = (float)(sh[1][0] * inv);
        out[0] = flow + coef * act;   // bcf.py:128
        out[1] = flow;
        out[2] = act;
    }
}
// gradients of the first time-embedding layer (K = 1): dwt1[k] = sum_r dT1[r][k] t[r], dbt1[k] = sum_r dT1[r][k]
__global__ void time_grad_partial_kernel(const float* __restrict__ dT1, const float* __restrict__ Tv, long long R, int T,
                                         int rows_per_chunk, float* __restrict__ partial) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= T) return;
    const long long lo = (long long)blockIdx.y * rows_per_chunk, hi = min(R, lo + rows_per_chunk);
    float aw = 0.f, ab = 0.f;
    for (long long r = lo; r < hi; ++r) {
        const float d = dT1[r * T + col];
        aw = fmaf(d, Tv[r], aw);
        ab += d;
    }
    partial[((size_t)blockIdx.y * 2) * T + col] = aw;
    partial[((size_t)blockIdx.y * 2 + 1) * T + col] = ab;
}
__global__ void time_grad_reduce_kernel(const float* __restrict__ partial, int chunks, int T, float* __restrict__ gw,
                                        float* __restrict__ gb, int accumulate) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= T) return;
    float aw = 0.f, ab = 0.f;
    for (int c = 0; c < chunks; ++c) { aw += partial[((size_t)c * 2) * T + col]; ab += partial[((size_t)c * 2 + 1) * T + col]; }
    gw[col] = accumulate ? gw[col] + aw : aw;
    gb[col] = accumulate ? gb[col] + ab : ab;
}
__global__ void copy_out_kernel(const float* __restrict__ V, int ldv, float* __restrict__ out, long long B, int A) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B * A) return;
    out[i] = V[(i / A) * ldv + i % A];
}


// ---- Euler action sampler (flow_policy.py:49-63), fused over all steps ---------------------------------------------
// The head's first layer is linear in the concatenated features, so per row and step only the action-embedding block
// of W1 is recurrent:  h1 = relu(W1[:,0:H] s_feat  +  (W1[:,H:2H] t_feat_i + b1)  +  W1[:,2H:3H] x_feat).
// The first term is constant over the steps (computed once per row), the second is the same for every row (one small
// launch for all steps: step_bias_kernel), the third is a [H x H] mat-vec per row and step whose weights are streamed
// from L2.  One CTA integrates R rows through all steps; thread (g, j) owns column j over the k-slice g of KS.
__global__ void __launch_bounds__(1024) step_bias_kernel(const float* __restrict__ wt1, const float* __restrict__ bt1,
                                                         const float* __restrict__ Wt2t, const float* __restrict__ bt2,
                                                         const float* __restrict__ W1t, const float* __restrict__ b1,
                                                         float* __restrict__ tb, int T, int H, double dt) {
    extern __shared__ float sm_tb[];
    float* T1 = sm_tb;
    float* T2 = sm_tb + T;
    const int i = blockIdx.x;
    const float t = (float)(i * dt);   // python double i*dt -> fp32 tensor (flow_policy.py:58)
    for (int k = threadIdx.x; k < T; k += blockDim.x) T1[k] = fmaxf(fmaf(t, wt1[k], bt1[k]), 0.f);
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float acc = bt2[j];
        for (int k = 0; k < T; ++k) acc = fmaf(Wt2t[(size_t)k * H + j], T1[k], acc);
        T2[j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float acc = b1[j];
        for (int k = 0; k < H; ++k) acc = fmaf(W1t[(size_t)(H + k) * H + j], T2[k], acc);
        tb[(size_t)i * H + j] = acc;
    }
}

template <int R>
__device__ __forceinline__ void slice_matvec(const float* __restrict__ W, const float* __restrict__ feat, float* __restrict__ part,
                                             int H, int KS) {
    const int g = threadIdx.x / H, j = threadIdx.x % H;
    const int klen = H / KS, k0 = g * klen;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int k = k0; k < k0 + klen; k += 4) {
        const float w0 = __ldg(W + (size_t)k * H + j), w1 = __ldg(W + (size_t)(k + 1) * H + j);
        const float w2 = __ldg(W + (size_t)(k + 2) * H + j), w3 = __ldg(W + (size_t)(k + 3) * H + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 f = *reinterpret_cast<const float4*>(feat + r * H + k);
            acc[r] = fmaf(w0, f.x, acc[r]);
            acc[r] = fmaf(w1, f.y, acc[r]);
            acc[r] = fmaf(w2, f.z, acc[r]);
            acc[r] = fmaf(w3, f.w, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) part[(g * R + r) * H + j] = acc[r];
}

template <int R>
__global__ void __launch_bounds__(1024) vnet_sample_kernel(const float* __restrict__ s, const float* __restrict__ x0,
                                                           float* __restrict__ action, const float* __restrict__ tb,
                                                           const float* __restrict__ Wst, const float* __restrict__ bs,
                                                           const float* __restrict__ Wat, const float* __restrict__ ba,
                                                           const float* __restrict__ W1t, const float* __restrict__ W2,
                                                           const float* __restrict__ b2, long long B, int S, int A, int H,
                                                           int KS, int n_steps, float dt) {
    extern __shared__ __align__(16) float sm_s[];
    float* feat = sm_s;                    // [R][H]  s_feat, then x_feat / h1 per step
    float* cs = feat + R * H;              // [R][H]  W1[:,0:H] s_feat
    float* part = cs + R * H;              // [KS][R][H]
    float* xs = part + KS * R * H;         // [R][A]
    float* ss = xs + R * A;                // [R][S]
    const long long row0 = (long long)blockIdx.x * R;
    const int nt = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < R * S; i += nt) { const long long r = row0 + i / S; ss[i] = r < B ? s[r * S + i % S] : 0.f; }
    for (int i = tid; i < R * A; i += nt) { const long long r = row0 + i / A; xs[i] = r < B ? x0[r * A + i % A] : 0.f; }
    __syncthreads();
    for (int i = tid; i < R * H; i += nt) {
        const int r = i / H, j = i % H;
        float acc = bs[j];
        for (int k = 0; k < S; ++k) acc = fmaf(__ldg(Wst + (size_t)k * H + j), ss[r * S + k], acc);
        feat[i] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    slice_matvec<R>(W1t, feat, part, H, KS);
    __syncthreads();
    for (int i = tid; i < R * H; i += nt) {
        float acc = part[i];
        for (int g = 1; g < KS; ++g) acc += part[g * R * H + i];
        cs[i] = acc;
    }
    const int n_warps = nt >> 5, warp = tid >> 5, lane = tid & 31;
    const float* W1x = W1t + (size_t)2 * H * H;
    for (int step = 0; step < n_steps; ++step) {
        __syncthreads();   // xs of the previous step / cs complete; feat free
        for (int i = tid; i < R * H; i += nt) {
            const int r = i / H, j = i % H;
            float acc = ba[j];
            for (int k = 0; k < A; ++k) acc = fmaf(__ldg(Wat + (size_t)k * H + j), xs[r * A + k], acc);
            feat[i] = fmaxf(acc, 0.f);
        }
        __syncthreads();
        slice_matvec<R>(W1x, feat, part, H, KS);
        __syncthreads();
        const float* tbi = tb + (size_t)step * H;
        for (int i = tid; i < R * H; i += nt) {
            float acc = part[i];
            for (int g = 1; g < KS; ++g) acc += part[g * R * H + i];
            feat[i] = fmaxf(acc + (cs[i] + __ldg(tbi + i % H)), 0.f);
        }
        __syncthreads();
        if (warp < n_warps) {   // full warps only: (row, action component) dot products with a shuffle reduction
            for (int p = warp; p < R * A; p += n_warps) {
                const int r = p / A, a = p % A;
                float acc = 0.f;
                for (int j = lane; j < H; j += 32) acc = fmaf(__ldg(W2 + (size_t)a * H + j), feat[r * H + j], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) xs[p] = fmaf(acc + b2[a], dt, xs[p]);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < R * A; i += nt) { const long long r = row0 + i / A; if (r < B) action[r * A + i % A] = xs[i]; }
}

struct VBuf {
    float *Sp_in, *Xp_in, *T1, *Tv, *Hc, *H1, *V;
};

// embeddings + head for R rows whose padded inputs / T1 are already in place
static int vnet_trunk(frl_vnet* n, const VBuf& b, int R, bool do_state, cudaStream_t st) {
    const int H = n->cfg.hidden, T = n->cfg.time_dim;
    const float* pk = n->pack;
    int rc;
    if (do_state) {
        rc = run_gemm<false, EPI_BIAS_RELU>(b.Sp_in, n->Sp, pk + n->o_Wst, H, b.Hc, 3 * H, R, H, n->Sp, pk + n->o_bs, nullptr, 0, 1, n->Sp, st);
        if (rc != FRL_OK) return rc;
    }
    rc = run_gemm<false, EPI_BIAS_RELU>(b.T1, T, pk + n->o_Wt2t, H, b.Hc + H, 3 * H, R, H, T, pk + n->o_bt2, nullptr, 0, 1, T, st);
    if (rc != FRL_OK) return rc;
    rc = run_gemm<false, EPI_BIAS_RELU>(b.Xp_in, n->Ap, pk + n->o_Wat, H, b.Hc + 2 * H, 3 * H, R, H, n->Ap, pk + n->o_ba, nullptr, 0, 1, n->Ap, st);
    if (rc != FRL_OK) return rc;
    rc = run_gemm<false, EPI_BIAS_RELU>(b.Hc, 3 * H, pk + n->o_W1t, H, b.H1, H, R, H, 3 * H, pk + n->o_b1, nullptr, 0, 1, 3 * H, st);
    if (rc != FRL_OK) return rc;
    return run_gemm<false, EPI_BIAS>(b.H1, H, pk + n->o_W2t, n->Aop, b.V, n->Aop, R, n->Aop, H, pk + n->o_b2, nullptr, 0, 1, H, st);
}

static VBuf carve(frl_vnet* n, float* ws, size_t R, size_t* used) {
    const int H = n->cfg.hidden, T = n->cfg.time_dim;
    size_t off = 0;
    auto take = [&](size_t f) { size_t o = off; off += (f + 3) / 4 * 4; return ws ? ws + o : (float*)nullptr; };
    VBuf b;
    b.Sp_in = take(R * n->Sp); b.Xp_in = take(R * n->Ap); b.T1 = take(R * T); b.Tv = take(R);
    b.Hc = take(R * 3 * H); b.H1 = take(R * H); b.V = take(R * n->Aop);
    *used = off;
    return b;
}

}  // namespace frl

using namespace frl;

extern "C" int frl_vnet_create(const frl_vnet_config* cfg, frl_vnet** out) {
    FRL_REQUIRE(cfg && out, "null config/out");
    FRL_REQUIRE(cfg->s_dim >= 1 && cfg->s_dim <= 1024 && cfg->a_dim >= 1 && cfg->a_dim <= 64, "bad s_dim / a_dim");
    FRL_REQUIRE(cfg->hidden >= 16 && cfg->hidden % 16 == 0 && cfg->hidden <= 1024, "hidden must be a multiple of 16, <= 1024");
    FRL_REQUIRE(cfg->time_dim >= 16 && cfg->time_dim % 16 == 0 && cfg->time_dim <= 1024, "time_dim must be a multiple of 16");
    int ndev = 0;
    FRL_CUDA(cudaGetDeviceCount(&ndev));
    FRL_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
    FRL_CUDA(cudaSetDevice(cfg->device));
    frl_vnet* n = new (std::nothrow) frl_vnet();
    if (!n) { set_error("out of host memory"); return FRL_ERR_NOMEM; }
    n->cfg = *cfg;
    const int H = cfg->hidden, T = cfg->time_dim, S = cfg->s_dim, A = cfg->a_dim;
    n->Sp = round_up(S, 16); n->Ap = round_up(A, 16); n->Aop = round_up(A, 16);
    const long long sizes[12] = {T, T, (long long)H * T, H, (long long)H * S, H, (long long)H * A, H,
                                 (long long)H * 3 * H, H, (long long)A * H, A};
    long long off = 0;
    for (int i = 0; i < 12; ++i) { n->off[i] = off; off += sizes[i]; }
    n->n_params = off;
    size_t f = 0;
    auto take = [&](size_t x) { size_t o = f; f += (x + 3) / 4 * 4; return o; };
    n->o_wt1 = take(T); n->o_bt1 = take(T); n->o_Wt2t = take((size_t)T * H); n->o_Wt2 = take((size_t)H * T); n->o_bt2 = take(H);
    n->o_Wst = take((size_t)n->Sp * H); n->o_bs = take(H); n->o_Wat = take((size_t)n->Ap * H); n->o_ba = take(H);
    n->o_W1t = take((size_t)3 * H * H); n->o_W1 = take((size_t)H * 3 * H); n->o_b1 = take(H);
    n->o_W2t = take((size_t)H * n->Aop); n->o_W2 = take((size_t)n->Aop * H); n->o_b2 = take(n->Aop);
    n->pack_floats = f;
    cudaError_t e = cudaMalloc(&n->pack, f * sizeof(float));
    if (e != cudaSuccess) { delete n; return cuda_fail(e, "cudaMalloc(vnet pack)", __FILE__, __LINE__); }
    *out = n;
    return FRL_OK;
}

extern "C" int frl_vnet_destroy(frl_vnet* n) {
    if (!n) return FRL_OK;
    cudaSetDevice(n->cfg.device);
    if (n->pack) cudaFree(n->pack);
    if (n->ws) cudaFree(n->ws);
    for (float* r : n->retired) cudaFree(r);
    delete n;
    return FRL_OK;
}

extern "C" long long frl_vnet_num_params(const frl_vnet* n) { return n ? n->n_params : -1; }

extern "C" int frl_vnet_set_params(frl_vnet* n, const float* params_dev, void* /*stream*/) {
    FRL_REQUIRE(n && params_dev, "null net/params");
    n->params_dev = params_dev;
    n->version += 1;
    return FRL_OK;
}

extern "C" int frl_vnet_forward(frl_vnet* n, const float* x, const float* s, const float* t, float* v, long long B, void* stream) {
    FRL_REQUIRE(n && n->version > 0, "frl_vnet_set_params has not been called");
    if (B <= 0) return FRL_OK;
    FRL_REQUIRE(x && s && t && v, "null pointer");
    FRL_REQUIRE(B < (1ll << 31) / (3 * n->cfg.hidden), "batch too large");
    FRL_CUDA(cudaSetDevice(n->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = vnet_ensure_pack(n, st);
    if (rc != FRL_OK) return rc;
    size_t used = 0;
    carve(n, nullptr, (size_t)B, &used);
    rc = vnet_ws(n, used);
    if (rc != FRL_OK) return rc;
    VBuf b = carve(n, n->ws, (size_t)B, &used);
    const int S = n->cfg.s_dim, A = n->cfg.a_dim, T = n->cfg.time_dim;
    const long long W = n->Sp + n->Ap + T;
    vnet_prepare_kernel<<<(unsigned)((B * W + 255) / 256), 256, 0, st>>>(s, x, nullptr, t, 0.f, n->pack + n->o_wt1, n->pack + n->o_bt1,
                                                                       b.Sp_in, b.Xp_in, b.T1, nullptr, B, B, S, A, n->Sp, n->Ap, T, 0, 1);
    FRL_CUDA(cudaGetLastError());
    rc = vnet_trunk(n, b, (int)B, true, st);
    if (rc != FRL_OK) return rc;
    copy_out_kernel<<<(unsigned)((B * A + 255) / 256), 256, 0, st>>>(b.V, n->Aop, v, B, A);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_vnet_sample(frl_vnet* n, const float* s, const float* x0, int n_steps, float* action, long long B, void* stream) {
    FRL_REQUIRE(n && n->version > 0, "frl_vnet_set_params has not been called");
    if (B <= 0) return FRL_OK;
    FRL_REQUIRE(s && x0 && action, "null pointer");
    FRL_REQUIRE(n_steps >= 1 && n_steps <= 4096, "n_steps must be in 1..4096");
    FRL_CUDA(cudaSetDevice(n->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = vnet_ensure_pack(n, st);
    if (rc != FRL_OK) return rc;
    const int S = n->cfg.s_dim, A = n->cfg.a_dim, T = n->cfg.time_dim, H = n->cfg.hidden;
    rc = vnet_ws(n, (size_t)n_steps * H);
    if (rc != FRL_OK) return rc;
    float* tb = n->ws;
    const float* pk = n->pack;
    const double dt = 1.0 / n_steps;   // python double, flow_policy.py:55
    step_bias_kernel<<<n_steps, std::min(1024, round_up(std::max(T, H), 32)), (size_t)(T + H) * sizeof(float), st>>>(
        pk + n->o_wt1, pk + n->o_bt1, pk + n->o_Wt2t, pk + n->o_bt2, pk + n->o_W1t, pk + n->o_b1, tb, T, H, dt);
    FRL_CUDA(cudaGetLastError());
    int KS = 4;
    while (KS > 1 && H * KS > 1024) KS >>= 1;
    auto smem_for = [&](int R) { return (size_t)(R * H * (2 + KS) + R * (A + S)) * sizeof(float); };
    // 16 rows per CTA once that still fills the SMs (weights are streamed from L2 once per CTA and step)
    const bool wide = B >= 16ll * 296 && smem_for(16) <= 200 * 1024;
    const int R = wide ? 16 : 4;
    const size_t smem = smem_for(R);
    FRL_REQUIRE(smem <= 227 * 1024, "state too wide for the fused sampler");
    const unsigned grid = (unsigned)((B + R - 1) / R);
    auto kern = wide ? vnet_sample_kernel<16> : vnet_sample_kernel<4>;
    static bool attr_done[2][64] = {};
    if (!attr_done[wide][n->cfg.device & 63]) {
        FRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done[wide][n->cfg.device & 63] = true;
    }
    kern<<<grid, H * KS, smem, st>>>(s, x0, action, tb, pk + n->o_Wst, pk + n->o_bs, pk + n->o_Wat, pk + n->o_ba, pk + n->o_W1t,
                                     pk + n->o_W2, pk + n->o_b2, B, S, A, H, KS, n_steps, (float)dt);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_bcf_loss_grad(frl_vnet* n, const float* s, const float* actions, const float* x0, const float* t, long long B,
                                 long long mean_divisor, float bc_coef, float grad_scale, int accumulate, float* losses,
                                 float* grad, void* stream) {
    FRL_REQUIRE(n && n->version > 0, "frl_vnet_set_params has not been called");
    FRL_REQUIRE(s && actions && x0 && t, "null input");
    FRL_REQUIRE(B > 0 && 2 * B < (1ll << 31) / (3 * n->cfg.hidden), "batch size out of range");
    FRL_CUDA(cudaSetDevice(n->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = vnet_ensure_pack(n, st);
    if (rc != FRL_OK) return rc;
    const int H = n->cfg.hidden, T = n->cfg.time_dim, S = n->cfg.s_dim, A = n->cfg.a_dim, Aop = n->Aop;
    const int R = (int)(2 * B);
    const double div = (double)(mean_divisor > 0 ? mean_divisor : B) * A;   // F.mse_loss: mean over B*A elements

    auto plan_splits = [&](int tiles) {
        int want = std::max(1, 296 / std::max(1, tiles));
        int maxs = std::max(1, (R + 63) / 64);
        return std::min(want, maxs);
    };
    const int sp_big = plan_splits(((H + 127) / 128) * ((3 * H + 127) / 128));
    const int sp_hh = plan_splits(((H + 127) / 128) * ((H + 127) / 128));
    const int sp_small = plan_splits(2);
    const int max_sp = std::max(sp_big, std::max(sp_hh, sp_small));
    const int cs_chunks = std::max(1, std::min(148, (R + 255) / 256));
    const int cs_rows = (R + cs_chunks - 1) / cs_chunks;
    const int loss_blocks = (R + 255) / 256;

    size_t used = 0;
    carve(n, nullptr, (size_t)R, &used);
    size_t off = used;
    auto take = [&](size_t f) { size_t o = off; off += (f + 3) / 4 * 4; return o; };
    const size_t o_dzh = take((size_t)R * H), o_dhc = take((size_t)R * 3 * H), o_dt1 = take((size_t)R * T);
    const size_t o_part = take((size_t)max_sp * H * 3 * H);
    const size_t o_cs = take((size_t)cs_chunks * 2 * std::max(3 * H, T));
    const size_t o_lp = take((size_t)loss_blocks * 4 + 4);
    rc = vnet_ws(n, off);
    if (rc != FRL_OK) return rc;
    VBuf b = carve(n, n->ws, (size_t)R, &used);
    float* ws = n->ws;
    float *dZh = ws + o_dzh, *dHc = ws + o_dhc, *dT1 = ws + o_dt1, *part = ws + o_part, *cs = ws + o_cs;
    double* lpart = reinterpret_cast<double*>(ws + o_lp);
    const float* pk = n->pack;

    // ---- forward on the 2B stacked rows ----
    const long long W = n->Sp + n->Ap + T;
    vnet_prepare_kernel<<<(unsigned)(((long long)R * W + 255) / 256), 256, 0, st>>>(s, actions, x0, t, 0.f, pk + n->o_wt1, pk + n->o_bt1,
                                                                                  b.Sp_in, b.Xp_in, b.T1, b.Tv, B, R, S, A, n->Sp,
                                                                                  n->Ap, T, 1, 1);
    FRL_CUDA(cudaGetLastError());
    rc = vnet_trunk(n, b, R, true, st);
    if (rc != FRL_OK) return rc;
    bcf_loss_kernel<<<loss_blocks, 256, 0, st>>>(b.V, Aop, actions, x0, B, A, (float)((double)grad_scale * 2.0 / div), bc_coef, lpart);
    FRL_CUDA(cudaGetLastError());
    if (losses) {
        bcf_loss_final_kernel<<<1, 256, 0, st>>>(lpart, loss_blocks, 1.0 / div, bc_coef, losses);
        FRL_CUDA(cudaGetLastError());
    }
    if (!grad) return FRL_OK;

    auto ksplit = [&](int splits) { return round_up((R + splits - 1) / splits, BK); };
    // wgrad: grad[off .. off + rows*K) (+)= dZ^T X, dZ [R][ldz] (first `rows` columns used), X [R][ldx] (K columns used)
    auto wgrad = [&](const float* dZ, int ldz, int rows, const float* X, int ldx, int Kp, int K, long long goff, int sp) -> int {
        int r = run_gemm<true, EPI_NONE>(dZ, ldz, X, ldx, part, Kp, rows, Kp, R, nullptr, nullptr, 0, sp, ksplit(sp), st);
        if (r != FRL_OK) return r;
        reduce_wgrad_kernel<<<(rows * K + 255) / 256, 256, 0, st>>>(part, sp, (size_t)rows * Kp, Kp, rows, K, grad + goff, accumulate);
        FRL_CUDA(cudaGetLastError());
        return FRL_OK;
    };
    auto bgrad = [&](const float* dZ, int ldz, int cols, long long goff) -> int {
        colsum_partial_kernel<<<dim3((cols + 127) / 128, cs_chunks), 128, 0, st>>>(dZ, ldz, R, cols, cs_rows, cs);
        reduce_bias_kernel<<<(cols + 127) / 128, 128, 0, st>>>(cs, cs_chunks, cols, 0, cols, grad + goff, accumulate);
        FRL_CUDA(cudaGetLastError());
        return FRL_OK;
    };
    // ---- backward ----
    rc = wgrad(b.V, Aop, A, b.H1, H, H, H, n->off[10], sp_small);                          // velocity_head.2
    if (rc != FRL_OK) return rc;
    rc = bgrad(b.V, Aop, A, n->off[11]);
    if (rc != FRL_OK) return rc;
    rc = run_gemm<false, EPI_MASK>(b.V, Aop, pk + n->o_W2, H, dZh, H, R, H, Aop, nullptr, b.H1, H, 1, Aop, st);
    if (rc != FRL_OK) return rc;
    rc = wgrad(dZh, H, H, b.Hc, 3 * H, 3 * H, 3 * H, n->off[8], sp_big);                  // velocity_head.0
    if (rc != FRL_OK) return rc;
    rc = bgrad(dZh, H, H, n->off[9]);
    if (rc != FRL_OK) return rc;
    rc = run_gemm<false, EPI_MASK>(dZh, H, pk + n->o_W1, 3 * H, dHc, 3 * H, R, 3 * H, H, nullptr, b.Hc, 3 * H, 1, H, st);
    if (rc != FRL_OK) return rc;
    rc = wgrad(dHc, 3 * H, H, b.Sp_in, n->Sp, n->Sp, S, n->off[4], sp_small);              // state_embed
    if (rc != FRL_OK) return rc;
    rc = bgrad(dHc, 3 * H, H, n->off[5]);
    if (rc != FRL_OK) return rc;
    rc = wgrad(dHc + 2 * H, 3 * H, H, b.Xp_in, n->Ap, n->Ap, A, n->off[6], sp_small);      // action_embed
    if (rc != FRL_OK) return rc;
    rc = bgrad(dHc + 2 * H, 3 * H, H, n->off[7]);
    if (rc != FRL_OK) return rc;
    rc = wgrad(dHc + H, 3 * H, H, b.T1, T, T, T, n->off[2], sp_hh);                        // time_embed.2
    if (rc != FRL_OK) return rc;
    rc = bgrad(dHc + H, 3 * H, H, n->off[3]);
    if (rc != FRL_OK) return rc;
    rc = run_gemm<false, EPI_MASK>(dHc + H, 3 * H, pk + n->o_Wt2, T, dT1, T, R, T, H, nullptr, b.T1, T, 1, H, st);
    if (rc != FRL_OK) return rc;
    time_grad_partial_kernel<<<dim3((T + 127) / 128, cs_chunks), 128, 0, st>>>(dT1, b.Tv, R, T, cs_rows, cs);   // time_embed.0
    time_grad_reduce_kernel<<<(T + 127) / 128, 128, 0, st>>>(cs, cs_chunks, T, grad + n->off[0], grad + n->off[1], accumulate);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}
