// FP32 (FFMA) fused log-density integrator and velocity-field forward.
//
// One persistent CTA per SM walks over tiles of TS = 32 samples x role.  For a tile, ALL n_steps Euler steps of
//   a_{k+1} = a_k + delta * v(a_k, s, t_k),   log_det -= delta * probe^T (dv/da) probe
// (modril/flowril/pretrain.py:192-214 / :112-134) run without leaving the SM: the primal activations and the
// forward-mode tangent (which replaces the reference's autograd VJP, pretrain.py:163-177 -- same scalar, see
// oracle/closed_form.py) live in shared memory, k-major, and every weight chunk streamed from L2 through a
// cp.async double buffer is used for both.  Each thread owns a 4-sample x (H/32)-column register tile of BOTH the
// primal and the tangent, so the ReLU mask of the primal is applied to the tangent in registers.
//
// This is the exact-arithmetic (1e-4 parity) path; the tensor-core path is logprob_tc.cu.
#include <cmath>

#include "frl_internal.h"

namespace frl {

constexpr int TS = 32;        // samples per tile
constexpr int KC = 16;        // k-rows per weight chunk
constexpr int NTHREADS = 256;

struct LogprobK {
    NetF32 net[2];
    float sign[2];
    float* out[2];       // per role slot: logp [B]   (vfield: out[0] = v_c / v, out[1] = tanh(r))
    const float* s;
    const float* a;
    const float* probe;  // may be null (basis >= 0 or vfield)
    const float* t_rows; // vfield only: per-row t
    long long probe_step_stride;  // 0: shared probe
    long long B;
    int n_roles, n_heads, depth, s_dim, a_dim, d_in, d_in_p, n_out;
    int n_steps, basis, accumulate;
    float delta;
    float t_mid[kMaxSteps];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int H>
__device__ __forceinline__ void load_chunk(float* dst, const float* __restrict__ src) {
    constexpr int V = KC * H / 4;  // float4 count
#pragma unroll
    for (int i = 0; i < V / NTHREADS; ++i) {
        int idx = threadIdx.x + i * NTHREADS;
        cp_async16(dst + idx * 4, src + idx * 4);
    }
}

// One trunk layer for the tile: Xout = relu(W Xin + b) and, if TAN, Xout_t = mask * (W[:, :Kt] Xin_t).
template <int H, bool TAN>
__device__ __forceinline__ void trunk_layer(const float* __restrict__ Wt_g, const float* __restrict__ bias_g, int Kp,
                                            int Kt, const float* Xin_p, const float* Xin_t, float* Xout_p,
                                            float* Xout_t, float* Wc) {
    constexpr int CG = H / 128;  // groups of 4 columns per thread
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = warp * 4;
    float accp[4][CG * 4];
    float acct[4][CG * 4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CG * 4; ++c) { accp[r][c] = 0.f; acct[r][c] = 0.f; }

    const int nchunk = Kp / KC;
    load_chunk<H>(Wc, Wt_g);
    cp_async_commit();
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            load_chunk<H>(Wc + ((c + 1) & 1) * KC * H, Wt_g + (size_t)(c + 1) * KC * H);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* wc = Wc + (c & 1) * KC * H;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int k = c * KC + kk;
            const float4 ap = *reinterpret_cast<const float4*>(Xin_p + k * TS + r0);
            float4 w[CG];
#pragma unroll
            for (int g = 0; g < CG; ++g) w[g] = *reinterpret_cast<const float4*>(wc + kk * H + g * 128 + lane * 4);
            const float apv[4] = {ap.x, ap.y, ap.z, ap.w};
#pragma unroll
            for (int g = 0; g < CG; ++g) {
                const float wv[4] = {w[g].x, w[g].y, w[g].z, w[g].w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j) accp[r][g * 4 + j] = fmaf(apv[r], wv[j], accp[r][g * 4 + j]);
            }
            if (TAN && k < Kt) {
                const float4 at = *reinterpret_cast<const float4*>(Xin_t + k * TS + r0);
                const float atv[4] = {at.x, at.y, at.z, at.w};
#pragma unroll
                for (int g = 0; g < CG; ++g) {
                    const float wv[4] = {w[g].x, w[g].y, w[g].z, w[g].w};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acct[r][g * 4 + j] = fmaf(atv[r], wv[j], acct[r][g * 4 + j]);
                }
            }
        }
        __syncthreads();
    }
    // epilogue: bias, ReLU (strict mask, relu'(0) = 0 like threshold_backward), mask the tangent
#pragma unroll
    for (int g = 0; g < CG; ++g) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = g * 128 + lane * 4 + j;
            const float bias = __ldg(bias_g + n);
            float hp[4], ht[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float z = accp[r][g * 4 + j] + bias;
                const bool m = z > 0.f;
                hp[r] = m ? z : 0.f;
                ht[r] = m ? acct[r][g * 4 + j] : 0.f;
            }
            *reinterpret_cast<float4*>(Xout_p + n * TS + r0) = make_float4(hp[0], hp[1], hp[2], hp[3]);
            if (TAN) *reinterpret_cast<float4*>(Xout_t + n * TS + r0) = make_float4(ht[0], ht[1], ht[2], ht[3]);
        }
    }
    __syncthreads();
}

// Heads: out[o][r] = sum_k Wh[o][k] X[k][r] (+ bias for the primal).  Warp w handles outputs w, w+8, ...;
// lane = sample.  Wh rows are read warp-uniformly from global/L1.
template <int H, bool TAN>
__device__ __forceinline__ void heads(const float* __restrict__ Wh, const float* __restrict__ bh, int n_out,
                                      const float* Xp, const float* Xt, float* Hout_p, float* Hout_t) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < n_out; o += NTHREADS / 32) {
        const float* wrow = Wh + (size_t)o * H;
        float accp = 0.f, acct = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + k));
            accp = fmaf(w.x, Xp[(k + 0) * TS + lane], accp);
            accp = fmaf(w.y, Xp[(k + 1) * TS + lane], accp);
            accp = fmaf(w.z, Xp[(k + 2) * TS + lane], accp);
            accp = fmaf(w.w, Xp[(k + 3) * TS + lane], accp);
            if (TAN) {
                acct = fmaf(w.x, Xt[(k + 0) * TS + lane], acct);
                acct = fmaf(w.y, Xt[(k + 1) * TS + lane], acct);
                acct = fmaf(w.z, Xt[(k + 2) * TS + lane], acct);
                acct = fmaf(w.w, Xt[(k + 3) * TS + lane], acct);
            }
        }
        Hout_p[o * TS + lane] = accp + __ldg(bh + o);
        if (TAN) Hout_t[o * TS + lane] = acct;
    }
    __syncthreads();
}

template <int H, bool VFIELD>
__global__ void __launch_bounds__(NTHREADS, 1) logprob_f32_kernel(const __grid_constant__ LogprobK p) {
    constexpr bool TAN = !VFIELD;
    extern __shared__ __align__(16) float smem[];
    float* X0p = smem;                             // [d_in_p][TS]
    float* X0t = X0p + p.d_in_p * TS;              // [32][TS] (a_dim <= 32)
    float* Xp = X0t + 32 * TS;                     // [2][H][TS]
    float* Xt = Xp + 2 * H * TS;                   // [2][H][TS]
    float* Wc = Xt + (TAN ? 2 * H * TS : 0);       // [2][KC][H]
    float* Hout_p = Wc + 2 * KC * H;               // [n_out][TS]
    float* Hout_t = Hout_p + 64 * TS;              // [n_out][TS]  (n_out <= 64)

    const int tid = threadIdx.x;
    const long long tiles_per_role = (p.B + TS - 1) / TS;
    const long long n_tiles = tiles_per_role * p.n_roles;
    const int t_row = p.a_dim + p.s_dim;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int role = (int)(tile % p.n_roles);
        const long long row0 = (tile / p.n_roles) * TS;
        const NetF32& net = p.net[role];
        const float sign = p.sign[role];

        __syncthreads();  // previous tile fully consumed
        for (int idx = tid; idx < p.d_in_p * TS; idx += NTHREADS) {
            const int k = idx / TS, r = idx % TS;
            const long long row = row0 + r;
            float v = 0.f;
            if (row < p.B) {
                if (k < p.a_dim) v = p.a[row * p.a_dim + k];
                else if (k < t_row) v = p.s[row * p.s_dim + (k - p.a_dim)];
                else if (k == t_row) v = VFIELD ? p.t_rows[row] : p.t_mid[0];
            }
            X0p[idx] = v;
        }
        if (TAN && p.probe_step_stride == 0) {
            for (int idx = tid; idx < p.a_dim * TS; idx += NTHREADS) {
                const int k = idx / TS, r = idx % TS;
                const long long row = row0 + r;
                float v = 0.f;
                if (row < p.B) v = p.basis >= 0 ? (k == p.basis ? 1.f : 0.f) : p.probe[row * p.a_dim + k];
                X0t[idx] = v;
            }
        }
        float log_det = 0.f;
        const int n_steps = VFIELD ? 1 : p.n_steps;
        for (int step = 0; step < n_steps; ++step) {
            if (TAN && p.probe_step_stride != 0) {
                __syncthreads();
                for (int idx = tid; idx < p.a_dim * TS; idx += NTHREADS) {
                    const int k = idx / TS, r = idx % TS;
                    const long long row = row0 + r;
                    X0t[idx] = row < p.B ? p.probe[step * p.probe_step_stride + row * p.a_dim + k] : 0.f;
                }
            }
            // (layer 0's first __syncthreads orders the input writes above / the state update below)
            trunk_layer<H, TAN>(net.Wt[0], net.b[0], p.d_in_p, p.a_dim, X0p, X0t, Xp, Xt, Wc);
            int cur = 0;
            for (int l = 1; l < p.depth; ++l) {
                trunk_layer<H, TAN>(net.Wt[l], net.b[l], H, H, Xp + cur * H * TS, Xt + cur * H * TS,
                                    Xp + (cur ^ 1) * H * TS, Xt + (cur ^ 1) * H * TS, Wc);
                cur ^= 1;
            }
            heads<H, TAN>(net.Wh, net.bh, p.n_out, Xp + cur * H * TS, Xt + cur * H * TS, Hout_p, Hout_t);

            if (tid < TS) {
                const int r = tid;
                const long long row = row0 + r;
                if (VFIELD) {
                    if (row < p.B) {
                        for (int j = 0; j < p.a_dim; ++j) {
                            p.out[0][row * p.a_dim + j] = Hout_p[j * TS + r];
                            if (p.n_heads == 2 && p.out[1])
                                p.out[1][row * p.a_dim + j] = tanhf(Hout_p[(p.a_dim + j) * TS + r]);
                        }
                    }
                } else {
                    float div = 0.f;
                    for (int j = 0; j < p.a_dim; ++j) {
                        float v = Hout_p[j * TS + r];
                        float dv = Hout_t[j * TS + r];
                        if (p.n_heads == 2) {
                            const float th = tanhf(Hout_p[(p.a_dim + j) * TS + r]);
                            v += sign * th;
                            dv += sign * (1.f - th * th) * Hout_t[(p.a_dim + j) * TS + r];
                        }
                        div = fmaf(dv, X0t[j * TS + r], div);
                        X0p[j * TS + r] += v * p.delta;                     // pretrain.py:211 / :131
                    }
                    log_det -= p.delta * div;                                // pretrain.py:210 / :130
                    if (step + 1 < n_steps) X0p[t_row * TS + r] = p.t_mid[step + 1];
                }
            }
        }
        if (!VFIELD && tid < TS) {
            const long long row = row0 + tid;
            if (row < p.B) {
                if (p.accumulate) {
                    p.out[role][row] += log_det;
                } else {
                    float prior = 0.f;
                    for (int j = 0; j < p.a_dim; ++j) {
                        const float x = X0p[j * TS + tid];
                        prior += -0.5f * x * x - 0.91893853320467274178f;    // Normal(0,1).log_prob, pretrain.py:213
                    }
                    p.out[role][row] = prior + log_det;
                }
            }
        }
    }
}

__global__ void reward_kernel(const float* __restrict__ le, const float* __restrict__ la, float* __restrict__ out,
                              long long B, int type) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float r = le[i] - la[i];  // flowril.py:191
    float y;
    switch (type) {
        case FRL_REWARD_NORM: {
            const float sg = 1.f / (1.f + expf(-r));
            y = logf(sg + 1e-20f) - logf(1.f - sg + 1e-20f);
            break;
        }
        case FRL_REWARD_EXP: { const float e = expf(r); y = -e / (1.f + e); break; }
        case FRL_REWARD_CLIP: y = fminf(fmaxf(r, -20.f), 20.f); break;
        case FRL_REWARD_SMOOTH: y = -log1pf(expf(-r)); break;
        default: y = r;
    }
    out[i] = y;
}

static int sm_count(int device) {
    static int cached[64] = {0};
    if (device < 64 && cached[device]) return cached[device];
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (device < 64) cached[device] = n;
    return n;
}

template <int H, bool VFIELD>
static int launch(const LogprobK& p, int device, cudaStream_t stream) {
    const bool tan = !VFIELD;
    size_t floats = (size_t)p.d_in_p * TS + 32 * TS + 2 * H * TS + (tan ? 2 * H * TS : 0) + 2 * KC * H + 2 * 64 * TS;
    size_t bytes = floats * sizeof(float);
    auto kern = logprob_f32_kernel<H, VFIELD>;
    FRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    long long n_tiles = ((p.B + TS - 1) / TS) * p.n_roles;
    int grid = (int)std::min<long long>(n_tiles, sm_count(device));
    if (grid < 1) return FRL_OK;
    kern<<<grid, NTHREADS, bytes, stream>>>(p);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

static void fill_common(LogprobK& p, const frl_net* n) {
    p.n_heads = n->cfg.n_heads;
    p.depth = n->cfg.depth;
    p.s_dim = n->cfg.s_dim;
    p.a_dim = n->cfg.a_dim;
    p.d_in = n->d_in;
    p.d_in_p = n->d_in_p;
    p.n_out = n->n_out;
}

int launch_reward(const float* le, const float* la, float* reward, long long B, int reward_type, cudaStream_t stream) {
    if (B <= 0) return FRL_OK;
    reward_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(le, la, reward, B, reward_type);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

int launch_logprob_f32(const ScoreArgs& a, cudaStream_t stream) {
    const frl_net* n0 = a.net[0];
    for (int r = 0; r < a.n_roles; ++r) {
        int rcp = ensure_pack(a.net[r], PACK_F32, stream);
        if (rcp != FRL_OK) return rcp;
    }
    LogprobK p{};
    fill_common(p, n0);
    for (int r = 0; r < a.n_roles; ++r) {
        p.net[r] = a.net[r]->f32;
        p.sign[r] = a.sign[r];
        p.out[r] = a.out_logp[r];
    }
    p.s = a.s; p.a = a.a; p.probe = a.probe; p.t_rows = nullptr;
    p.B = a.B; p.n_roles = a.n_roles; p.n_steps = a.n_steps;
    p.delta = (float)(1.0 / a.n_steps);  // python float 1.0/n_steps folded to fp32 (pretrain.py:201)
    for (int k = 0; k < a.n_steps; ++k) p.t_mid[k] = a.t_mid_host[k];
    p.probe_step_stride = a.probe_mode == FRL_PROBE_PER_STEP ? a.B * (long long)p.a_dim : 0;
    const int H = n0->cfg.hidden;
    auto run = [&](const LogprobK& q) {
        return H == 256 ? launch<256, false>(q, n0->cfg.device, stream) : launch<128, false>(q, n0->cfg.device, stream);
    };
    if (a.probe_mode == FRL_PROBE_EXACT) {
        // trace = sum_i e_i^T J e_i: a_dim passes over the same primal trajectory; pass 0 writes prior + its
        // share of log_det, the others add theirs.
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

int launch_vfield_f32(const frl_net* net, const float* s, const float* a, const float* t, float* out0, float* out1,
                      long long B, cudaStream_t stream) {
    {
        int rcp = ensure_pack(net, PACK_F32, stream);
        if (rcp != FRL_OK) return rcp;
    }
    LogprobK p{};
    fill_common(p, net);
    p.net[0] = net->f32;
    p.sign[0] = 1.f;
    p.out[0] = out0;
    p.out[1] = out1;
    p.s = s; p.a = a; p.probe = nullptr; p.t_rows = t;
    p.B = B; p.n_roles = 1; p.n_steps = 1; p.delta = 0.f; p.basis = -1;
    return net->cfg.hidden == 256 ? launch<256, true>(p, net->cfg.device, stream)
                                  : launch<128, true>(p, net->cfg.device, stream);
}

}  // namespace frl
