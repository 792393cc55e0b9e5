// PPO consumer of the rewards (scope row f4, second part): the actor-critic policy the reference builds for
// --alg flowril and the clipped-PPO minibatch loss + gradient, FP32 on the SGEMM building block of sgemm.cuh.
// Replaces  DistActorCritic.forward / get_action / evaluate_actions
// (deps/rl-toolkit/rlf/policies/actor_critic/dist_actor_critic.py:48-86; MLPBasic actor / critic with tanh after every
// layer, rlf/rl/model.py:246-296; critic_head Linear(H,1), rlf/policies/utils.py:16-19; DiagGaussian,
// rlf/rl/distributions.py:60-75)  and the loss / backward of  PPO.update  (rlf/algos/on_policy/ppo.py:23-49).
//
//   value = wv . hc_L + bv,   hc_l = tanh(Wc_l hc_{l-1} + bc_l),  hc_0 = obs
//   mean  = Wm ha_L + bm,     ha_l = tanh(Wa_l ha_{l-1} + ba_l),  ha_0 = obs,   std = exp(logstd)
// The minibatch gather of RolloutStorage.feed_forward_generator (rlf/storage/rollout_storage.py:270-323) is folded into
// the kernels (optional int64 row index).  Batch reductions are split and reduced in a fixed order (deterministic).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "frl_internal.h"
#include "sgemm.cuh"

constexpr int kAcMaxLayers = 8;

struct frl_acnet {
    frl_acnet_config cfg;
    int Sp, Aop;                                   // obs width / action width padded to 16
    // flat offsets (state_dict order): [net 0 = critic, 1 = actor][layer]
    long long off_w[2][kAcMaxLayers], off_b[2][kAcMaxLayers], off_hw[2], off_hb[2], off_logstd;
    long long n_params;
    float* pack = nullptr;
    size_t o_Wt[2][kAcMaxLayers], o_W[2][kAcMaxLayers], o_b[2][kAcMaxLayers], o_Wht[2], o_Wh[2], o_bh[2], o_logstd;
    size_t pack_floats = 0;
    const float* params_dev = nullptr;
    unsigned long long version = 0, packed = 0;
    float* ws = nullptr;
    size_t ws_floats = 0;
    // Workspaces that were outgrown stay allocated until the handle is destroyed: a captured CUDA graph (PPOUpdateGraph)
    // may still reference them, and a later, larger eager call must not pull the memory from under its replays.
    std::vector<float*> retired;
};

namespace frl {

struct APackJob {
    long long src_off, dst_off, first;
    int rows, cols, dst_ld, transpose;
};
struct APackTable {
    APackJob job[6 * kAcMaxLayers + 8];
    int n_jobs;
    long long total;
};
__global__ void acnet_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, APackTable tab) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tab.total; i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < tab.n_jobs && tab.job[j + 1].first <= i) ++j;
        const APackJob& jb = tab.job[j];
        const long long e = i - jb.first;
        const int r = (int)(e / jb.cols), c = (int)(e % jb.cols);
        const float v = src[jb.src_off + e];
        if (jb.transpose) dst[jb.dst_off + (long long)c * jb.dst_ld + r] = v;
        else dst[jb.dst_off + (long long)r * jb.dst_ld + c] = v;
    }
}

static int acnet_ensure_pack(frl_acnet* n, cudaStream_t st) {
    if (n->packed == n->version) return FRL_OK;
    const int H = n->cfg.hidden, S = n->cfg.obs_dim, A = n->cfg.a_dim, L = n->cfg.layers;
    FRL_CUDA(cudaMemsetAsync(n->pack, 0, n->pack_floats * sizeof(float), st));
    APackTable tab{};
    long long total = 0;
    auto add = [&](long long src, size_t dst, int rows, int cols, int dst_ld, int transpose) {
        APackJob& j = tab.job[tab.n_jobs++];
        j = APackJob{src, (long long)dst, total, rows, cols, dst_ld, transpose};
        total += (long long)rows * cols;
    };
    for (int net = 0; net < 2; ++net) {
        for (int l = 0; l < L; ++l) {
            const int K = l == 0 ? S : H, Kp = l == 0 ? n->Sp : H;
            add(n->off_w[net][l], n->o_Wt[net][l], H, K, H, 1);     // [Kp][H]
            add(n->off_w[net][l], n->o_W[net][l], H, K, Kp, 0);     // [H][Kp]
            add(n->off_b[net][l], n->o_b[net][l], 1, H, H, 0);
        }
        const int NO = net == 0 ? 1 : A;
        add(n->off_hw[net], n->o_Wht[net], NO, H, n->Aop, 1);       // [H][Aop]
        add(n->off_hw[net], n->o_Wh[net], NO, H, H, 0);             // [Aop][H]
        add(n->off_hb[net], n->o_bh[net], 1, NO, n->Aop, 0);
    }
    add(n->off_logstd, n->o_logstd, 1, A, n->Aop, 0);
    tab.total = total;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 592);
    acnet_pack_kernel<<<blocks, 256, 0, st>>>(n->params_dev, n->pack, tab);
    FRL_CUDA(cudaGetLastError());
    n->packed = n->version;
    return FRL_OK;
}

static int acnet_ws(frl_acnet* n, size_t floats) {
    if (n->ws_floats >= floats) return FRL_OK;
    if (n->ws) n->retired.push_back(n->ws);
    n->ws = nullptr;
    n->ws_floats = 0;
    FRL_CUDA(cudaMalloc(&n->ws, floats * sizeof(float)));
    n->ws_floats = floats;
    return FRL_OK;
}

// observations (optionally gathered through idx) into the zero-padded GEMM operand [B][Sp]
__global__ void acnet_gather_obs_kernel(const float* __restrict__ obs, const long long* __restrict__ idx,
                                        float* __restrict__ Xp, long long B, int S, int Sp) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B * Sp) return;
    const long long r = i / Sp;
    const int c = (int)(i % Sp);
    const long long src = idx ? idx[r] : r;
    Xp[i] = c < S ? obs[src * S + c] : 0.f;
}

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;   // math.log(math.sqrt(2 * math.pi)), torch Normal.log_prob

// FixedNormal.log_probs of one row (distributions.py:31-34) in torch's operation order
__device__ __forceinline__ float row_logp(const float* __restrict__ x, const float* __restrict__ mean,
                                          const float* __restrict__ logstd, int A) {
    float acc = 0.f;
    for (int j = 0; j < A; ++j) {
        const float scale = expf(logstd[j]);
        const float d = x[j] - mean[j];
        acc += -(d * d) / (2.f * (scale * scale)) - logf(scale) - kLogSqrt2Pi;
    }
    return acc;
}
__device__ __forceinline__ float row_entropy(const float* __restrict__ logstd, int A) {
    float acc = 0.f;
    for (int j = 0; j < A; ++j) acc += 0.5f + kLogSqrt2Pi + logf(expf(logstd[j]));   // Normal.entropy summed (:36-37)
    return acc;
}

// get_action / evaluate_actions epilogue.  mode 0: outputs value, mean.  mode 1 (act): action = mean + std * noise
// (noise NULL: dist.mode()), log-prob of it.  mode 2 (evaluate): log-prob of the given action, entropy.
__global__ void acnet_head_kernel(const float* __restrict__ V, const float* __restrict__ M, int ld,
                                  const float* __restrict__ logstd, const float* __restrict__ noise,
                                  const float* __restrict__ action_in, float* __restrict__ value, float* __restrict__ mean_out,
                                  float* __restrict__ action_out, float* __restrict__ logp, float* __restrict__ ent, long long B,
                                  int A, int mode) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float* m = M + r * ld;
    if (value) value[r] = V[r * ld];
    if (mean_out)
        for (int j = 0; j < A; ++j) mean_out[r * A + j] = m[j];
    if (mode == 1) {
        float acc = 0.f;
        for (int j = 0; j < A; ++j) {
            const float scale = expf(logstd[j]);
            const float x = noise ? __fadd_rn(m[j], __fmul_rn(scale, noise[r * A + j])) : m[j];
            action_out[r * A + j] = x;
            const float d = x - m[j];
            acc += -(d * d) / (2.f * (scale * scale)) - logf(scale) - kLogSqrt2Pi;
        }
        if (logp) logp[r] = acc;
    } else if (mode == 2) {
        if (logp) logp[r] = row_logp(action_in + r * A, m, logstd, A);
    }
    if (ent) ent[r] = row_entropy(logstd, A);
}

// ppo.py:29-49 per row + dL/dvalue and dL/dmean written in place of value / mean.  partial[block] = {sum of
// max(vl, vlc), sum of min(surr1, surr2), d logstd_j sums (A)} in double.
__global__ void __launch_bounds__(256) ppo_loss_kernel(float* __restrict__ V, float* __restrict__ M, int ld,
                                                       const float* __restrict__ logstd, const float* __restrict__ action,
                                                       const float* __restrict__ old_logp, const float* __restrict__ adv,
                                                       const float* __restrict__ ret, const float* __restrict__ old_value,
                                                       const long long* __restrict__ idx, long long B, int A, float clip,
                                                       float scale_actor, float scale_value, double* __restrict__ partial) {
    extern __shared__ double sh_p[];   // [8 warps][2 + A]
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int W = 2 + A, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double vsum = 0.0, asum = 0.0;
    float d_logp = 0.f;
    const long long src = r < B ? (idx ? idx[r] : r) : 0;
    if (r < B) {
        float* m = M + r * ld;
        const float value = V[r * ld];
        const float logp = row_logp(action + src * A, m, logstd, A);
        const float ratio = expf(logp - old_logp[src]);
        const float ad = adv[src];
        const float surr1 = ratio * ad;
        const float surr2 = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip) * ad;
        asum = (double)fminf(surr1, surr2);
        // torch.min splits the gradient on ties; clamp passes it on the closed interval
        const float w1 = surr1 < surr2 ? 1.f : (surr1 > surr2 ? 0.f : 0.5f);
        const float in_r = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? 1.f : 0.f;
        d_logp = -(ad * scale_actor) * (w1 + (1.f - w1) * in_r) * ratio;
        const float ov = old_value[src], rt = ret[src];
        const float dv = value - ov;
        const float vpc = ov + fminf(fmaxf(dv, -clip), clip);
        const float e1 = value - rt, e2 = vpc - rt;
        const float vl = e1 * e1, vlc = e2 * e2;
        vsum = (double)fmaxf(vl, vlc);
        const float wv = vl > vlc ? 1.f : (vl < vlc ? 0.f : 0.5f);
        const float in_v = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
        V[r * ld] = scale_value * (wv * 2.f * e1 + (1.f - wv) * 2.f * e2 * in_v);
        for (int j = 1; j < ld; ++j) V[r * ld + j] = 0.f;
    }
    for (int o = 16; o > 0; o >>= 1) {
        vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        asum += __shfl_xor_sync(0xffffffffu, asum, o);
    }
    if (lane == 0) { sh_p[warp * W] = vsum; sh_p[warp * W + 1] = asum; }
    for (int j = 0; j < A; ++j) {
        double g = 0.0;
        if (r < B) {
            float* m = M + r * ld;
            const float scale = expf(logstd[j]);
            const float var = scale * scale;
            const float d = action[src * A + j] - m[j];
            g = (double)(d_logp * (d * d / var - 1.f));
            m[j] = d_logp * d / var;
        }
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == 0) sh_p[warp * W + 2 + j] = g;
    }
    if (r < B)
        for (int j = A; j < ld; ++j) M[r * ld + j] = 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        double acc = 0.0;
        for (int w = 0; w < 8; ++w) acc += sh_p[w * W + c];
        partial[(size_t)blockIdx.x * W + c] = acc;
    }
}

// losses = {loss, value_loss, actor_loss, entropy}; d logstd (+)= sums - entropy_coef * local share
__global__ void ppo_loss_final_kernel(const double* __restrict__ partial, int n_blocks, int A, double inv_div, float value_coef,
                                      float ent_coef, float ent_share, float grad_scale, const float* __restrict__ logstd,
                                      float* __restrict__ losses, float* __restrict__ g_logstd, int accumulate) {
    const int c = threadIdx.x;
    double acc = 0.0;
    if (c < 2 + A)
        for (int b = 0; b < n_blocks; ++b) acc += partial[(size_t)b * (2 + A) + c];
    __shared__ float sh[2];
    if (c < 2) sh[c] = c == 0 ? (float)(0.5 * acc * inv_div) : (float)(-acc * inv_div);
    __syncthreads();
    if (c >= 2 && c < 2 + A && g_logstd) {
        const float g = (float)acc - grad_scale * ent_coef * ent_share;
        g_logstd[c - 2] = accumulate ? g_logstd[c - 2] + g : g;
    }
    if (c == 0 && losses) {
        const float ent = row_entropy(logstd, A);
        losses[0] = sh[0] * value_coef + sh[1] - ent * ent_coef * ent_share;   // ppo.py:48-49 (this rank's share)
        losses[1] = sh[0];
        losses[2] = sh[1];
        losses[3] = ent;
    }
}

struct ABuf {
    float *Xp, *act[2], *out[2];   // act[net]: L activation matrices [B][H] back to back; out[net]: [B][Aop]
};

static ABuf ac_carve(const frl_acnet* n, float* ws, size_t B, size_t* used) {
    const size_t H = n->cfg.hidden, L = n->cfg.layers;
    size_t off = 0;
    auto take = [&](size_t f) { size_t o = off; off += (f + 3) / 4 * 4; return ws ? ws + o : (float*)nullptr; };
    ABuf b;
    b.Xp = take(B * n->Sp);
    for (int net = 0; net < 2; ++net) { b.act[net] = take(B * H * L); b.out[net] = take(B * n->Aop); }
    *used = off;
    return b;
}

// both trunks + heads for B rows whose padded observations are in b.Xp (net_mask: bit 0 critic, bit 1 actor)
static int ac_forward(frl_acnet* n, const ABuf& b, int B, int net_mask, cudaStream_t st) {
    const int H = n->cfg.hidden, L = n->cfg.layers;
    const float* pk = n->pack;
    for (int net = 0; net < 2; ++net) {
        if (!(net_mask >> net & 1)) continue;
        const float* x = b.Xp;
        int K = n->Sp;
        for (int l = 0; l < L; ++l) {
            float* h = b.act[net] + (size_t)l * B * H;
            int rc = run_gemm<false, EPI_BIAS_TANH>(x, K, pk + n->o_Wt[net][l], H, h, H, B, H, K, pk + n->o_b[net][l], nullptr, 0, 1, K, st);
            if (rc != FRL_OK) return rc;
            x = h;
            K = H;
        }
        int rc = run_gemm<false, EPI_BIAS>(x, H, pk + n->o_Wht[net], n->Aop, b.out[net], n->Aop, B, n->Aop, H, pk + n->o_bh[net], nullptr, 0, 1, H, st);
        if (rc != FRL_OK) return rc;
    }
    return FRL_OK;
}

static int ac_prepare(frl_acnet* n, const float* obs, const long long* idx, long long B, size_t extra_floats, ABuf* out,
                      size_t* used, cudaStream_t st) {
    int rc = acnet_ensure_pack(n, st);
    if (rc != FRL_OK) return rc;
    ac_carve(n, nullptr, (size_t)B, used);
    rc = acnet_ws(n, *used + extra_floats);
    if (rc != FRL_OK) return rc;
    *out = ac_carve(n, n->ws, (size_t)B, used);
    acnet_gather_obs_kernel<<<(unsigned)((B * n->Sp + 255) / 256), 256, 0, st>>>(obs, idx, out->Xp, B, n->cfg.obs_dim, n->Sp);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

// ---- fused minibatch kernel for nets that fit in shared memory (the reference's 64 x 2 default) ------------------------
// One CTA keeps BOTH nets' weights in shared memory (natural [H][K] layout read straight from the flat parameter vector,
// rows padded to K+1 so that the forward pass (threads over j) and the dgrad (threads over k) are both bank-conflict
// free) and walks its row tiles: gather -> critic / actor forward -> PPO loss per row -> both backward passes -> weight
// gradient partials.  A CTA accumulates its partials in its own row of `part` (no atomics: fixed tile -> CTA mapping, fixed
// order), ppo_fused_reduce_kernel sums the rows.  Two launches per minibatch instead of ~35.
struct FusedOffs {
    int off_w[2][kAcMaxLayers], off_b[2][kAcMaxLayers], off_hw[2], off_hb[2], off_logstd, n_params;
};

__device__ __forceinline__ void emit(float* p, float v, bool first) { *p = first ? v : *p + v; }

__global__ void __launch_bounds__(512) ppo_fused_kernel(const float* __restrict__ params, FusedOffs fo, const float* __restrict__ obs,
                                                        const float* __restrict__ action, const float* __restrict__ old_logp,
                                                        const float* __restrict__ adv, const float* __restrict__ ret,
                                                        const float* __restrict__ old_value, const long long* __restrict__ idx,
                                                        long long B, int S, int A, int H, int L, int R, float clip,
                                                        float scale_actor, float scale_value, float* __restrict__ part,
                                                        int part_stride) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- carve shared memory ----
    // weights: [critic trunk][actor trunk][critic head][actor head]; layer l of a trunk at a closed-form offset
    const int L0 = H * (S + 1) + H, LH = H * (H + 1) + H, T = L0 + (L - 1) * LH;
    auto sWl = [&](int net, int l) { return sm + net * T + (l == 0 ? 0 : L0 + (l - 1) * LH); };
    auto sbl = [&](int net, int l) { return sWl(net, l) + H * ((l == 0 ? S : H) + 1); };
    auto sWhd = [&](int net) { return sm + 2 * T + (net == 0 ? 0 : (H + 1) + 1); };
    auto sbhd = [&](int net) { return sWhd(net) + (net == 0 ? 1 : A) * (H + 1); };
    float* p = sm + 2 * T + (1 + A) * (H + 1) + 1 + A;
    float* slog = p; p += A;
    float* X = p; p += R * S;
    float* act[2];
    for (int net = 0; net < 2; ++net) { act[net] = p; p += L * R * H; }
    float* dz0 = p; p += R * H;
    float* dz1 = p; p += R * H;
    float* outV = p; p += R;            // value, then dL/dvalue
    float* outM = p; p += R * A;        // mean, then dL/dmean
    float* rstat = p; p += R * (2 + A); // per row: max(vl, vlc), min(surr), d logstd_j contributions
    long long* src = reinterpret_cast<long long*>(p + ((p - sm) & 1));   // 8-byte aligned

    // ---- weights: once per CTA ----
    for (int net = 0; net < 2; ++net) {
        for (int l = 0; l < L; ++l) {
            const int K = l == 0 ? S : H;
            const float* g = params + fo.off_w[net][l];
            for (int i = tid; i < H * K; i += nt) sWl(net, l)[(i / K) * (K + 1) + i % K] = g[i];
            for (int i = tid; i < H; i += nt) sbl(net, l)[i] = params[fo.off_b[net][l] + i];
        }
        const int NO = net == 0 ? 1 : A;
        for (int i = tid; i < NO * H; i += nt) sWhd(net)[(i / H) * (H + 1) + i % H] = params[fo.off_hw[net] + i];
        for (int i = tid; i < NO; i += nt) sbhd(net)[i] = params[fo.off_hb[net] + i];
    }
    for (int i = tid; i < A; i += nt) slog[i] = params[fo.off_logstd + i];
    float* mypart = part + (size_t)blockIdx.x * part_stride;
    const long long n_tiles = (B + R - 1) / R;
    bool first = true;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        const long long row0 = tile * R;
        __syncthreads();   // weights loaded / previous tile done with the buffers
        for (int r = tid; r < R; r += nt) src[r] = row0 + r < B ? (idx ? idx[row0 + r] : row0 + r) : -1;
        __syncthreads();
        for (int i = tid; i < R * S; i += nt) { const long long q = src[i / S]; X[i] = q >= 0 ? obs[q * S + i % S] : 0.f; }
        __syncthreads();
        // ---- forward, both nets ----
        for (int l = 0; l < L; ++l) {
            const int K = l == 0 ? S : H;
            for (int o = tid; o < 2 * R * H; o += nt) {
                const int net = o / (R * H), q = o % (R * H), r = q / H, j = q % H;
                const float* x = l == 0 ? X + r * S : act[net] + ((size_t)(l - 1) * R + r) * H;
                const float* w = sWl(net, l) + j * (K + 1);
                float acc = sbl(net, l)[j];
                for (int k = 0; k < K; ++k) acc = fmaf(x[k], w[k], acc);
                act[net][((size_t)l * R + r) * H + j] = tanhf(acc);
            }
            __syncthreads();
        }
        for (int o = tid; o < R * (1 + A); o += nt) {
            const int r = o / (1 + A), c = o % (1 + A), net = c == 0 ? 0 : 1, oo = c == 0 ? 0 : c - 1;
            const float* x = act[net] + ((size_t)(L - 1) * R + r) * H;
            const float* w = sWhd(net) + oo * (H + 1);
            float acc = sbhd(net)[oo];
            for (int k = 0; k < H; ++k) acc = fmaf(x[k], w[k], acc);
            if (c == 0) outV[r] = acc; else outM[r * A + oo] = acc;
        }
        __syncthreads();
        // ---- PPO loss per row (ppo.py:29-49), same statements as ppo_loss_kernel ----
        for (int r = tid; r < R; r += nt) {
            const long long q = src[r];
            float* st = rstat + r * (2 + A);
            if (q < 0) {
                for (int c = 0; c < 2 + A; ++c) st[c] = 0.f;
                outV[r] = 0.f;
                for (int j = 0; j < A; ++j) outM[r * A + j] = 0.f;
                continue;
            }
            float* m = outM + r * A;
            const float value = outV[r];
            const float logp = row_logp(action + q * A, m, slog, A);
            const float ratio = expf(logp - old_logp[q]);
            const float ad = adv[q];
            const float surr1 = ratio * ad;
            const float surr2 = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip) * ad;
            const float w1 = surr1 < surr2 ? 1.f : (surr1 > surr2 ? 0.f : 0.5f);
            const float in_r = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? 1.f : 0.f;
            const float d_logp = -(ad * scale_actor) * (w1 + (1.f - w1) * in_r) * ratio;
            const float ov = old_value[q], rt = ret[q];
            const float dv = value - ov;
            const float vpc = ov + fminf(fmaxf(dv, -clip), clip);
            const float e1 = value - rt, e2 = vpc - rt;
            const float vl = e1 * e1, vlc = e2 * e2;
            const float wv = vl > vlc ? 1.f : (vl < vlc ? 0.f : 0.5f);
            const float in_v = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
            st[0] = fmaxf(vl, vlc);
            st[1] = fminf(surr1, surr2);
            outV[r] = scale_value * (wv * 2.f * e1 + (1.f - wv) * 2.f * e2 * in_v);
            for (int j = 0; j < A; ++j) {
                const float scale = expf(slog[j]);
                const float var = scale * scale;
                const float d = action[q * A + j] - m[j];
                st[2 + j] = d_logp * (d * d / var - 1.f);
                m[j] = d_logp * d / var;
            }
        }
        __syncthreads();
        for (int c = tid; c < 2 + A; c += nt) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc += rstat[r * (2 + A) + c];
            emit(mypart + (c < 2 ? fo.n_params + c : fo.off_logstd + c - 2), acc, first);
        }
        // ---- backward, one net after the other (dz buffers are shared) ----
        for (int net = 0; net < 2; ++net) {
            const int NO = net == 0 ? 1 : A;
            const float* dOut = net == 0 ? outV : outM;
            const float* hL = act[net] + (size_t)(L - 1) * R * H;
            for (int o = tid; o < NO * H + NO; o += nt) {
                float acc = 0.f;
                if (o < NO * H) {
                    const int oo = o / H, k = o % H;
                    for (int r = 0; r < R; ++r) acc = fmaf(dOut[r * NO + oo], hL[r * H + k], acc);
                    emit(mypart + fo.off_hw[net] + o, acc, first);
                } else {
                    const int oo = o - NO * H;
                    for (int r = 0; r < R; ++r) acc += dOut[r * NO + oo];
                    emit(mypart + fo.off_hb[net] + oo, acc, first);
                }
            }
            for (int o = tid; o < R * H; o += nt) {
                const int r = o / H, k = o % H;
                float acc = 0.f;
                for (int oo = 0; oo < NO; ++oo) acc = fmaf(dOut[r * NO + oo], sWhd(net)[oo * (H + 1) + k], acc);
                const float h = hL[o];
                dz0[o] = acc * (1.f - h * h);
            }
            __syncthreads();
            float *dz = dz0, *dzn = dz1;
            for (int l = L - 1; l >= 0; --l) {
                const int K = l == 0 ? S : H;
                const float* in = l == 0 ? X : act[net] + (size_t)(l - 1) * R * H;
                for (int o = tid; o < H * K + H; o += nt) {
                    float acc = 0.f;
                    if (o < H * K) {
                        const int j = o / K, k = o % K;
                        for (int r = 0; r < R; ++r) acc = fmaf(dz[r * H + j], in[r * K + k], acc);
                        emit(mypart + fo.off_w[net][l] + o, acc, first);
                    } else {
                        const int j = o - H * K;
                        for (int r = 0; r < R; ++r) acc += dz[r * H + j];
                        emit(mypart + fo.off_b[net][l] + j, acc, first);
                    }
                }
                if (l > 0) {
                    for (int o = tid; o < R * H; o += nt) {
                        const int r = o / H, k = o % H;
                        float acc = 0.f;
                        for (int j = 0; j < H; ++j) acc = fmaf(dz[r * H + j], sWl(net, l)[j * (H + 1) + k], acc);
                        const float h = in[o];
                        dzn[o] = acc * (1.f - h * h);
                    }
                }
                __syncthreads();
                float* t = dz; dz = dzn; dzn = t;
            }
        }
    }
}

// grad (+)= sum over the CTAs' partial rows; the entropy term of d logstd; losses like ppo_loss_final_kernel
__global__ void ppo_fused_reduce_kernel(const float* __restrict__ part, int n_cta, int part_stride, int n_params, int off_logstd,
                                        int A, double inv_div, float value_coef, float ent_coef, float ent_share, float grad_scale,
                                        const float* __restrict__ logstd, float* __restrict__ losses, float* __restrict__ grad,
                                        int accumulate) {
    // block (32 parameters, 8 row groups): group y adds the partial rows y, y+8, ...; the 8 sums are added in a fixed order
    __shared__ float sh[8][33];
    const int i = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (i < n_params && grad)
        for (int c = threadIdx.y; c < n_cta; c += 8) acc += part[(size_t)c * part_stride + i];
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && i < n_params && grad) {
        acc = sh[0][threadIdx.x];
        for (int y = 1; y < 8; ++y) acc += sh[y][threadIdx.x];
        if (i >= off_logstd && i < off_logstd + A) acc -= grad_scale * ent_coef * ent_share;
        grad[i] = accumulate ? grad[i] + acc : acc;
    }
    if (i == 0 && threadIdx.y == 1 && losses) {
        double v = 0.0, a = 0.0;
        for (int c = 0; c < n_cta; ++c) { v += part[(size_t)c * part_stride + n_params]; a += part[(size_t)c * part_stride + n_params + 1]; }
        const float vl = (float)(0.5 * v * inv_div), al = (float)(-a * inv_div);
        const float ent = row_entropy(logstd, A);
        losses[0] = vl * value_coef + al - ent * ent_coef * ent_share;
        losses[1] = vl;
        losses[2] = al;
        losses[3] = ent;
    }
}

// ---- fused acting / evaluation kernel (same residency rule as ppo_fused_kernel) -----------------------------------------
// get_action is called once per environment step with one row per environment (32 by default): one launch, 8 rows per
// CTA, the needed nets' weights read straight from the flat parameter vector into shared memory.
__global__ void __launch_bounds__(256) acnet_fused_eval_kernel(const float* __restrict__ params, FusedOffs fo,
                                                               const float* __restrict__ obs, const float* __restrict__ noise,
                                                               const float* __restrict__ action_in, float* __restrict__ value,
                                                               float* __restrict__ mean_out, float* __restrict__ action_out,
                                                               float* __restrict__ logp, float* __restrict__ ent, long long B,
                                                               int S, int A, int H, int L, int R, int net_mask, int mode) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int L0 = H * (S + 1) + H, LH = H * (H + 1) + H, T = L0 + (L - 1) * LH;
    auto sWl = [&](int net, int l) { return sm + net * T + (l == 0 ? 0 : L0 + (l - 1) * LH); };
    auto sbl = [&](int net, int l) { return sWl(net, l) + H * ((l == 0 ? S : H) + 1); };
    auto sWhd = [&](int net) { return sm + 2 * T + (net == 0 ? 0 : (H + 1) + 1); };
    auto sbhd = [&](int net) { return sWhd(net) + (net == 0 ? 1 : A) * (H + 1); };
    float* p = sm + 2 * T + (1 + A) * (H + 1) + 1 + A;
    float* slog = p; p += A;
    float* X = p; p += R * S;
    float* hbuf[2] = {p, p + 2 * R * H}; p += 4 * R * H;   // ping-pong activations, both nets side by side
    float* outV = p; p += R;
    float* outM = p; p += R * A;
    for (int net = 0; net < 2; ++net) {
        if (!(net_mask >> net & 1)) continue;
        for (int l = 0; l < L; ++l) {
            const int K = l == 0 ? S : H;
            const float* g = params + fo.off_w[net][l];
            for (int i = tid; i < H * K; i += nt) sWl(net, l)[(i / K) * (K + 1) + i % K] = g[i];
            for (int i = tid; i < H; i += nt) sbl(net, l)[i] = params[fo.off_b[net][l] + i];
        }
        const int NO = net == 0 ? 1 : A;
        for (int i = tid; i < NO * H; i += nt) sWhd(net)[(i / H) * (H + 1) + i % H] = params[fo.off_hw[net] + i];
        for (int i = tid; i < NO; i += nt) sbhd(net)[i] = params[fo.off_hb[net] + i];
    }
    for (int i = tid; i < A; i += nt) slog[i] = params[fo.off_logstd + i];
    const long long n_tiles = (B + R - 1) / R;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * R;
        __syncthreads();
        for (int i = tid; i < R * S; i += nt) { const long long q = row0 + i / S; X[i] = q < B ? obs[q * S + i % S] : 0.f; }
        __syncthreads();
        for (int l = 0; l < L; ++l) {
            const float* in = hbuf[(l + 1) & 1];
            float* out = hbuf[l & 1];
            const int K = l == 0 ? S : H;
            for (int o = tid; o < 2 * R * H; o += nt) {
                const int net = o / (R * H), q = o % (R * H), r = q / H, j = q % H;
                if (!(net_mask >> net & 1)) continue;
                const float* x = l == 0 ? X + r * S : in + net * R * H + r * H;
                const float* w = sWl(net, l) + j * (K + 1);
                float acc = sbl(net, l)[j];
                for (int k = 0; k < K; ++k) acc = fmaf(x[k], w[k], acc);
                out[o] = tanhf(acc);
            }
            __syncthreads();
        }
        const float* hL = hbuf[(L - 1) & 1];
        for (int o = tid; o < R * (1 + A); o += nt) {
            const int r = o / (1 + A), c = o % (1 + A), net = c == 0 ? 0 : 1, oo = c == 0 ? 0 : c - 1;
            if (!(net_mask >> net & 1)) continue;
            const float* x = hL + net * R * H + r * H;
            const float* w = sWhd(net) + oo * (H + 1);
            float acc = sbhd(net)[oo];
            for (int k = 0; k < H; ++k) acc = fmaf(x[k], w[k], acc);
            if (c == 0) outV[r] = acc; else outM[r * A + oo] = acc;
        }
        __syncthreads();
        for (int r = tid; r < R; r += nt) {
            const long long q = row0 + r;
            if (q >= B) continue;
            const float* m = outM + r * A;
            if (value) value[q] = outV[r];
            if (mean_out)
                for (int j = 0; j < A; ++j) mean_out[q * A + j] = m[j];
            if (mode == 1) {
                float acc = 0.f;
                for (int j = 0; j < A; ++j) {
                    const float scale = expf(slog[j]);
                    const float x = noise ? __fadd_rn(m[j], __fmul_rn(scale, noise[q * A + j])) : m[j];
                    action_out[q * A + j] = x;
                    const float d = x - m[j];
                    acc += -(d * d) / (2.f * (scale * scale)) - logf(scale) - kLogSqrt2Pi;
                }
                if (logp) logp[q] = acc;
            } else if (mode == 2) {
                if (logp) logp[q] = row_logp(action_in + q * A, m, slog, A);
            }
            if (ent) ent[q] = row_entropy(slog, A);
        }
    }
}

struct FusedPlan {
    bool ok;
    int R, n_cta;
    size_t smem;
};
static FusedPlan plan_fused(const frl_acnet* n, long long B) {
    const int H = n->cfg.hidden, S = n->cfg.obs_dim, A = n->cfg.a_dim, L = n->cfg.layers;
    size_t w = 0;
    for (int l = 0; l < L; ++l) w += (size_t)H * ((l == 0 ? S : H) + 1) + H;
    w = 2 * w + (size_t)(1 + A) * (H + 1) + 1 + A + A;
    FusedPlan best{false, 0, 0, 0};
    const int cand[3] = {32, 16, 8};
    for (int R : cand) {
        // small minibatches: 8-row tiles spread the rows over the SMs
        if (R > 8 && (B + 7) / 8 <= 296) continue;
        const size_t acts = (size_t)R * S + 2 * (size_t)L * R * H + 2 * (size_t)R * H + R + (size_t)R * A + (size_t)R * (2 + A);
        const size_t bytes = (w + acts + 2) * sizeof(float) + (size_t)R * sizeof(long long);
        if (bytes > 200 * 1024) continue;
        best = FusedPlan{true, R, (int)std::min<long long>((B + R - 1) / R, 296), bytes};
        break;
    }
    return best;
}

}  // namespace frl

using namespace frl;

extern "C" int frl_acnet_create(const frl_acnet_config* cfg, frl_acnet** out) {
    FRL_REQUIRE(cfg && out, "null config/out");
    FRL_REQUIRE(cfg->obs_dim >= 1 && cfg->obs_dim <= 4096 && cfg->a_dim >= 1 && cfg->a_dim <= 64, "bad obs_dim / a_dim");
    FRL_REQUIRE(cfg->hidden >= 16 && cfg->hidden % 16 == 0 && cfg->hidden <= 1024, "hidden must be a multiple of 16, <= 1024");
    FRL_REQUIRE(cfg->layers >= 1 && cfg->layers <= kAcMaxLayers, "layers must be in 1..8");
    int ndev = 0;
    FRL_CUDA(cudaGetDeviceCount(&ndev));
    FRL_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
    FRL_CUDA(cudaSetDevice(cfg->device));
    frl_acnet* n = new (std::nothrow) frl_acnet();
    if (!n) { set_error("out of host memory"); return FRL_ERR_NOMEM; }
    n->cfg = *cfg;
    const int H = cfg->hidden, S = cfg->obs_dim, A = cfg->a_dim, L = cfg->layers;
    n->Sp = round_up(S, 16);
    n->Aop = round_up(A, 16);
    long long off = 0;
    auto adv = [&](long long k) { long long o = off; off += k; return o; };
    // state_dict order: critic.net.*, critic_head.*, actor.net.*, dist.logstd, dist.fc_mean.*
    for (int l = 0; l < L; ++l) { n->off_w[0][l] = adv((long long)H * (l == 0 ? S : H)); n->off_b[0][l] = adv(H); }
    n->off_hw[0] = adv(H);
    n->off_hb[0] = adv(1);
    for (int l = 0; l < L; ++l) { n->off_w[1][l] = adv((long long)H * (l == 0 ? S : H)); n->off_b[1][l] = adv(H); }
    n->off_logstd = adv(A);
    n->off_hw[1] = adv((long long)A * H);
    n->off_hb[1] = adv(A);
    n->n_params = off;
    size_t f = 0;
    auto take = [&](size_t x) { size_t o = f; f += (x + 3) / 4 * 4; return o; };
    for (int net = 0; net < 2; ++net) {
        for (int l = 0; l < L; ++l) {
            const size_t Kp = l == 0 ? n->Sp : H;
            n->o_Wt[net][l] = take(Kp * H);
            n->o_W[net][l] = take((size_t)H * Kp);
            n->o_b[net][l] = take(H);
        }
        n->o_Wht[net] = take((size_t)H * n->Aop);
        n->o_Wh[net] = take((size_t)n->Aop * H);
        n->o_bh[net] = take(n->Aop);
    }
    n->o_logstd = take(n->Aop);
    n->pack_floats = f;
    cudaError_t e = cudaMalloc(&n->pack, f * sizeof(float));
    if (e != cudaSuccess) { delete n; return cuda_fail(e, "cudaMalloc(acnet pack)", __FILE__, __LINE__); }
    *out = n;
    return FRL_OK;
}

extern "C" int frl_acnet_destroy(frl_acnet* n) {
    if (!n) return FRL_OK;
    cudaSetDevice(n->cfg.device);
    if (n->pack) cudaFree(n->pack);
    if (n->ws) cudaFree(n->ws);
    for (float* r : n->retired) cudaFree(r);
    delete n;
    return FRL_OK;
}

extern "C" long long frl_acnet_num_params(const frl_acnet* n) { return n ? n->n_params : -1; }

extern "C" int frl_acnet_set_params(frl_acnet* n, const float* params_dev, void* /*stream*/) {
    FRL_REQUIRE(n && params_dev, "null net/params");
    n->params_dev = params_dev;
    n->version += 1;
    return FRL_OK;
}

static int acnet_eval_common(frl_acnet* n, const float* obs, const float* noise, const float* action_in, long long B,
                             float* value, float* mean, float* action_out, float* logp, float* ent, int mode, void* stream) {
    FRL_REQUIRE(n && n->version > 0, "frl_acnet_set_params has not been called");
    if (B <= 0) return FRL_OK;
    FRL_REQUIRE(obs, "null observations");
    FRL_REQUIRE(B < (1ll << 31) / std::max(n->cfg.hidden, n->Sp), "batch too large");
    FRL_CUDA(cudaSetDevice(n->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool need_actor = mean || action_out || logp;
    {   // nets that fit in shared memory: one launch (FRL_PPO_GEMM=1 forces the per-layer GEMM path)
        const int H = n->cfg.hidden, S = n->cfg.obs_dim, A = n->cfg.a_dim, L = n->cfg.layers, R = 8;
        size_t w = 0;
        for (int l = 0; l < L; ++l) w += (size_t)H * ((l == 0 ? S : H) + 1) + H;
        w = 2 * w + (size_t)(1 + A) * (H + 1) + 1 + A + A;
        const size_t smem = (w + (size_t)R * S + 4 * (size_t)R * H + R + (size_t)R * A) * sizeof(float);
        if (smem <= 200 * 1024 && !getenv("FRL_PPO_GEMM")) {
            static bool attr_done[64] = {};
            if (!attr_done[n->cfg.device & 63]) {
                FRL_CUDA(cudaFuncSetAttribute(acnet_fused_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_done[n->cfg.device & 63] = true;
            }
            FusedOffs fo;
            for (int net = 0; net < 2; ++net) {
                for (int l = 0; l < L; ++l) { fo.off_w[net][l] = (int)n->off_w[net][l]; fo.off_b[net][l] = (int)n->off_b[net][l]; }
                fo.off_hw[net] = (int)n->off_hw[net];
                fo.off_hb[net] = (int)n->off_hb[net];
            }
            fo.off_logstd = (int)n->off_logstd;
            fo.n_params = (int)n->n_params;
            const int grid = (int)std::min<long long>((B + R - 1) / R, 592);
            acnet_fused_eval_kernel<<<grid, 256, smem, st>>>(n->params_dev, fo, obs, noise, action_in, value, mean, action_out, logp,
                                                            ent, B, S, A, H, L, R, (value ? 1 : 0) | (need_actor ? 2 : 0), mode);
            FRL_CUDA(cudaGetLastError());
            return FRL_OK;
        }
    }
    ABuf b;
    size_t used = 0;
    int rc = ac_prepare(n, obs, nullptr, B, 0, &b, &used, st);
    if (rc != FRL_OK) return rc;
    rc = ac_forward(n, b, (int)B, (value ? 1 : 0) | (need_actor ? 2 : 0), st);
    if (rc != FRL_OK) return rc;
    acnet_head_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(b.out[0], b.out[1], n->Aop, n->pack + n->o_logstd, noise, action_in,
                                                                 value, mean, action_out, logp, ent, B, n->cfg.a_dim, mode);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_acnet_forward(frl_acnet* n, const float* obs, long long B, float* value, float* mean, void* stream) {
    return acnet_eval_common(n, obs, nullptr, nullptr, B, value, mean, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int frl_acnet_act(frl_acnet* n, const float* obs, const float* noise, long long B, float* value, float* action,
                             float* logp, float* ent, void* stream) {
    FRL_REQUIRE(B <= 0 || action, "null action output");
    return acnet_eval_common(n, obs, noise, nullptr, B, value, nullptr, action, logp, ent, 1, stream);
}

extern "C" int frl_acnet_evaluate(frl_acnet* n, const float* obs, const float* action, long long B, float* value, float* logp,
                                  float* ent, void* stream) {
    FRL_REQUIRE(B <= 0 || (action && logp), "null action / log-prob");
    return acnet_eval_common(n, obs, nullptr, action, B, value, nullptr, nullptr, logp, ent, 2, stream);
}

extern "C" int frl_ppo_loss_grad(frl_acnet* n, const float* obs, const float* action, const float* old_logp, const float* adv,
                                 const float* ret, const float* old_value, const long long* idx, long long B,
                                 long long mean_divisor, float clip_param, float value_loss_coef, float entropy_coef,
                                 float grad_scale, int accumulate, float* losses, float* grad, void* stream) {
    FRL_REQUIRE(n && n->version > 0, "frl_acnet_set_params has not been called");
    FRL_REQUIRE(obs && action && old_logp && adv && ret && old_value, "null input");
    FRL_REQUIRE(B > 0 && B < (1ll << 31) / std::max(n->cfg.hidden, n->Sp), "batch size out of range");
    FRL_REQUIRE(clip_param >= 0.f, "clip_param must be >= 0");
    FRL_CUDA(cudaSetDevice(n->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int H = n->cfg.hidden, S = n->cfg.obs_dim, A = n->cfg.a_dim, L = n->cfg.layers, Aop = n->Aop;
    const int R = (int)B;
    const double div = (double)(mean_divisor > 0 ? mean_divisor : B);

    // nets that fit in shared memory: one fused kernel + a reduction (FRL_PPO_GEMM=1 forces the per-layer GEMM path)
    const FusedPlan fp = plan_fused(n, B);
    if (fp.ok && !getenv("FRL_PPO_GEMM")) {
        const int stride = round_up((int)n->n_params + 2, 4);
        int rc = acnet_ws(n, (size_t)fp.n_cta * stride);
        if (rc != FRL_OK) return rc;
        static bool attr_done[64] = {};
        if (!attr_done[n->cfg.device & 63]) {
            FRL_CUDA(cudaFuncSetAttribute(ppo_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_done[n->cfg.device & 63] = true;
        }
        FusedOffs fo;
        for (int net = 0; net < 2; ++net) {
            for (int l = 0; l < L; ++l) { fo.off_w[net][l] = (int)n->off_w[net][l]; fo.off_b[net][l] = (int)n->off_b[net][l]; }
            fo.off_hw[net] = (int)n->off_hw[net];
            fo.off_hb[net] = (int)n->off_hb[net];
        }
        fo.off_logstd = (int)n->off_logstd;
        fo.n_params = (int)n->n_params;
        ppo_fused_kernel<<<fp.n_cta, 512, fp.smem, st>>>(n->params_dev, fo, obs, action, old_logp, adv, ret, old_value, idx, B, S, A, H,
                                                        L, fp.R, clip_param, (float)((double)grad_scale / div),
                                                        (float)((double)grad_scale * 0.5 * value_loss_coef / div), n->ws, stride);
        FRL_CUDA(cudaGetLastError());
        ppo_fused_reduce_kernel<<<((int)n->n_params + 31) / 32, dim3(32, 8), 0, st>>>(
            n->ws, fp.n_cta, stride, (int)n->n_params, (int)n->off_logstd, A, 1.0 / div, value_loss_coef, entropy_coef,
            (float)((double)B / div), grad_scale, n->params_dev + n->off_logstd, losses, grad, accumulate);
        FRL_CUDA(cudaGetLastError());
        return FRL_OK;
    }

    auto plan_splits = [&](int tiles) {
        int want = std::max(1, 296 / std::max(1, tiles));
        int maxs = std::max(1, (R + 63) / 64);
        return std::min(want, maxs);
    };
    const int sp_hh = plan_splits(((H + 127) / 128) * ((std::max(H, n->Sp) + 127) / 128));
    const int sp_small = plan_splits(2);
    const int max_sp = std::max(sp_hh, sp_small);
    const int cs_chunks = std::max(1, std::min(148, (R + 255) / 256));
    const int cs_rows = (R + cs_chunks - 1) / cs_chunks;
    const int loss_blocks = (R + 255) / 256;
    const int Wp = 2 + A;

    size_t used = 0;
    ac_carve(n, nullptr, (size_t)R, &used);
    size_t off = used;
    auto take = [&](size_t f) { size_t o = off; off += (f + 3) / 4 * 4; return o; };
    const size_t o_dz0 = take((size_t)R * H), o_dz1 = take((size_t)R * H);
    const size_t o_part = take((size_t)max_sp * H * std::max(H, n->Sp));
    const size_t o_cs = take((size_t)cs_chunks * H);
    const size_t o_lp = take((size_t)loss_blocks * Wp * 2 + 4);
    ABuf b;
    int rc = ac_prepare(n, obs, idx, B, off - used, &b, &used, st);
    if (rc != FRL_OK) return rc;
    float* ws = n->ws;
    float *dZ[2] = {ws + o_dz0, ws + o_dz1}, *part = ws + o_part, *cs = ws + o_cs;
    double* lpart = reinterpret_cast<double*>(ws + o_lp);   // offsets are multiples of 4 floats
    const float* pk = n->pack;

    rc = ac_forward(n, b, R, 3, st);
    if (rc != FRL_OK) return rc;
    ppo_loss_kernel<<<loss_blocks, 256, (size_t)8 * Wp * sizeof(double), st>>>(
        b.out[0], b.out[1], Aop, pk + n->o_logstd, action, old_logp, adv, ret, old_value, idx, B, A, clip_param,
        (float)((double)grad_scale / div), (float)((double)grad_scale * 0.5 * value_loss_coef / div), lpart);
    FRL_CUDA(cudaGetLastError());
    if (losses || grad) {
        ppo_loss_final_kernel<<<1, round_up(Wp, 32), 0, st>>>(lpart, loss_blocks, A, 1.0 / div, value_loss_coef, entropy_coef,
                                                              (float)((double)B / div), grad_scale, pk + n->o_logstd, losses,
                                                              grad ? grad + n->off_logstd : nullptr, accumulate);
        FRL_CUDA(cudaGetLastError());
    }
    if (!grad) return FRL_OK;

    auto ksplit = [&](int splits) { return round_up((R + splits - 1) / splits, BK); };
    auto wgrad = [&](const float* dZm, int ldz, int rows, const float* X, int ldx, int Kp, int K, long long goff, int sp) -> int {
        int r = run_gemm<true, EPI_NONE>(dZm, ldz, X, ldx, part, Kp, rows, Kp, R, nullptr, nullptr, 0, sp, ksplit(sp), st);
        if (r != FRL_OK) return r;
        reduce_wgrad_kernel<<<(rows * K + 255) / 256, 256, 0, st>>>(part, sp, (size_t)rows * Kp, Kp, rows, K, grad + goff, accumulate);
        FRL_CUDA(cudaGetLastError());
        return FRL_OK;
    };
    auto bgrad = [&](const float* dZm, int ldz, int cols, long long goff) -> int {
        colsum_partial_kernel<<<dim3((cols + 127) / 128, cs_chunks), 128, 0, st>>>(dZm, ldz, R, cols, cs_rows, cs);
        reduce_bias_kernel<<<(cols + 127) / 128, 128, 0, st>>>(cs, cs_chunks, cols, 0, cols, grad + goff, accumulate);
        FRL_CUDA(cudaGetLastError());
        return FRL_OK;
    };
    for (int net = 0; net < 2; ++net) {
        const int NO = net == 0 ? 1 : A;
        const float* hL = b.act[net] + (size_t)(L - 1) * R * H;
        rc = wgrad(b.out[net], Aop, NO, hL, H, H, H, n->off_hw[net], sp_small);           // critic_head / dist.fc_mean
        if (rc != FRL_OK) return rc;
        rc = bgrad(b.out[net], Aop, NO, n->off_hb[net]);
        if (rc != FRL_OK) return rc;
        rc = run_gemm<false, EPI_DTANH>(b.out[net], Aop, pk + n->o_Wh[net], H, dZ[0], H, R, H, Aop, nullptr, hL, H, 1, Aop, st);
        if (rc != FRL_OK) return rc;
        int cur = 0;
        for (int l = L - 1; l >= 0; --l) {
            const float* X = l == 0 ? b.Xp : b.act[net] + (size_t)(l - 1) * R * H;
            const int Kp = l == 0 ? n->Sp : H, K = l == 0 ? S : H;
            rc = wgrad(dZ[cur], H, H, X, Kp, Kp, K, n->off_w[net][l], l == 0 ? sp_small : sp_hh);
            if (rc != FRL_OK) return rc;
            rc = bgrad(dZ[cur], H, H, n->off_b[net][l]);
            if (rc != FRL_OK) return rc;
            if (l > 0) {
                rc = run_gemm<false, EPI_DTANH>(dZ[cur], H, pk + n->o_W[net][l], H, dZ[cur ^ 1], H, R, H, H, nullptr, X, H, 1, H, st);
                if (rc != FRL_OK) return rc;
                cur ^= 1;
            }
        }
    }
    return FRL_OK;
}
