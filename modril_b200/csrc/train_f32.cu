// FP32 conditional-flow-matching loss + gradient (forward, backward) for one role batch.
// Replaces  fm_loss(...) ; loss.backward()  of the reference (modril/flowril/pretrain.py:179-190, :94-109;
// modril/flowril/flowril.py:317-322) with hand-written SGEMM-class kernels:
//   forward  : X0 = [a_t | s | t],  H_l = relu(H_{l-1} W_l^T + b_l),  V = H_L Wh^T + bh
//   loss     : L = (1/Bdiv) sum_b (1-t_b)^1.5 sum_j (v_bj - (noise - a0)_bj)^2 ;  dV in place of V
//   backward : dWh = dV^T H_L, dbh = colsum dV, dH_L = dV Wh ; per layer dZ = dH * (H>0),
//              dW_l = dZ^T H_{l-1} (split over the batch, partials reduced in fixed order), db_l = colsum dZ,
//              dH_{l-1} = dZ W_l
// Everything is deterministic: no float atomics, fixed split counts for a given (B, net).
#include <algorithm>
#include <cmath>

#include "frl_internal.h"
#include "sgemm.cuh"

namespace frl {

// X0[row] = [ (1-t) a0 + t noise | s | t | 0 ... ]   (pretrain.py:184 / :102; column order networks.py:24)
__global__ void cfm_prepare_kernel(const float* __restrict__ s, const float* __restrict__ a0,
                                   const float* __restrict__ t, const float* __restrict__ noise,
                                   float* __restrict__ X0, long long B, int s_dim, int a_dim, int ld) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B * ld) return;
    const long long row = i / ld;
    const int k = (int)(i % ld);
    float v = 0.f;
    const float tt = t[row];
    if (k < a_dim) {
        v = __fadd_rn(__fmul_rn(1.f - tt, a0[row * a_dim + k]), __fmul_rn(tt, noise[row * a_dim + k]));
    } else if (k < a_dim + s_dim) {
        v = s[row * s_dim + (k - a_dim)];
    } else if (k == a_dim + s_dim) {
        v = tt;
    }
    X0[i] = v;
}

// Per row: loss term and dL/d(head pre-activations), written over V in place.
__global__ void __launch_bounds__(256) cfm_loss_kernel(float* __restrict__ V, int ldv, const float* __restrict__ a0,
                                                       const float* __restrict__ noise, const float* __restrict__ t,
                                                       long long B, int a_dim, int n_heads, float sign, float dscale,
                                                       double* __restrict__ partial) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double term = 0.0;
    if (row < B) {
        const float w = powf(1.f - t[row], 1.5f);  // pretrain.py:188 / :105
        float* v = V + row * ldv;
        float acc = 0.f;
        for (int j = 0; j < a_dim; ++j) {
            float pred = v[j];
            float th = 0.f;
            if (n_heads == 2) {
                th = tanhf(v[a_dim + j]);
                pred += sign * th;
            }
            const float diff = pred - (noise[row * a_dim + j] - a0[row * a_dim + j]);
            acc += diff * diff;
            const float d = dscale * w * diff;  // dscale = grad_scale * 2 / Bdiv
            v[j] = d;
            if (n_heads == 2) v[a_dim + j] = d * sign * (1.f - th * th);
        }
        term = (double)(w * acc);
    }
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int i = 0; i < 8; ++i) tsum += sh[i];
        partial[blockIdx.x] = tsum;
    }
}

__global__ void loss_final_kernel(const double* __restrict__ partial, int n, double inv_div, float* __restrict__ loss) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(sh[0] * inv_div);
}

static int ensure_ws(frl_net* net, size_t floats) {
    if (net->ws_floats >= floats) return FRL_OK;
    if (net->ws) net->retired.push_back(net->ws);   // kept alive: see frl_net::retired
    net->ws = nullptr;
    net->ws_floats = 0;
    FRL_CUDA(cudaMalloc(&net->ws, floats * sizeof(float)));
    net->ws_floats = floats;
    return FRL_OK;
}

int cfm_loss_grad_f32(frl_net* net, float role_sign, const float* s, const float* a0, const float* t,
                      const float* noise, long long B, long long mean_divisor, float grad_scale, int accumulate,
                      float* loss, float* grad, cudaStream_t st) {
    const int H = net->cfg.hidden, A = net->cfg.a_dim, depth = net->cfg.depth;
    const int K0p = net->d_in_p, NOp = net->n_out_p;
    FRL_REQUIRE(B > 0 && B < (1ll << 31) / std::max(H, K0p), "batch size out of range");
    {
        int rcp = ensure_pack(net, PACK_F32, st);
        if (rcp != FRL_OK) return rcp;
    }
    const int Bi = (int)B;
    const double div = (double)(mean_divisor > 0 ? mean_divisor : B);

    // split plan for the batch-reduction GEMMs (deterministic for a given B)
    const int k_chunk_min = 64;
    auto plan_splits = [&](int tiles) {
        int want = std::max(1, 296 / std::max(1, tiles));
        int maxs = std::max(1, (Bi + k_chunk_min - 1) / k_chunk_min);
        return std::min(want, maxs);
    };
    const int splits_hh = plan_splits(((H + 127) / 128) * ((H + 127) / 128));
    const int splits_small = plan_splits(2);
    const int max_splits = std::max(splits_hh, splits_small);
    const int cs_chunks = std::max(1, std::min(148, (Bi + 255) / 256));
    const int cs_rows = (Bi + cs_chunks - 1) / cs_chunks;
    const int loss_blocks = (Bi + 255) / 256;

    // workspace carve-up (floats)
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) / 4 * 4; return o; };
    const size_t o_x0 = take((size_t)B * K0p);
    size_t o_h[kMaxDepth];
    for (int l = 0; l < depth; ++l) o_h[l] = take((size_t)B * H);
    const size_t o_v = take((size_t)B * NOp);
    const size_t o_dz0 = take((size_t)B * H);
    const size_t o_dz1 = take((size_t)B * H);
    const size_t o_part = take((size_t)max_splits * H * std::max(H, std::max(K0p, NOp)));
    const size_t o_cs = take((size_t)cs_chunks * std::max(H, NOp));
    const size_t o_lp = take((size_t)loss_blocks * 2 + 4);  // doubles
    int rc = ensure_ws(net, off);
    if (rc != FRL_OK) return rc;
    float* ws = net->ws;
    float* X0 = ws + o_x0;
    float* V = ws + o_v;
    float* dZ[2] = {ws + o_dz0, ws + o_dz1};
    float* part = ws + o_part;
    float* cs = ws + o_cs;
    double* lpart = reinterpret_cast<double*>(ws + o_lp);
    const NetF32& w = net->f32;

    // ---- forward ----
    {
        long long n = B * K0p;
        cfm_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, a0, t, noise, X0, B, net->cfg.s_dim, A, K0p);
        FRL_CUDA(cudaGetLastError());
    }
    const float* in = X0;
    int Kin = K0p;
    for (int l = 0; l < depth; ++l) {
        float* out = ws + o_h[l];
        rc = run_gemm<false, EPI_BIAS_RELU>(in, Kin, w.Wt[l], H, out, H, Bi, H, Kin, w.b[l], nullptr, 0, 1, Kin, st);
        if (rc != FRL_OK) return rc;
        in = out;
        Kin = H;
    }
    const float* Hlast = in;
    rc = run_gemm<false, EPI_BIAS>(Hlast, H, w.Wht, NOp, V, NOp, Bi, NOp, H, w.bh, nullptr, 0, 1, H, st);
    if (rc != FRL_OK) return rc;

    // ---- loss + dV (in place) ----
    cfm_loss_kernel<<<loss_blocks, 256, 0, st>>>(V, NOp, a0, noise, t, B, A, net->cfg.n_heads, role_sign,
                                                 (float)((double)grad_scale * 2.0 / div), lpart);
    FRL_CUDA(cudaGetLastError());
    if (loss) {
        loss_final_kernel<<<1, 256, 0, st>>>(lpart, loss_blocks, 1.0 / div, loss);
        FRL_CUDA(cudaGetLastError());
    }
    if (!grad) return FRL_OK;

    auto ksplit = [&](int splits) { return round_up((Bi + splits - 1) / splits, BK); };

    // ---- heads backward ----
    {
        // dWh [NOp][H] = dV^T [NOp x B] * Hlast [B][H]
        const int sp = splits_small;
        rc = run_gemm<true, EPI_NONE>(V, NOp, Hlast, H, part, H, NOp, H, Bi, nullptr, nullptr, 0, sp, ksplit(sp), st);
        if (rc != FRL_OK) return rc;
        for (int h = 0; h < net->cfg.n_heads; ++h) {
            // partial layout [split][NOp][H]; head h's rows start at h*A
            reduce_wgrad_kernel<<<(A * H + 255) / 256, 256, 0, st>>>(part + (size_t)h * A * H, sp, (size_t)NOp * H, H,
                                                                     A, H, grad + net->off_wh[h], accumulate);
        }
        FRL_CUDA(cudaGetLastError());
        colsum_partial_kernel<<<dim3((NOp + 63) / 64, cs_chunks), 64, 0, st>>>(V, NOp, B, NOp, cs_rows, cs);
        for (int h = 0; h < net->cfg.n_heads; ++h)
            reduce_bias_kernel<<<1, 64, 0, st>>>(cs, cs_chunks, NOp, h * A, A, grad + net->off_bh[h], accumulate);
        FRL_CUDA(cudaGetLastError());
        // dZ_L = (dV [B][NOp] * Wh [NOp][H]) masked by H_L > 0
        rc = run_gemm<false, EPI_MASK>(V, NOp, w.Wh, H, dZ[0], H, Bi, H, NOp, nullptr, Hlast, H, 1, NOp, st);
        if (rc != FRL_OK) return rc;
    }
    // ---- trunk backward ----
    int cur = 0;
    for (int l = depth - 1; l >= 0; --l) {
        const float* Xin = l == 0 ? X0 : ws + o_h[l - 1];
        const int Kp = l == 0 ? K0p : H;
        const int K = l == 0 ? net->d_in : H;
        const int sp = (Kp == H) ? splits_hh : splits_small;
        // dW_l [H][Kp] = dZ^T [H x B] * Xin [B][Kp]
        rc = run_gemm<true, EPI_NONE>(dZ[cur], H, Xin, Kp, part, Kp, H, Kp, Bi, nullptr, nullptr, 0, sp, ksplit(sp), st);
        if (rc != FRL_OK) return rc;
        reduce_wgrad_kernel<<<(H * K + 255) / 256, 256, 0, st>>>(part, sp, (size_t)H * Kp, Kp, H, K,
                                                                 grad + net->off_w[l], accumulate);
        colsum_partial_kernel<<<dim3((H + 127) / 128, cs_chunks), 128, 0, st>>>(dZ[cur], H, B, H, cs_rows, cs);
        reduce_bias_kernel<<<(H + 127) / 128, 128, 0, st>>>(cs, cs_chunks, H, 0, H, grad + net->off_b[l], accumulate);
        FRL_CUDA(cudaGetLastError());
        if (l > 0) {
            // dZ_{l-1} = (dZ_l [B][H] * W_l [H][H]) masked by H_{l-1} > 0
            rc = run_gemm<false, EPI_MASK>(dZ[cur], H, w.W[l], H, dZ[cur ^ 1], H, Bi, H, H, nullptr, Xin, H, 1, H, st);
            if (rc != FRL_OK) return rc;
            cur ^= 1;
        }
    }
    return FRL_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Anti-divergence and Jacobian-stability regularisers of the scrf reward model and their gradients
// (modril/flowril/flowril.py:255-264;  _hutch_div pretrain.py:163-177;  _jacobian_frobenius pretrain.py:55-59).
//
//   loss_anti   = mean_b ( v^T J_r v )^2,           J_r = d tanh(head_r)/d a
//   loss_stable = mean_b || J_{v_c + r}^T v ||^2
//
// The reference gets the parameter gradients from autograd's double backward.  Here they are written out: ReLU has
// no second derivative, so every pass is a masked chain through the same weights, and each loss needs
//   anti  : tangent forward u (with the primal), one backward over the STACKED rows [primal-adjoint ; tangent-adjoint]
//   stable: reverse chain nu, a forward tangent pass mubar seeded with 2g/N, one primal backward; the weight
//           gradients are again stacked products  [dz ; nu]^T [h ; mubar]
// Rows 0..N-1 of every stacked buffer belong to the primal, rows N..2N-1 to the tangent-like quantity; both halves
// share the primal's ReLU masks (sgemm `mask_rows`).
// ---------------------------------------------------------------------------------------------------------------
__global__ void reg_prepare_kernel(const float* __restrict__ s, const float* __restrict__ a,
                                   const float* __restrict__ t, const float* __restrict__ probe,
                                   float* __restrict__ X0, long long N, int s_dim, int a_dim, int ld) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= 2 * N * ld) return;
    const long long row = i / ld;
    const int k = (int)(i % ld);
    float v = 0.f;
    if (row < N) {
        if (k < a_dim) v = a[row * a_dim + k];
        else if (k < a_dim + s_dim) v = s[row * s_dim + (k - a_dim)];
        else if (k == a_dim + s_dim) v = t[row];
    } else if (k < a_dim) {
        v = probe[(row - N) * a_dim + k];
    }
    X0[i] = v;
}

__device__ __forceinline__ void block_partial(double term, double* __restrict__ partial) {
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int i = 0; i < 8; ++i) tsum += sh[i];
        partial[blockIdx.x] = tsum;
    }
}

// anti: q = sum_j v_j tau_j w_j;  DV top = [0 | dq v w dtau/drho], DV bottom = [0 | dq v tau],  dq = 2 q / Ndiv
__global__ void __launch_bounds__(256) reg_anti_heads_kernel(const float* __restrict__ HV, const float* __restrict__ TV,
                                                             const float* __restrict__ probe, float* __restrict__ DV,
                                                             long long N, int a_dim, int ld, float dscale,
                                                             double* __restrict__ partial) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double term = 0.0;
    if (row < N) {
        float q = 0.f;
        for (int j = 0; j < a_dim; ++j) {
            const float r = tanhf(HV[row * ld + a_dim + j]);
            q = fmaf(probe[row * a_dim + j] * (1.f - r * r), TV[row * ld + a_dim + j], q);
        }
        const float dq = dscale * q;
        float* top = DV + row * ld;
        float* bot = DV + (N + row) * ld;
        for (int j = 0; j < ld; ++j) { top[j] = 0.f; bot[j] = 0.f; }
        for (int j = 0; j < a_dim; ++j) {
            const float r = tanhf(HV[row * ld + a_dim + j]);
            const float tau = 1.f - r * r;
            const float v = probe[row * a_dim + j];
            top[a_dim + j] = dq * v * TV[row * ld + a_dim + j] * (-2.f * r * tau);
            bot[a_dim + j] = dq * v * tau;
        }
        term = (double)q * (double)q;
    }
    block_partial(term, partial);
}

// stable, start of the reverse chain: DV bottom = [v | v tau]
__global__ void reg_stable_start_kernel(const float* __restrict__ HV, const float* __restrict__ probe,
                                        float* __restrict__ DV, long long N, int a_dim, int ld) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= N) return;
    float* bot = DV + (N + row) * ld;
    for (int j = 0; j < ld; ++j) bot[j] = 0.f;
    for (int j = 0; j < a_dim; ++j) {
        const float r = tanhf(HV[row * ld + a_dim + j]);
        const float v = probe[row * a_dim + j];
        bot[j] = v;
        bot[a_dim + j] = v * (1.f - r * r);
    }
}

// stable: p = |g|^2 over the action columns of G; seed of the forward tangent pass X0 bottom = [2 g / Ndiv | 0]
__global__ void __launch_bounds__(256) reg_stable_g_kernel(const float* __restrict__ G, float* __restrict__ X0,
                                                           long long N, int a_dim, int ld, float dscale,
                                                           double* __restrict__ partial) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double term = 0.0;
    if (row < N) {
        float acc = 0.f;
        float* bot = X0 + (N + row) * ld;
        for (int j = 0; j < ld; ++j) {
            float g = 0.f;
            if (j < a_dim) {
                g = G[row * ld + j];
                acc = fmaf(g, g, acc);
            }
            bot[j] = dscale * g;
        }
        term = (double)acc;
    }
    block_partial(term, partial);
}

// stable: cotangent of the head_r pre-activation  DV top = [0 | v (W_r mubar_L) dtau/drho]
__global__ void reg_stable_rho_kernel(const float* __restrict__ HV, const float* __restrict__ TV,
                                      const float* __restrict__ probe, float* __restrict__ DV, long long N, int a_dim,
                                      int ld) {
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= N) return;
    float* top = DV + row * ld;
    for (int j = 0; j < ld; ++j) top[j] = 0.f;
    for (int j = 0; j < a_dim; ++j) {
        const float r = tanhf(HV[row * ld + a_dim + j]);
        top[a_dim + j] = probe[row * a_dim + j] * TV[row * ld + a_dim + j] * (-2.f * r * (1.f - r * r));
    }
}

int reg_loss_grad_f32(frl_net* net, const float* s, const float* a, const float* t, const float* probe, long long N,
                      long long mean_divisor, int which, float* loss_anti, float* loss_stable, float* grad_anti,
                      float* grad_stable, cudaStream_t st) {
    const int H = net->cfg.hidden, A = net->cfg.a_dim, depth = net->cfg.depth;
    const int K0p = net->d_in_p, NOp = net->n_out_p;
    FRL_REQUIRE(N > 0 && 2 * N < (1ll << 31) / std::max(H, K0p), "batch size out of range");
    {
        int rcp = ensure_pack(net, PACK_F32, st);
        if (rcp != FRL_OK) return rcp;
    }
    const int Ni = (int)N, R2 = 2 * Ni;
    const double div = (double)(mean_divisor > 0 ? mean_divisor : N);
    const float dscale = (float)(2.0 / div);

    auto plan_splits = [&](int tiles) {
        int want = std::max(1, 296 / std::max(1, tiles));
        int maxs = std::max(1, (R2 + 63) / 64);
        return std::min(want, maxs);
    };
    const int splits_hh = plan_splits(((H + 127) / 128) * ((H + 127) / 128));
    const int splits_small = plan_splits(2);
    const int max_splits = std::max(splits_hh, splits_small);
    const int cs_chunks = std::max(1, std::min(148, (Ni + 255) / 256));
    const int cs_rows = (Ni + cs_chunks - 1) / cs_chunks;
    const int row_blocks = (Ni + 255) / 256;

    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) / 4 * 4; return o; };
    size_t o_act[kMaxDepth + 1], o_dz[kMaxDepth];
    o_act[0] = take((size_t)R2 * K0p);
    for (int l = 1; l <= depth; ++l) o_act[l] = take((size_t)R2 * H);
    for (int l = 0; l < depth; ++l) o_dz[l] = take((size_t)R2 * H);
    const size_t o_hv = take((size_t)N * NOp), o_tv = take((size_t)N * NOp), o_dv = take((size_t)R2 * NOp);
    const size_t o_g = take((size_t)N * K0p);
    const size_t o_part = take((size_t)max_splits * H * std::max(H, std::max(K0p, NOp)));
    const size_t o_cs = take((size_t)cs_chunks * std::max(H, NOp));
    const size_t o_lp = take((size_t)row_blocks * 2 + 4);
    int rc = ensure_ws(net, off);
    if (rc != FRL_OK) return rc;
    float* ws = net->ws;
    auto ACT = [&](int l) { return ws + o_act[l]; };
    auto ACTb = [&](int l) { return ws + o_act[l] + (size_t)N * (l == 0 ? K0p : H); };  // rows N..2N-1
    auto DZ = [&](int l) { return ws + o_dz[l]; };
    auto DZb = [&](int l) { return ws + o_dz[l] + (size_t)N * H; };
    float *HV = ws + o_hv, *TV = ws + o_tv, *DV = ws + o_dv, *DVb = ws + o_dv + (size_t)N * NOp, *G = ws + o_g;
    float *part = ws + o_part, *cs = ws + o_cs;
    double* lpart = reinterpret_cast<double*>(ws + o_lp);
    const NetF32& w = net->f32;
    auto ldact = [&](int l) { return l == 0 ? K0p : H; };
    auto ksplit = [&](int splits) { return round_up((R2 + splits - 1) / splits, BK); };

    // ---- primal forward (shared by both losses) ----
    {
        const long long n = 2 * N * K0p;
        reg_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, a, t, probe, ACT(0), N, net->cfg.s_dim, A, K0p);
        FRL_CUDA(cudaGetLastError());
    }
    for (int l = 0; l < depth; ++l) {
        rc = run_gemm<false, EPI_BIAS_RELU>(ACT(l), ldact(l), w.Wt[l], H, ACT(l + 1), H, Ni, H, ldact(l), w.b[l], nullptr,
                                            0, 1, ldact(l), st);
        if (rc != FRL_OK) return rc;
    }
    rc = run_gemm<false, EPI_BIAS>(ACT(depth), H, w.Wht, NOp, HV, NOp, Ni, NOp, H, w.bh, nullptr, 0, 1, H, st);
    if (rc != FRL_OK) return rc;

    // tangent-like forward pass of the bottom rows through the masked trunk, then both heads (no bias)
    auto tangent_forward = [&]() -> int {
        for (int l = 0; l < depth; ++l) {
            int r = run_gemm<false, EPI_MASK>(ACTb(l), ldact(l), w.Wt[l], H, ACTb(l + 1), H, Ni, H, ldact(l), nullptr,
                                              ACT(l + 1), H, 1, ldact(l), st);
            if (r != FRL_OK) return r;
        }
        return run_gemm<false, EPI_NONE>(ACTb(depth), H, w.Wht, NOp, TV, NOp, Ni, NOp, H, nullptr, nullptr, 0, 1, H, st);
    };
    // masked backward chain of `rows` rows starting at row `r0` of DV / DZ (mask = primal activations)
    auto backward_chain = [&](int r0, int rows) -> int {
        const size_t ro = (size_t)r0;
        int r = run_gemm<false, EPI_MASK>(DV + ro * NOp, NOp, w.Wh, H, DZ(depth - 1) + ro * H, H, rows, H, NOp, nullptr,
                                          ACT(depth), H, 1, NOp, st, r0 == 0 ? Ni : rows);
        if (r != FRL_OK) return r;
        for (int l = depth - 1; l >= 1; --l) {
            r = run_gemm<false, EPI_MASK>(DZ(l) + ro * H, H, w.W[l], H, DZ(l - 1) + ro * H, H, rows, H, H, nullptr,
                                          ACT(l), H, 1, H, st, r0 == 0 ? Ni : rows);
            if (r != FRL_OK) return r;
        }
        return FRL_OK;
    };
    // weight / bias gradients from the stacked buffers: dW_l = DZ_l^T ACT_l (2N rows), db_l = colsum of the top rows
    auto weight_grads = [&](float* grad) -> int {
        int r = run_gemm<true, EPI_NONE>(DV, NOp, ACT(depth), H, part, H, NOp, H, R2, nullptr, nullptr, 0, splits_small,
                                         ksplit(splits_small), st);
        if (r != FRL_OK) return r;
        for (int h = 0; h < 2; ++h)
            reduce_wgrad_kernel<<<(A * H + 255) / 256, 256, 0, st>>>(part + (size_t)h * A * H, splits_small,
                                                                     (size_t)NOp * H, H, A, H, grad + net->off_wh[h], 0);
        colsum_partial_kernel<<<dim3((NOp + 63) / 64, cs_chunks), 64, 0, st>>>(DV, NOp, N, NOp, cs_rows, cs);
        for (int h = 0; h < 2; ++h)
            reduce_bias_kernel<<<1, 64, 0, st>>>(cs, cs_chunks, NOp, h * A, A, grad + net->off_bh[h], 0);
        FRL_CUDA(cudaGetLastError());
        for (int l = depth - 1; l >= 0; --l) {
            const int Kp = ldact(l), K = l == 0 ? net->d_in : H;
            const int sp = (Kp == H) ? splits_hh : splits_small;
            r = run_gemm<true, EPI_NONE>(DZ(l), H, ACT(l), Kp, part, Kp, H, Kp, R2, nullptr, nullptr, 0, sp, ksplit(sp), st);
            if (r != FRL_OK) return r;
            reduce_wgrad_kernel<<<(H * K + 255) / 256, 256, 0, st>>>(part, sp, (size_t)H * Kp, Kp, H, K,
                                                                     grad + net->off_w[l], 0);
            colsum_partial_kernel<<<dim3((H + 127) / 128, cs_chunks), 128, 0, st>>>(DZ(l), H, N, H, cs_rows, cs);
            reduce_bias_kernel<<<(H + 127) / 128, 128, 0, st>>>(cs, cs_chunks, H, 0, H, grad + net->off_b[l], 0);
            FRL_CUDA(cudaGetLastError());
        }
        return FRL_OK;
    };

    if (which & 1) {
        // ---- anti-divergence ----
        rc = tangent_forward();  // ACT bottom = u_l, TV = [W_c u_L | W_r u_L]
        if (rc != FRL_OK) return rc;
        reg_anti_heads_kernel<<<row_blocks, 256, 0, st>>>(HV, TV, probe, DV, N, A, NOp, dscale, lpart);
        FRL_CUDA(cudaGetLastError());
        if (loss_anti) loss_final_kernel<<<1, 256, 0, st>>>(lpart, row_blocks, 1.0 / div, loss_anti);
        if (grad_anti) {
            rc = backward_chain(0, R2);  // primal-adjoint (top) and tangent-adjoint (bottom) in one stacked pass
            if (rc != FRL_OK) return rc;
            rc = weight_grads(grad_anti);
            if (rc != FRL_OK) return rc;
        }
    }
    if (which & 2) {
        // ---- Jacobian stability ----
        reg_stable_start_kernel<<<row_blocks, 256, 0, st>>>(HV, probe, DV, N, A, NOp);
        FRL_CUDA(cudaGetLastError());
        rc = backward_chain(Ni, Ni);  // reverse chain: DZ bottom = nu_l
        if (rc != FRL_OK) return rc;
        rc = run_gemm<false, EPI_NONE>(DZb(0), H, w.W[0], K0p, G, K0p, Ni, K0p, H, nullptr, nullptr, 0, 1, H, st);
        if (rc != FRL_OK) return rc;
        reg_stable_g_kernel<<<row_blocks, 256, 0, st>>>(G, ACT(0), N, A, K0p, dscale, lpart);
        FRL_CUDA(cudaGetLastError());
        if (loss_stable) loss_final_kernel<<<1, 256, 0, st>>>(lpart, row_blocks, 1.0 / div, loss_stable);
        if (grad_stable) {
            rc = tangent_forward();  // ACT bottom = mubar_l, TV = [W_c mubar_L | W_r mubar_L]
            if (rc != FRL_OK) return rc;
            reg_stable_rho_kernel<<<row_blocks, 256, 0, st>>>(HV, TV, probe, DV, N, A, NOp);
            FRL_CUDA(cudaGetLastError());
            rc = backward_chain(0, Ni);  // primal backward of the tanh'' term: DZ top
            if (rc != FRL_OK) return rc;
            rc = weight_grads(grad_stable);
            if (rc != FRL_OK) return rc;
        }
    }
    return FRL_OK;
}

// y[i] += c * x[i] with c = alpha * num / (den + eps) (num only: alpha * num; neither: alpha) and c forced to 0 when
// gate[0] <= 0.  num / den / gate are optional device scalars: the warm-up weights of flowril.py:272-284 are formed
// without a host round trip, and `gate` (the previous weight) reproduces `use_anti = args.loss_anti > 0`
// (flowril.py:255-256), which stays false for good once a weight hits zero.  coef_out[0] receives c.
__global__ void grad_axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long long n, float alpha,
                                 const float* __restrict__ num, const float* __restrict__ den, float eps,
                                 const float* __restrict__ gate, float* __restrict__ coef_out) {
    float c = alpha;
    if (num) c *= den ? num[0] / (den[0] + eps) : num[0];
    if (gate && !(gate[0] > 0.f)) c = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fmaf(c, x[i], y[i]);
    if (coef_out) {
        // every block has read gate[] (which may alias coef_out) before the last block to finish writes it
        __shared__ unsigned last;
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned*>(coef_out + 1), 1u) == gridDim.x - 1;
        __syncthreads();
        if (last && threadIdx.x == 0) {
            coef_out[0] = c;
            reinterpret_cast<unsigned*>(coef_out + 1)[0] = 0u;
        }
    }
}

int grad_axpy(float* y, const float* x, long long n, float alpha, const float* num, const float* den, float eps,
              const float* gate, float* coef_out, cudaStream_t st) {
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 592));
    grad_axpy_kernel<<<blocks, 256, 0, st>>>(y, x, n, alpha, num, den, eps, gate, coef_out);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

}  // namespace frl
