// Global grad-norm clip + Adam over flat fp32 segments, deterministic (fixed reduction order, no float atomics).
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step as called at
// modril/flowril/flowril.py:324-330 (and modril/flowril/pretrain.py:417-420).
#include <cmath>

#include "frl_internal.h"

namespace frl {

constexpr int kNormBlocks = 296;  // 2 per SM
constexpr int kNormThreads = 256;

struct Segs {
    frl_adam_segment s[4];
    long long first[5];
    int n;
};

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[kNormThreads / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kNormThreads / 32; ++i) t += sh[i];
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(kNormThreads) sumsq_kernel(const __grid_constant__ Segs sg, double* __restrict__ partial,
                                                             long long* __restrict__ step_dev) {
    // device-resident step counter (CUDA-graph replay): advance it here, the Adam kernel that follows reads it
    if (step_dev && blockIdx.x == 0 && threadIdx.x == 0) step_dev[0] += 1;
    const long long total = sg.first[sg.n];
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = blockIdx.x * per, hi = min(total, lo + per);
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        // segment lookup with compile-time indices: the parameter struct stays in the constant bank
        const float* gp = sg.s[0].grad + i;
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if (k < sg.n && i >= sg.first[k]) gp = sg.s[k].grad + (i - sg.first[k]);
        const float g = *gp;
        acc += (double)g * (double)g;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

struct AdamScalars {
    float one_minus_b1, b2, one_minus_b2, sqrt_bc2, eps, neg_step_size, max_norm;
    double lr, beta1, beta2;   // used only with a device-resident step counter
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ Segs sg, const double* __restrict__ partial, int n_partial,
                                                   AdamScalars sc, float* __restrict__ norm_out,
                                                   const long long* __restrict__ step_dev) {
    __shared__ float s_coef, s_sqrt_bc2, s_neg_step;
    // total of the per-block partial sums: every block forms it with the same fixed tree (all threads load in parallel;
    // a serial loop in thread 0 is ~75 dependent L2 round trips in EVERY block)
    double part = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) part += partial[i];
    const double total_sq = block_sum(part);   // valid in thread 0
    if (threadIdx.x == 0) {
        s_sqrt_bc2 = sc.sqrt_bc2;
        s_neg_step = sc.neg_step_size;
        if (step_dev) {   // bias corrections from the device counter (same formulas as the host path below)
            const double st = (double)step_dev[0];
            s_sqrt_bc2 = (float)sqrt(1.0 - pow(sc.beta2, st));
            s_neg_step = (float)(-sc.lr / (1.0 - pow(sc.beta1, st)));
        }
        const float norm = (float)sqrt(total_sq);
        float coef = sc.max_norm / (norm + 1e-6f);               // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
        s_coef = coef < 1.f ? coef : 1.f;
        if (blockIdx.x == 0 && norm_out) norm_out[0] = norm;
    }
    __syncthreads();
    const float coef = s_coef;
    sc.sqrt_bc2 = s_sqrt_bc2;
    sc.neg_step_size = s_neg_step;
    const long long total = sg.first[sg.n];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // segment lookup with compile-time indices (see sumsq_kernel)
        float* pp = sg.s[0].params + i;
        const float* gp = sg.s[0].grad + i;
        float* mp = sg.s[0].exp_avg + i;
        float* vp = sg.s[0].exp_avg_sq + i;
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if (k < sg.n && i >= sg.first[k]) {
                const long long j = i - sg.first[k];
                pp = sg.s[k].params + j;
                gp = sg.s[k].grad + j;
                mp = sg.s[k].exp_avg + j;
                vp = sg.s[k].exp_avg_sq + j;
            }
        const float g = *gp * coef;
        float m = *mp;
        float v = *vp;
        m = m + (g - m) * sc.one_minus_b1;                         // exp_avg.lerp_(grad, 1 - beta1)
        v = v * sc.b2 + sc.one_minus_b2 * g * g;                   // mul_(beta2).addcmul_(g, g, 1 - beta2)
        const float denom = sqrtf(v) / sc.sqrt_bc2 + sc.eps;  // sqrt(v)/sqrt(bc2) + eps
        *pp = *pp + sc.neg_step_size * (m / denom);
        *mp = m;
        *vp = v;
    }
}

int clip_adam(const frl_adam_segment* segs, int n_segs, long long step, long long* step_dev, double lr, double beta1,
              double beta2, double eps, double max_norm, float* norm_out, cudaStream_t stream) {
    Segs sg{};
    sg.n = n_segs;
    long long total = 0;
    for (int i = 0; i < n_segs; ++i) {
        sg.s[i] = segs[i];
        sg.first[i] = total;
        total += segs[i].n;
    }
    sg.first[n_segs] = total;
    if (total == 0) return FRL_OK;
    double* partial = nullptr;
    FRL_CUDA(cudaMallocAsync((void**)&partial, kNormBlocks * sizeof(double), stream));
    sumsq_kernel<<<kNormBlocks, kNormThreads, 0, stream>>>(sg, partial, step_dev);
    FRL_CUDA(cudaGetLastError());
    const double bc1 = 1.0 - std::pow(beta1, (double)step);
    const double bc2 = 1.0 - std::pow(beta2, (double)step);
    AdamScalars sc;
    sc.one_minus_b1 = (float)(1.0 - beta1);
    sc.b2 = (float)beta2;
    sc.one_minus_b2 = (float)(1.0 - beta2);
    sc.sqrt_bc2 = (float)std::sqrt(bc2);
    sc.eps = (float)eps;
    sc.neg_step_size = (float)(-lr / bc1);
    sc.max_norm = (float)max_norm;
    sc.lr = lr; sc.beta1 = beta1; sc.beta2 = beta2;
    int blocks = (int)std::min<long long>((total + 255) / 256, 592);
    adam_kernel<<<blocks, 256, 0, stream>>>(sg, partial, kNormBlocks, sc, norm_out, step_dev);
    FRL_CUDA(cudaGetLastError());
    FRL_CUDA(cudaFreeAsync(partial, stream));
    return FRL_OK;
}


// Per-tensor L2 norms of a flat gradient (GradNorm2D.update sums g.norm() over the parameter tensors,
// modril/flowril/networks.py:108-117).  One block per tensor, fixed reduction order.
struct NormTable {
    long long first[2 * kMaxDepth + 5];
    int n;
};
__global__ void __launch_bounds__(kNormThreads) tensor_norms_kernel(const float* __restrict__ flat, NormTable tab,
                                                                    float* __restrict__ out) {
    const long long lo = tab.first[blockIdx.x], hi = tab.first[blockIdx.x + 1];
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += (double)flat[i] * (double)flat[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = (float)sqrt(acc);
}

int tensor_norms(const float* flat, const long long* offsets_host, int n_tensors, float* out, cudaStream_t stream) {
    NormTable tab{};
    FRL_REQUIRE(n_tensors >= 1 && n_tensors <= 2 * kMaxDepth + 4, "1..20 tensors");
    tab.n = n_tensors;
    for (int i = 0; i <= n_tensors; ++i) tab.first[i] = offsets_host[i];
    tensor_norms_kernel<<<n_tensors, kNormThreads, 0, stream>>>(flat, tab, out);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

}  // namespace frl
