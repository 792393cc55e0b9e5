// Returns / GAE recursion and advantage normalisation of the PPO consumer of the rewards (scope row f4, first part).
// Replaces  RolloutStorage.compute_returns  (deps/rl-toolkit/rlf/storage/rollout_storage.py:178-233) and
// compute_advantages  (:235-239): T sequential python-loop steps of small elementwise torch ops become one launch.
// One thread owns one (environment, value-dim) column and walks the T steps backwards with the reference's float32
// operation order (explicit _rn intrinsics, no FMA contraction), so the scan is bit-identical to the torch loop.
#include <algorithm>

#include "frl_internal.h"

namespace frl {

__global__ void compute_returns_kernel(const float* __restrict__ rewards, float* __restrict__ value_preds,
                                       const float* __restrict__ masks, const float* __restrict__ bad_masks,
                                       const float* __restrict__ next_value, int T, int N, int D, float gamma,
                                       float gamma_lambda, int use_gae, float* __restrict__ returns) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // column = n * D + d
    if (c >= N * D) return;
    const int n = c / D;
    const size_t ND = (size_t)N * D;
    if (use_gae) {
        value_preds[(size_t)T * ND + c] = next_value[c];     // rollout_storage.py:185 / :210
        float gae = 0.f;
        float v_next = next_value[c];
        for (int t = T - 1; t >= 0; --t) {
            const float m = masks[(size_t)(t + 1) * N + n];
            const float v = value_preds[(size_t)t * ND + c];
            // delta = r + gamma * V[t+1] * mask[t+1] - V[t]
            const float delta = __fsub_rn(__fadd_rn(rewards[(size_t)t * N + n], __fmul_rn(__fmul_rn(gamma, v_next), m)), v);
            // gae = delta + (gamma * lambda) * mask[t+1] * gae
            gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, m), gae));
            if (bad_masks) gae = __fmul_rn(gae, bad_masks[(size_t)(t + 1) * N + n]);   // :198
            returns[(size_t)t * ND + c] = __fadd_rn(gae, v);
            v_next = v;
        }
    } else {
        float ret = next_value[c];
        returns[(size_t)T * ND + c] = ret;                   // :201 / :226
        for (int t = T - 1; t >= 0; --t) {
            const float m = masks[(size_t)(t + 1) * N + n];
            ret = __fadd_rn(__fmul_rn(__fmul_rn(ret, gamma), m), rewards[(size_t)t * N + n]);
            if (bad_masks) {                                  // :205-211
                const float b = bad_masks[(size_t)(t + 1) * N + n];
                ret = __fadd_rn(__fmul_rn(ret, b), __fmul_rn(__fsub_rn(1.f, b), value_preds[(size_t)t * ND + c]));
            }
            returns[(size_t)t * ND + c] = ret;
        }
    }
}

constexpr int kAdvThreads = 256;

// adv = returns - value_preds over the first n elements; per-block partial sums of adv and adv^2 in double
__global__ void __launch_bounds__(kAdvThreads) adv_partial_kernel(const float* __restrict__ returns,
                                                                  const float* __restrict__ value_preds, long long n,
                                                                  float* __restrict__ adv, double* __restrict__ partial) {
    double s1 = 0.0, s2 = 0.0;
    for (long long i = blockIdx.x * (long long)kAdvThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kAdvThreads) {
        const float a = __fsub_rn(returns[i], value_preds[i]);
        adv[i] = a;
        s1 += (double)a;
        s2 += (double)a * (double)a;
    }
    __shared__ double sh[2][kAdvThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < kAdvThreads / 32; ++i) { a += sh[0][i]; b += sh[1][i]; }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
    }
}

// (adv - mean) / (std + eps), std with Bessel's correction (torch.Tensor.std default); every block re-reduces the
// few partials in the same order
__global__ void __launch_bounds__(kAdvThreads) adv_normalise_kernel(float* __restrict__ adv, long long n,
                                                                    const double* __restrict__ partial, int n_partial,
                                                                    float eps, float* __restrict__ stats_out) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < n_partial; ++i) { s1 += partial[2 * i]; s2 += partial[2 * i + 1]; }
    const double mean = s1 / (double)n;
    const double var = n > 1 ? fmax(0.0, (s2 - s1 * mean) / (double)(n - 1)) : nan("");
    const float mean_f = (float)mean, std_f = (float)sqrt(var);
    if (stats_out && blockIdx.x == 0 && threadIdx.x == 0) { stats_out[0] = mean_f; stats_out[1] = std_f; }
    const float den = __fadd_rn(std_f, eps);
    for (long long i = blockIdx.x * (long long)kAdvThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kAdvThreads)
        adv[i] = __fdiv_rn(__fsub_rn(adv[i], mean_f), den);
}

}  // namespace frl

using namespace frl;

extern "C" int frl_compute_returns(const float* rewards, float* value_preds, const float* masks, const float* bad_masks,
                                   const float* next_value, int T, int N, int D, double gamma, double gae_lambda,
                                   int use_gae, float* returns, int device, void* stream) {
    FRL_REQUIRE(rewards && value_preds && masks && next_value && returns, "null pointer");
    FRL_REQUIRE(T >= 1 && N >= 1 && D >= 1 && (long long)N * D < (1ll << 31), "bad T / N / D");
    FRL_CUDA(cudaSetDevice(device));
    const int cols = N * D;
    compute_returns_kernel<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        rewards, value_preds, masks, bad_masks, next_value, T, N, D, (float)gamma, (float)(gamma * gae_lambda), use_gae, returns);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_normalise_advantages(const float* returns, const float* value_preds, long long n, float eps, float* adv,
                                        double* scratch, float* stats_out, int device, void* stream) {
    FRL_REQUIRE(returns && value_preds && adv && scratch, "null pointer");
    FRL_REQUIRE(n >= 1, "empty rollout");
    FRL_CUDA(cudaSetDevice(device));
    const int blocks = (int)std::min<long long>((n + kAdvThreads - 1) / kAdvThreads, 128);
    cudaStream_t st = (cudaStream_t)stream;
    adv_partial_kernel<<<blocks, kAdvThreads, 0, st>>>(returns, value_preds, n, adv, scratch);
    const int nblocks = (int)std::min<long long>((n + kAdvThreads - 1) / kAdvThreads, 592);
    adv_normalise_kernel<<<nblocks, kAdvThreads, 0, st>>>(adv, n, scratch, blocks, eps, stats_out);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}
