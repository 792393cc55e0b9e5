// Return normalisation of the reward hook for a whole rollout in ONE launch.
// Replaces the per-step host loop of  FlowMatchingEstimation._get_reward  (modril/flowril/flowril.py:221-228):
//     returns = returns * masks[step] * gamma + reward          (float32 torch ops, three roundings)
//     ret_rms.update(returns.cpu().numpy())                      (RunningMeanStd, float32 numpy,
//                                                                 deps/rl-toolkit/rlf/baselines/common/running_mean_std.py:3-35)
//     reward / sqrt(ret_rms.var[0] + 1e-8)
// which costs one device->host sync per rollout step (128 per update).  The scan over steps is inherently sequential
// (the running moments of step t scale the reward of step t), so one CTA walks the T steps and reduces over the N
// environments inside each step.  Batch moments are accumulated in double and rounded to float32 (numpy's float32
// pairwise sums differ from that by at most an ulp); the running-moment merge follows numpy's float32 arithmetic
// with the count kept in double like the reference's python float.
#include "frl_internal.h"

namespace frl {

constexpr int kRetThreads = 1024;

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // protect sh from the previous use
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kRetThreads / 32; ++i) t += sh[i];  // same order in every thread
    return t;
}

__global__ void __launch_bounds__(kRetThreads) normalise_returns_kernel(const float* __restrict__ reward,
                                                                        const float* __restrict__ masks, int T, int N,
                                                                        float gamma, float* __restrict__ returns,
                                                                        int returns_valid, float* __restrict__ mean_var,
                                                                        double* __restrict__ count_io,
                                                                        float* __restrict__ out) {
    __shared__ double sh[kRetThreads / 32];
    float mean = mean_var[0], var = mean_var[1];
    double count = count_io[0];
    for (int t = 0; t < T; ++t) {
        const float* r = reward + (size_t)t * N;
        const float* m = masks + (size_t)t * N;
        // returns recursion (first ever call: returns = reward.clone() and the recursion is still applied)
        double s1 = 0.0;
        for (int n = threadIdx.x; n < N; n += kRetThreads) {
            const float prev = (t == 0 && !returns_valid) ? r[n] : returns[n];
            const float ret = __fadd_rn(__fmul_rn(__fmul_rn(prev, m[n]), gamma), r[n]);
            returns[n] = ret;
            s1 += (double)ret;
        }
        const float bm = (float)(block_reduce_sum(s1, sh) / (double)N);
        double s2 = 0.0;
        for (int n = threadIdx.x; n < N; n += kRetThreads) {
            const float d = __fsub_rn(returns[n], bm);
            s2 += (double)__fmul_rn(d, d);
        }
        const float bv = (float)(block_reduce_sum(s2, sh) / (double)N);
        // RunningMeanStd.update (float32 array arithmetic, python-float count)
        const float delta = __fsub_rn(bm, mean);
        const double tot_d = count + (double)N;
        const float tot = (float)tot_d, cnt = (float)count, nf = (float)N;
        const float new_mean = __fadd_rn(mean, __fdiv_rn(__fmul_rn(delta, nf), tot));
        const float m_a = __fmul_rn(var, cnt);
        const float m_b = __fmul_rn(bv, nf);
        const float m_c = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(delta, delta), cnt), nf), tot);
        const float m2 = __fadd_rn(__fadd_rn(m_a, m_b), m_c);
        mean = new_mean;
        var = __fdiv_rn(m2, tot);
        count = tot_d;
        const float scale = sqrtf(__fadd_rn(var, 1e-8f));
        for (int n = threadIdx.x; n < N; n += kRetThreads) out[(size_t)t * N + n] = __fdiv_rn(r[n], scale);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mean_var[0] = mean;
        mean_var[1] = var;
        count_io[0] = count;
    }
}

int normalise_returns(const float* reward, const float* masks, int T, int N, float gamma, float* returns,
                      int returns_valid, float* mean_var, double* count, float* out, cudaStream_t stream) {
    normalise_returns_kernel<<<1, kRetThreads, 0, stream>>>(reward, masks, T, N, gamma, returns, returns_valid,
                                                            mean_var, count, out);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

}  // namespace frl
