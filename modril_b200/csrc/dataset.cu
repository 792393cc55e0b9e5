// Demonstration pre-processing of the offline pre-trainer (modril/flowril/pretrain.py:33-39, :269-285):
//   mean = x.mean(0), std = x.std(0) (unbiased), x <- clamp((x - mean) / (std + 1e-8), -10, 10)
// One CTA per column accumulates the moments in double (two passes, fixed order: deterministic), a second kernel
// normalises.  One-time data preparation; the demonstrations then stay resident in HBM for the whole run.
#include "frl_internal.h"

namespace frl {

__global__ void __launch_bounds__(256) column_moments_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ mean,
                                                             float* __restrict__ stdv) {
    __shared__ double sh[256];
    __shared__ double s_mean;
    const int c = blockIdx.x, tid = threadIdx.x;
    double acc = 0.0;
    for (long long r = tid; r < n; r += 256) acc += (double)x[r * d + c];
    sh[tid] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    if (tid == 0) s_mean = sh[0] / (double)n;
    __syncthreads();
    const double m = s_mean;
    acc = 0.0;
    for (long long r = tid; r < n; r += 256) { const double dv = (double)x[r * d + c] - m; acc += dv * dv; }
    __syncthreads();
    sh[tid] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    if (tid == 0) {
        mean[c] = (float)m;
        stdv[c] = (float)sqrt(sh[0] / (double)(n - 1));   // torch.std: Bessel's correction (nan for n == 1, like torch)
    }
}

__global__ void norm_vec_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv,
                                float* __restrict__ out, long long total, int d, float eps, float lim) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % d);
    const float v = __fdiv_rn(x[i] - mean[c], stdv[c] + eps);
    out[i] = fminf(fmaxf(v, -lim), lim);
}

}  // namespace frl

using namespace frl;

extern "C" int frl_column_moments(const float* x_dev, long long n, int d, float* mean_dev, float* std_dev, int device, void* stream) {
    FRL_REQUIRE(x_dev && mean_dev && std_dev, "null pointer");
    FRL_REQUIRE(n >= 1 && d >= 1 && d <= 65535, "bad shape");
    FRL_CUDA(cudaSetDevice(device));
    column_moments_kernel<<<d, 256, 0, (cudaStream_t)stream>>>(x_dev, n, d, mean_dev, std_dev);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

extern "C" int frl_norm_vec(const float* x_dev, const float* mean_dev, const float* std_dev, long long n, int d, float* out_dev,
                            int device, void* stream) {
    FRL_REQUIRE(x_dev && mean_dev && std_dev && out_dev, "null pointer");
    FRL_REQUIRE(n >= 0 && d >= 1, "bad shape");
    if (n == 0) return FRL_OK;
    FRL_CUDA(cudaSetDevice(device));
    const long long total = n * d;
    norm_vec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_dev, mean_dev, std_dev, out_dev, total, d,
                                                                                      1e-8f, 10.f);
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}
