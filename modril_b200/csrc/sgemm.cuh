// FP32 SGEMM building block shared by the FFMA training kernels (train_f32.cu, bcf_f32.cu, ppo_f32.cu).
#pragma once
#include <cuda_runtime.h>

#include "frl_internal.h"

namespace frl {

// EPI_BIAS_TANH: tanh(acc + bias) (the PPO actor / critic trunks); EPI_DTANH: acc * (1 - h^2) with h = mask (their backward)
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2, EPI_MASK = 3, EPI_BIAS_TANH = 4, EPI_DTANH = 5 };
constexpr int BK = 16;

// C[M][N] (+ split offset) = op(A) * B.   A_TRANS=false: A is [M][lda] (k contiguous); true: A is [K][lda]
// (m contiguous).  B is [K][ldb].  M guarded per row, N per float4, K per row (zero fill).
template <int BM, int BN, int TM, int TN, bool A_TRANS, int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                    int ldb, float* __restrict__ C, int ldc, int M, int N, int K,
                                                    const float* __restrict__ bias, const float* __restrict__ mask,
                                                    int ldmask, int k_split_len, int mask_rows) {
    static_assert((BM / TM) * (BN / TN) == 256, "thread tiling must cover the block tile with 256 threads");
    constexpr int AP = 4, BP = 4;
    __shared__ __align__(16) float As[BK][BM + AP];
    __shared__ __align__(16) float Bs[BK][BN + BP];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_lo = blockIdx.z * k_split_len;
    const int k_hi = min(K, k_lo + k_split_len);
    C += (size_t)blockIdx.z * M * ldc;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
        if (A_TRANS) {
            for (int idx = tid; idx < BK * BM / 4; idx += 256) {
                const int kk = idx / (BM / 4), mq = idx % (BM / 4);
                const int k = k0 + kk, m = m0 + mq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < k_hi && m < M) v = *reinterpret_cast<const float4*>(A + (size_t)k * lda + m);
                *reinterpret_cast<float4*>(&As[kk][mq * 4]) = v;
            }
        } else {
            for (int idx = tid; idx < BM * BK / 4; idx += 256) {
                const int mm = idx / (BK / 4), kq = idx % (BK / 4);
                const int m = m0 + mm, k = k0 + kq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < M && k < k_hi) v = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
                As[kq * 4 + 0][mm] = v.x;
                As[kq * 4 + 1][mm] = v.y;
                As[kq * 4 + 2][mm] = v.z;
                As[kq * 4 + 3][mm] = v.w;
            }
        }
        for (int idx = tid; idx < BK * BN / 4; idx += 256) {
            const int kk = idx / (BN / 4), nq = idx % (BN / 4);
            const int k = k0 + kk, n = n0 + nq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < k_hi && n < N) v = *reinterpret_cast<const float4*>(B + (size_t)k * ldb + n);
            *reinterpret_cast<float4*>(&Bs[kk][nq * 4]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_TANH) v += __ldg(bias + n);
            if (EPI == EPI_BIAS_RELU) v = v > 0.f ? v : 0.f;
            if (EPI == EPI_BIAS_TANH) v = tanhf(v);
            if (EPI == EPI_DTANH) { const float h = mask[(size_t)m * ldmask + n]; v *= 1.f - h * h; }
            // mask_rows < M: the rows are two stacked batches that share one mask matrix (regularisers)
            if (EPI == EPI_MASK) v = mask[(size_t)(m >= mask_rows ? m - mask_rows : m) * ldmask + n] > 0.f ? v : 0.f;
            C[(size_t)m * ldc + n] = v;
        }
    }
}

template <bool A_TRANS, int EPI>
static int run_gemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                    const float* bias, const float* mask, int ldmask, int splits, int k_split_len,
                    cudaStream_t st, int mask_rows = 0) {
    if (mask_rows <= 0) mask_rows = M;
    // pick a tile shape by the narrow dimension
    if (N <= 16) {
        dim3 grid((N + 15) / 16, (M + 127) / 128, splits);
        sgemm_kernel<128, 16, 4, 2, A_TRANS, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask,
                                                                         ldmask, k_split_len, mask_rows);
    } else if (M <= 16) {
        dim3 grid((N + 127) / 128, (M + 15) / 16, splits);
        sgemm_kernel<16, 128, 1, 8, A_TRANS, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask,
                                                                         ldmask, k_split_len, mask_rows);
    } else if ((long long)((M + 63) / 64) * ((N + 127) / 128) * splits < 74) {
        // small problems (the reference's 128-row updates): more, smaller CTAs; the k order per output is unchanged
        dim3 grid((N + 31) / 32, (M + 31) / 32, splits);
        sgemm_kernel<32, 32, 2, 2, A_TRANS, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask,
                                                                        ldmask, k_split_len, mask_rows);
    } else if (N <= 64) {
        dim3 grid((N + 63) / 64, (M + 127) / 128, splits);
        sgemm_kernel<128, 64, 8, 4, A_TRANS, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask,
                                                                         ldmask, k_split_len, mask_rows);
    } else if (M <= 64) {
        dim3 grid((N + 127) / 128, (M + 63) / 64, splits);
        sgemm_kernel<64, 128, 4, 8, A_TRANS, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask,
                                                                         ldmask, k_split_len, mask_rows);
    } else {
        dim3 grid((N + 127) / 128, (M + 127) / 128, splits);
        sgemm_kernel<128, 128, 8, 8, A_TRANS, EPI><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask,
                                                                          ldmask, k_split_len, mask_rows);
    }
    FRL_CUDA(cudaGetLastError());
    return FRL_OK;
}

// grad[n*K + k] (+)= sum_s partial[s*split_stride + n*ldp + k]  for n < n_rows, k < K  (ldp >= K)
static __global__ void reduce_wgrad_kernel(const float* __restrict__ partial, int splits, size_t split_stride, int ldp,
                                    int n_rows, int K, float* __restrict__ grad, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * K) return;
    const int n = i / K, k = i % K;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[(size_t)s * split_stride + (size_t)n * ldp + k];
    grad[i] = accumulate ? grad[i] + acc : acc;
}

// column sums of D [B][ld] over rows, split into chunks; partial[chunk][col]
static __global__ void colsum_partial_kernel(const float* __restrict__ D, int ld, long long B, int cols, int rows_per_chunk,
                                      float* __restrict__ partial) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const long long lo = (long long)blockIdx.y * rows_per_chunk;
    const long long hi = min(B, lo + rows_per_chunk);
    float acc = 0.f;
    for (long long r = lo; r < hi; ++r) acc += D[r * ld + col];
    partial[(size_t)blockIdx.y * cols + col] = acc;
}

static __global__ void reduce_bias_kernel(const float* __restrict__ partial, int chunks, int ldp, int col0, int n,
                                   float* __restrict__ grad, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += partial[(size_t)c * ldp + col0 + i];
    grad[i] = accumulate ? grad[i] + acc : acc;
}

}  // namespace frl
