// Internal declarations shared by the translation units of libflowril_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/flowril_b200.h"

namespace frl {

constexpr int kMaxDepth = FRL_MAX_DEPTH;
constexpr int kMaxSteps = FRL_MAX_STEPS;

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FRL_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::frl::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define FRL_REQUIRE(cond, msg)                         \
    do {                                               \
        if (!(cond)) {                                 \
            ::frl::set_error(std::string("invalid argument: ") + (msg)); \
            return FRL_ERR_INVALID;                    \
        }                                              \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Device-side view of one packed net for the FP32 (FFMA) kernels.
struct NetF32 {
    const float* Wt[kMaxDepth];  // trunk layer l, K-major (transposed): [Kp_l][H], Kp_0 = round_up(d_in,16) zero padded
    const float* W[kMaxDepth];   // trunk layer l, reference layout padded: [H][Kp_l]
    const float* b[kMaxDepth];   // [H]
    const float* Wh;             // heads, reference layout, rows padded: [NOp][H], NOp = round_up(n_heads*a_dim,16)
    const float* Wht;            // heads, K-major: [H][NOp]
    const float* bh;             // [NOp]
};

}  // namespace frl

// Opaque handle of the C ABI.
struct frl_net {
    frl_net_config cfg;
    int d_in, d_in_p, n_out, n_out_p;   // a+s+1, padded; n_heads*a_dim, padded
    long long n_params;
    long long off_w[frl::kMaxDepth], off_b[frl::kMaxDepth];  // offsets into the flat vector
    long long off_wh[2], off_bh[2];
    float* pack_f32 = nullptr;          // one allocation holding every NetF32 array
    size_t pack_f32_floats = 0;
    frl::NetF32 f32;
    // tensor-core packing (bf16 hi/mid/lo canonical UMMA tiles), see logprob_tc.cu
    void* pack_tc = nullptr;
    size_t pack_tc_bytes = 0;
    void* pack_tc3 = nullptr;           // per-CTA-rank block streams of the CTA-pair integrator (logprob_tc3.cu)
    void* pack_tc3x = nullptr;          // the same for the 3-term split integrator (logprob_tc3x.cu): scaled W_hi / W_lo blocks
    bool has_params = false;
    // Lazy packing: frl_net_set_params only records the master vector and bumps `version`; each kernel family
    // (re)builds its own layout on first use after a change (frl::ensure_pack), on the consumer's stream.
    const float* params_dev = nullptr;
    unsigned long long version = 0;
    unsigned long long packed[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t pack_event[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t pack_stream[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // Streams on which kernels that READ the packed layouts were launched since the layouts were built, with an event
    // recorded after the latest such launch (frl::note_reader): a re-pack after frl_net_set_params rewrites the layouts in
    // place, so its stream first waits for every reader on another stream.
    static constexpr int kMaxReaders = 4;
    cudaEvent_t reader_event[kMaxReaders] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t reader_stream[kMaxReaders] = {nullptr, nullptr, nullptr, nullptr};
    bool reader_used[kMaxReaders] = {false, false, false, false};
    // tensor-core training: fp16 weight images (train_tc.cu) and workspace
    void* pack_train = nullptr;
    void* ws_tc = nullptr;
    size_t ws_tc_bytes = 0;
    unsigned* train_flags = nullptr;   // inside ws_tc: [0] = time-out count of the fused training kernel's global waits
    // training workspace (grown on demand)
    float* ws = nullptr;
    size_t ws_floats = 0;
    // host-staging state for frl_score_host
    void* host_stage = nullptr;
    void* ws_x0s = nullptr;             // per-launch layer-0 operand images of wide-input nets (logprob_tc3.cu)
    size_t ws_x0s_bytes = 0;
    cudaEvent_t x0s_event = nullptr;    // recorded after the launch that reads ws_x0s (the next launch, on any stream, waits)
    // Workspaces that were outgrown stay allocated until frl_net_destroy: a captured CUDA graph of an update (or a launch
    // still in flight) may reference them, so a later, larger call must not free the memory under it.
    std::vector<void*> retired;
};

void frl_host_stage_free(frl_net* net);  // api.cu

namespace frl {

// net.cu: weight layouts, built lazily
enum PackKind { PACK_F32 = 0, PACK_TC1 = 1, PACK_TC2 = 2, PACK_TC3 = 3, PACK_TRAIN = 4, PACK_TC3X = 5, PACK_KINDS = 6 };
int ensure_pack(const frl_net* net, int kind, cudaStream_t stream);
int note_reader(const frl_net* net, cudaStream_t stream);   // call after launching kernels that read the net's packed layouts
int pack_f32(frl_net* net, const float* params_dev, cudaStream_t stream);

// logprob_f32.cu
struct ScoreArgs {
    const frl_net* net[2];  // role slots (expert, agent); net[1] may be null when n_roles == 1
    float sign[2];
    int n_roles;
    const float* s;
    const float* a;
    const float* probe;
    int probe_mode;
    const float* t_mid_host;
    int n_steps;
    int reward_type;        // -1: no reward output
    float* out_logp[2];
    float* out_reward;
    long long B;
};
int launch_logprob_f32(const ScoreArgs& args, cudaStream_t stream);
int launch_vfield_f32(const frl_net* net, const float* s, const float* a, const float* t, float* out0, float* out1,
                      long long B, cudaStream_t stream);
int launch_reward(const float* logp_e, const float* logp_a, float* reward, long long B, int reward_type,
                  cudaStream_t stream);

// logprob_tc.cu
namespace tc {
constexpr int kMaxBlocks = 72;
struct PackBlk {
    uint32_t goff;                   // bytes
    int nrows, klen, n0, k0, layer;  // layer == depth -> heads;  k0 < 0 -> bias block (rows k=0..2 = hi/mid/lo split)
    int part = 0;                    // 0: the rounded value; 1 / 2: hi / lo term of the 2-term 16-bit split of (scale * value)
};
}  // namespace tc
int tc_pack(frl_net* net, const float* params_dev, cudaStream_t stream);
int tc_pack_images(frl_net* net, const std::vector<tc::PackBlk>& pack, uint32_t image_bytes, void** dst,
                   const float* params_dev, cudaStream_t stream, float scale = 1.f);
int launch_logprob_tc(const ScoreArgs& args, int precision, cudaStream_t stream);

// logprob_tc3.cu
int tc3_pack(frl_net* net, const float* params_dev, cudaStream_t stream);
int launch_logprob_tc3(const ScoreArgs& args, int precision, cudaStream_t stream);

// logprob_tc3x.cu (FRL_PREC_FP16X3)
bool tc3x_supported(const frl_net* net);
int tc3x_pack(frl_net* net, const float* params_dev, cudaStream_t stream);
int launch_logprob_tc3x(const ScoreArgs& args, cudaStream_t stream);

// train_f32.cu
int cfm_loss_grad_f32(frl_net* net, float role_sign, const float* s, const float* a0, const float* t,
                      const float* noise, long long B, long long mean_divisor, float grad_scale, int accumulate,
                      float* loss, float* grad, cudaStream_t stream);

int reg_loss_grad_f32(frl_net* net, const float* s, const float* a, const float* t, const float* probe, long long N,
                      long long mean_divisor, int which, float* loss_anti, float* loss_stable, float* grad_anti,
                      float* grad_stable, cudaStream_t stream);
int grad_axpy(float* y, const float* x, long long n, float alpha, const float* num, const float* den, float eps,
              const float* gate, float* coef_out, cudaStream_t stream);

// train_tc.cu
int tc_train_pack(frl_net* net, const float* params_dev, cudaStream_t stream);
// one CFM loss term (a role's batch); up to two terms are stacked row-wise into ONE pass over the net
struct CfmTerm {
    float sign;
    const float *s, *a0, *t, *noise;
    long long B, mean_divisor;
    float grad_scale;
    float* loss;   // device scalar (may be null)
};
int cfm_loss_grad_tc(frl_net* net, const CfmTerm* terms, int n_terms, int accumulate, float* grad, cudaStream_t stream);
long long train_timeouts(const frl_net* net);

// adam.cu
int clip_adam(const frl_adam_segment* segs, int n_segs, long long step, long long* step_dev, double lr, double beta1,
              double beta2, double eps, double max_norm, float* norm_out, cudaStream_t stream);
int tensor_norms(const float* flat, const long long* offsets_host, int n_tensors, float* out, cudaStream_t stream);

// returns.cu
int normalise_returns(const float* reward, const float* masks, int T, int N, float gamma, float* returns,
                      int returns_valid, float* mean_var, double* count, float* out, cudaStream_t stream);

}  // namespace frl
