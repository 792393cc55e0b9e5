// C-ABI entry points (see include/flowril_b200.h for the contract and the reference interfaces replaced).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

#include "frl_internal.h"

using namespace frl;

namespace {

int check_net(const frl_net* n) {
    FRL_REQUIRE(n != nullptr, "null net");
    FRL_REQUIRE(n->has_params, "frl_net_set_params has not been called on this net");
    return FRL_OK;
}

int check_pair(const frl_net* e, const frl_net* a) {
    int rc = check_net(e);
    if (rc != FRL_OK) return rc;
    rc = check_net(a);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(e->cfg.s_dim == a->cfg.s_dim && e->cfg.a_dim == a->cfg.a_dim && e->cfg.hidden == a->cfg.hidden &&
                    e->cfg.depth == a->cfg.depth && e->cfg.n_heads == a->cfg.n_heads && e->cfg.device == a->cfg.device,
                "expert and agent nets must share dimensions and device");
    return FRL_OK;
}

int dispatch_score(const ScoreArgs& args, int precision, cudaStream_t st) {
    if (args.B <= 0) return FRL_OK;  // empty batch: nothing to do (pointers may legitimately be null)
    FRL_REQUIRE(args.n_steps >= 1 && args.n_steps <= kMaxSteps, "n_steps must be in 1..256");
    FRL_REQUIRE(args.t_mid_host != nullptr, "null t_mid");
    FRL_REQUIRE(args.s && args.a, "null s/a");
    FRL_REQUIRE(args.probe_mode == FRL_PROBE_EXACT || args.probe != nullptr, "null probe");
    FRL_REQUIRE(args.probe_mode >= 0 && args.probe_mode <= 2, "bad probe_mode");
    int rc;
    if (precision == FRL_PREC_FP32) {
        rc = launch_logprob_f32(args, st);
    } else if (precision == FRL_PREC_FP16X3) {
        // fp32-grade products on the tensor cores (3-term fp16 split).  Nets outside the kernel's plan (H != 256, very wide
        // inputs / outputs) take the FP32 FFMA kernel: the same accuracy class, so the mode never changes meaning.
        rc = 1;
        if (tc3x_supported(args.net[0])) {
            for (int r = 0; r < args.n_roles; ++r) {
                rc = ensure_pack(args.net[r], PACK_TC3X, st);
                if (rc != FRL_OK) return rc;
            }
            rc = launch_logprob_tc3x(args, st);
        }
        if (rc > 0) rc = launch_logprob_f32(args, st);
    } else if (precision == FRL_PREC_BF16 || precision == FRL_PREC_FP16) {
        rc = launch_logprob_tc(args, precision, st);
    } else {
        set_error("invalid argument: unknown precision");
        return FRL_ERR_INVALID;
    }
    if (rc != FRL_OK) return rc;
    for (int r = 0; r < args.n_roles; ++r) {
        rc = note_reader(args.net[r], st);
        if (rc != FRL_OK) return rc;
    }
    if (args.out_reward && args.reward_type >= 0)
        rc = launch_reward(args.out_logp[0], args.out_logp[1], args.out_reward, args.B, args.reward_type, st);
    return rc;
}

// ---- host staging for frl_score_host -------------------------------------------------------------------------
struct HostStage {
    static constexpr int kSlots = 2;
    cudaStream_t stream[kSlots] = {nullptr, nullptr};
    float* dev[kSlots] = {nullptr, nullptr};
    size_t cap_floats = 0;
};

}  // namespace

void frl_host_stage_free(frl_net* net) {
    HostStage* hs = static_cast<HostStage*>(net->host_stage);
    if (!hs) return;
    for (int i = 0; i < HostStage::kSlots; ++i) {
        if (hs->dev[i]) cudaFree(hs->dev[i]);
        if (hs->stream[i]) cudaStreamDestroy(hs->stream[i]);
    }
    delete hs;
    net->host_stage = nullptr;
}

extern "C" int frl_vfield(const frl_net* net, const float* s, const float* a, const float* t, float* out0, float* out1,
                          long long B, int precision, void* stream) {
    int rc = check_net(net);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(s && a && t && out0, "null pointer");
    FRL_REQUIRE(precision == FRL_PREC_FP32, "frl_vfield: only FRL_PREC_FP32 is implemented");
    if (B <= 0) return FRL_OK;
    FRL_CUDA(cudaSetDevice(net->cfg.device));
    rc = launch_vfield_f32(net, s, a, t, out0, out1, B, (cudaStream_t)stream);
    return rc != FRL_OK ? rc : note_reader(net, (cudaStream_t)stream);
}

extern "C" int frl_logprob(const frl_net* net, float role_sign, const float* s, const float* a, const float* probe,
                           int probe_mode, const float* t_mid_host, int n_steps, float* out_logp, long long B,
                           int precision, void* stream) {
    int rc = check_net(net);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(out_logp != nullptr, "null output");
    FRL_CUDA(cudaSetDevice(net->cfg.device));
    ScoreArgs args{};
    args.net[0] = net;
    args.net[1] = nullptr;
    args.sign[0] = role_sign >= 0.f ? 1.f : -1.f;
    args.n_roles = 1;
    args.s = s; args.a = a; args.probe = probe; args.probe_mode = probe_mode;
    args.t_mid_host = t_mid_host; args.n_steps = n_steps; args.reward_type = -1;
    args.out_logp[0] = out_logp; args.out_logp[1] = nullptr; args.out_reward = nullptr;
    args.B = B;
    return dispatch_score(args, precision, (cudaStream_t)stream);
}

static int score_device(const frl_net* ne, const frl_net* na, const float* s, const float* a, const float* probe,
                        int probe_mode, const float* t_mid_host, int n_steps, int reward_type, float* le, float* la,
                        float* rw, long long B, int precision, cudaStream_t st, float* scratch) {
    ScoreArgs args{};
    args.net[0] = ne; args.net[1] = na;
    args.sign[0] = 1.f; args.sign[1] = -1.f;
    args.n_roles = 2;
    args.s = s; args.a = a; args.probe = probe; args.probe_mode = probe_mode;
    args.t_mid_host = t_mid_host; args.n_steps = n_steps; args.reward_type = reward_type;
    args.out_logp[0] = le ? le : scratch;
    args.out_logp[1] = la ? la : scratch + B;
    args.out_reward = rw;
    args.B = B;
    return dispatch_score(args, precision, st);
}

extern "C" int frl_score(const frl_net* ne, const frl_net* na, const float* s, const float* a, const float* probe,
                         int probe_mode, const float* t_mid_host, int n_steps, int reward_type, float* le, float* la,
                         float* rw, long long B, int precision, void* stream) {
    int rc = check_pair(ne, na);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(reward_type >= 0 && reward_type <= 4, "bad reward_type");
    FRL_CUDA(cudaSetDevice(ne->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    float* scratch = nullptr;
    if ((!le || !la) && B > 0) FRL_CUDA(cudaMallocAsync((void**)&scratch, 2 * B * sizeof(float), st));
    rc = score_device(ne, na, s, a, probe, probe_mode, t_mid_host, n_steps, reward_type, le, la, rw, B, precision, st,
                      scratch);
    if (scratch) cudaFreeAsync(scratch, st);
    return rc;
}

extern "C" int frl_score_host(frl_net* ne, frl_net* na, const float* s_h, const float* a_h, const float* probe_h,
                              int probe_mode, const float* t_mid_host, int n_steps, int reward_type, float* le_h,
                              float* la_h, float* rw_h, long long B, int precision) {
    int rc = check_pair(ne, na);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(reward_type >= 0 && reward_type <= 4, "bad reward_type");
    FRL_REQUIRE(s_h && a_h, "null s/a");
    FRL_REQUIRE(probe_mode == FRL_PROBE_EXACT || probe_h, "null probe");
    FRL_REQUIRE(n_steps >= 1 && n_steps <= kMaxSteps, "n_steps must be in 1..256");
    if (B <= 0) return FRL_OK;
    FRL_CUDA(cudaSetDevice(ne->cfg.device));
    static const bool timing = getenv("FRL_HOST_TIMING") != nullptr;   // prints a host-clock breakdown of the call
    const auto t_entry = std::chrono::steady_clock::now();
    FRL_CUDA(cudaDeviceSynchronize());  // weight packs may still be in flight on the caller's streams
    const auto t_synced = std::chrono::steady_clock::now();
    const int S = ne->cfg.s_dim, A = ne->cfg.a_dim;
    const long long chunk = std::min<long long>(B, 262144);
    const int probe_rows = probe_mode == FRL_PROBE_PER_STEP ? n_steps : (probe_mode == FRL_PROBE_SHARED ? 1 : 0);
    const size_t per_row = (size_t)S + A + (size_t)probe_rows * A + 3;
    const size_t need = per_row * (size_t)chunk;

    HostStage* hs = static_cast<HostStage*>(ne->host_stage);
    if (!hs) {
        hs = new (std::nothrow) HostStage();
        if (!hs) { set_error("out of host memory"); return FRL_ERR_NOMEM; }
        ne->host_stage = hs;
        for (int i = 0; i < HostStage::kSlots; ++i) FRL_CUDA(cudaStreamCreateWithFlags(&hs->stream[i], cudaStreamNonBlocking));
    }
    if (hs->cap_floats < need) {
        for (int i = 0; i < HostStage::kSlots; ++i) {
            if (hs->dev[i]) FRL_CUDA(cudaFree(hs->dev[i]));
            hs->dev[i] = nullptr;
        }
        hs->cap_floats = 0;
        for (int i = 0; i < HostStage::kSlots; ++i) FRL_CUDA(cudaMalloc(&hs->dev[i], need * sizeof(float)));
        hs->cap_floats = need;
    }
    int slot = 0;
    for (long long r0 = 0; r0 < B; r0 += chunk, slot ^= 1) {
        const long long n = std::min(chunk, B - r0);
        cudaStream_t st = hs->stream[slot];
        float* d_s = hs->dev[slot];
        float* d_a = d_s + (size_t)n * S;
        float* d_p = d_a + (size_t)n * A;
        float* d_le = d_p + (size_t)n * A * probe_rows;
        float* d_la = d_le + n;
        float* d_rw = d_la + n;
        FRL_CUDA(cudaMemcpyAsync(d_s, s_h + r0 * S, (size_t)n * S * sizeof(float), cudaMemcpyHostToDevice, st));
        FRL_CUDA(cudaMemcpyAsync(d_a, a_h + r0 * A, (size_t)n * A * sizeof(float), cudaMemcpyHostToDevice, st));
        if (probe_rows == 1) {
            FRL_CUDA(cudaMemcpyAsync(d_p, probe_h + r0 * A, (size_t)n * A * sizeof(float), cudaMemcpyHostToDevice, st));
        } else if (probe_rows > 1) {
            FRL_CUDA(cudaMemcpy2DAsync(d_p, (size_t)n * A * sizeof(float), probe_h + r0 * A,
                                       (size_t)B * A * sizeof(float), (size_t)n * A * sizeof(float), n_steps,
                                       cudaMemcpyHostToDevice, st));
        }
        rc = score_device(ne, na, d_s, d_a, d_p, probe_mode, t_mid_host, n_steps, reward_type, d_le, d_la,
                          rw_h ? d_rw : nullptr, n, precision, st, nullptr);
        if (rc != FRL_OK) return rc;
        if (le_h) FRL_CUDA(cudaMemcpyAsync(le_h + r0, d_le, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (la_h) FRL_CUDA(cudaMemcpyAsync(la_h + r0, d_la, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (rw_h) FRL_CUDA(cudaMemcpyAsync(rw_h + r0, d_rw, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        // a slot is reused two chunks later: its stream order already serialises copy-in after the previous D2H
    }
    const auto t_enqueued = std::chrono::steady_clock::now();
    for (int i = 0; i < HostStage::kSlots; ++i) FRL_CUDA(cudaStreamSynchronize(hs->stream[i]));
    if (timing) {
        const auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "frl_score_host: device sync %.1f us, enqueue %.1f us, wait %.1f us\n", us(t_entry, t_synced),
                us(t_synced, t_enqueued), us(t_enqueued, std::chrono::steady_clock::now()));
    }
    return FRL_OK;
}

extern "C" int frl_cfm_loss_grad(frl_net* net, float role_sign, const float* s, const float* a0, const float* t,
                                 const float* noise, long long B, long long mean_divisor, float grad_scale,
                                 int accumulate, float* loss, float* grad, void* stream) {
    int rc = check_net(net);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(s && a0 && t && noise, "null input");
    FRL_REQUIRE(B > 0, "empty batch");
    FRL_CUDA(cudaSetDevice(net->cfg.device));
    rc = cfm_loss_grad_f32(net, role_sign >= 0.f ? 1.f : -1.f, s, a0, t, noise, B, mean_divisor, grad_scale,
                           accumulate, loss, grad, (cudaStream_t)stream);
    return rc != FRL_OK ? rc : note_reader(net, (cudaStream_t)stream);
}

extern "C" int frl_cfm_loss_grad_p(frl_net* net, float role_sign, const float* s, const float* a0, const float* t,
                                   const float* noise, long long B, long long mean_divisor, float grad_scale,
                                   int accumulate, float* loss, float* grad, int precision, void* stream) {
    if (precision == FRL_PREC_FP32)
        return frl_cfm_loss_grad(net, role_sign, s, a0, t, noise, B, mean_divisor, grad_scale, accumulate, loss, grad, stream);
    int rc = check_net(net);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(s && a0 && t && noise, "null input");
    FRL_REQUIRE(B > 0, "empty batch");
    if (precision != FRL_PREC_FP16) {
        set_error("frl_cfm_loss_grad_p: the tensor-core training path computes with fp16 operands (FRL_PREC_FP16)");
        return FRL_ERR_UNSUPPORTED;
    }
    if (net->n_out_p > 64) {
        set_error("tensor-core training is not available for this net (a_dim too large)");
        return FRL_ERR_UNSUPPORTED;
    }
    FRL_CUDA(cudaSetDevice(net->cfg.device));
    const CfmTerm term{role_sign >= 0.f ? 1.f : -1.f, s, a0, t, noise, B, mean_divisor, grad_scale, loss};
    rc = cfm_loss_grad_tc(net, &term, 1, accumulate, grad, (cudaStream_t)stream);
    return rc != FRL_OK ? rc : note_reader(net, (cudaStream_t)stream);
}

extern "C" int frl_cfm_loss_grad_pair(frl_net* net, const frl_cfm_term* terms, int accumulate, float* grad, int precision,
                                      void* stream) {
    FRL_REQUIRE(terms != nullptr, "null terms");
    if (precision == FRL_PREC_FP32) {   // FFMA path: one pass per term
        for (int i = 0; i < 2; ++i) {
            int rc = frl_cfm_loss_grad(net, terms[i].role_sign, terms[i].s_dev, terms[i].a0_dev, terms[i].t_dev, terms[i].noise_dev,
                                       terms[i].B, terms[i].mean_divisor, terms[i].grad_scale, i == 0 ? accumulate : 1,
                                       terms[i].loss_dev, grad, stream);
            if (rc != FRL_OK) return rc;
        }
        return FRL_OK;
    }
    int rc = check_net(net);
    if (rc != FRL_OK) return rc;
    if (precision != FRL_PREC_FP16) {
        set_error("frl_cfm_loss_grad_pair: the tensor-core training path computes with fp16 operands (FRL_PREC_FP16)");
        return FRL_ERR_UNSUPPORTED;
    }
    if (net->n_out_p > 64) {
        set_error("tensor-core training is not available for this net (a_dim too large)");
        return FRL_ERR_UNSUPPORTED;
    }
    CfmTerm t2[2];
    double scale[2];
    for (int i = 0; i < 2; ++i) {
        scale[i] = (double)terms[i].grad_scale / (double)(terms[i].mean_divisor > 0 ? terms[i].mean_divisor : std::max(1ll, terms[i].B));
    }
    if (scale[0] != 0.0 && scale[1] != 0.0 && (std::abs(scale[1] / scale[0]) > 256.0 || std::abs(scale[1] / scale[0]) < 1.0 / 256.0)) {
        // very unequal per-row scales would cost fp16 range in the stacked pass: one pass per term instead
        for (int i = 0; i < 2; ++i) {
            rc = frl_cfm_loss_grad_p(net, terms[i].role_sign, terms[i].s_dev, terms[i].a0_dev, terms[i].t_dev, terms[i].noise_dev,
                                     terms[i].B, terms[i].mean_divisor, terms[i].grad_scale, i == 0 ? accumulate : 1,
                                     terms[i].loss_dev, grad, precision, stream);
            if (rc != FRL_OK) return rc;
        }
        return FRL_OK;
    }
    for (int i = 0; i < 2; ++i) {
        FRL_REQUIRE(terms[i].s_dev && terms[i].a0_dev && terms[i].t_dev && terms[i].noise_dev, "null input");
        FRL_REQUIRE(terms[i].B > 0, "empty batch");
        t2[i] = CfmTerm{terms[i].role_sign >= 0.f ? 1.f : -1.f, terms[i].s_dev, terms[i].a0_dev, terms[i].t_dev, terms[i].noise_dev,
                        terms[i].B, terms[i].mean_divisor, terms[i].grad_scale, terms[i].loss_dev};
    }
    FRL_CUDA(cudaSetDevice(net->cfg.device));
    rc = cfm_loss_grad_tc(net, t2, 2, accumulate, grad, (cudaStream_t)stream);
    return rc != FRL_OK ? rc : note_reader(net, (cudaStream_t)stream);
}

extern "C" long long frl_train_timeouts(const frl_net* net) { return net ? train_timeouts(net) : -1; }

extern "C" int frl_reg_loss_grad(frl_net* net, const float* s, const float* a, const float* t, const float* probe,
                                 long long N, long long mean_divisor, int which, float* loss_anti, float* loss_stable,
                                 float* grad_anti, float* grad_stable, void* stream) {
    int rc = check_net(net);
    if (rc != FRL_OK) return rc;
    FRL_REQUIRE(net->cfg.n_heads == 2, "the regularisers are defined for the coupled-residual net (n_heads == 2) only");
    FRL_REQUIRE(s && a && t && probe, "null input");
    FRL_REQUIRE(N > 0, "empty batch");
    FRL_REQUIRE(which >= 1 && which <= 3, "which: bit 0 = anti, bit 1 = stable");
    FRL_CUDA(cudaSetDevice(net->cfg.device));
    rc = reg_loss_grad_f32(net, s, a, t, probe, N, mean_divisor, which, loss_anti, loss_stable, grad_anti,
                           grad_stable, (cudaStream_t)stream);
    return rc != FRL_OK ? rc : note_reader(net, (cudaStream_t)stream);
}

extern "C" int frl_grad_axpy(float* y, const float* x, long long n, float alpha, const float* num, const float* den,
                             float eps, const float* gate, float* coef_out, int device, void* stream) {
    FRL_REQUIRE(n >= 0 && (n == 0 || (x && y)), "null vector");
    FRL_REQUIRE(num || !den, "den without num");
    FRL_CUDA(cudaSetDevice(device));
    return grad_axpy(y, x, n, alpha, num, den, eps, gate, coef_out, (cudaStream_t)stream);
}

extern "C" int frl_normalise_rewards(const float* reward, const float* masks, int T, int N, float gamma,
                                     float* returns_io, int returns_valid, float* rms_mean_var, double* rms_count,
                                     float* out, int device, void* stream) {
    FRL_REQUIRE(reward && masks && returns_io && rms_mean_var && rms_count && out, "null pointer");
    FRL_REQUIRE(T >= 0 && N >= 1, "T >= 0 and N >= 1");
    if (T == 0) return FRL_OK;
    FRL_CUDA(cudaSetDevice(device));
    return normalise_returns(reward, masks, T, N, gamma, returns_io, returns_valid, rms_mean_var, rms_count, out,
                             (cudaStream_t)stream);
}

extern "C" int frl_tensor_norms(const float* flat, const long long* offsets_host, int n_tensors, float* out,
                                int device, void* stream) {
    FRL_REQUIRE(flat && offsets_host && out, "null pointer");
    FRL_CUDA(cudaSetDevice(device));
    return tensor_norms(flat, offsets_host, n_tensors, out, (cudaStream_t)stream);
}

extern "C" int frl_clip_adam(const frl_adam_segment* segs, int n_segs, long long step, double lr, double beta1,
                             double beta2, double eps, double max_norm, float* norm_out, int device, void* stream) {
    FRL_REQUIRE(segs && n_segs >= 1 && n_segs <= 4, "1..4 segments");
    FRL_REQUIRE(step >= 1, "step is the 1-based count after increment");
    for (int i = 0; i < n_segs; ++i)
        FRL_REQUIRE(segs[i].n >= 0 && (segs[i].n == 0 || (segs[i].params && segs[i].grad && segs[i].exp_avg &&
                                                          segs[i].exp_avg_sq)), "null segment pointer");
    FRL_CUDA(cudaSetDevice(device));
    return clip_adam(segs, n_segs, step, nullptr, lr, beta1, beta2, eps, max_norm, norm_out, (cudaStream_t)stream);
}

extern "C" int frl_clip_adam_dev(const frl_adam_segment* segs, int n_segs, long long* step_dev, double lr, double beta1,
                                 double beta2, double eps, double max_norm, float* norm_out, int device, void* stream) {
    FRL_REQUIRE(segs && n_segs >= 1 && n_segs <= 4, "1..4 segments");
    FRL_REQUIRE(step_dev != nullptr, "null step counter");
    for (int i = 0; i < n_segs; ++i)
        FRL_REQUIRE(segs[i].n >= 0 && (segs[i].n == 0 || (segs[i].params && segs[i].grad && segs[i].exp_avg &&
                                                          segs[i].exp_avg_sq)), "null segment pointer");
    FRL_CUDA(cudaSetDevice(device));
    return clip_adam(segs, n_segs, 1, step_dev, lr, beta1, beta2, eps, max_norm, norm_out, (cudaStream_t)stream);
}
