/*
 * flowril_b200.h -- C ABI of libflowril_b200.so: the B200 (sm_100a) implementation of the FLOWRIL/MODRIL
 * flow-matching occupancy-density hot path (velocity-field MLP, ODE log-density with divergence, coupled-residual
 * reward, conditional-flow-matching gradient, grad-clip + Adam).
 *
 * The reference (hhhhzl/MODRIL) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes for this path is a Python class contract.  Each entry point below is what a binding for
 * that contract calls, and cites the reference interface it replaces (paths relative to the reference checkout).
 * The reference-side binding (ctypes) is shown in INTEGRATION.md; modril_b200/_native.py is that binding.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.  All "dev" pointers are CUDA device pointers on the
 *     device the net was created on; "host" pointers are ordinary host memory.
 *   - every function returns 0 on success, a negative frl_status otherwise; frl_last_error() gives the message
 *     (thread-local).  Nothing throws, nothing calls exit().
 *   - work is enqueued on the caller's `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     with no hidden synchronisation, except the *_host entry points which return after their results are in
 *     host memory.
 *   - a frl_net is not thread-safe; distinct nets may be used from distinct threads.
 *   - matrices are row-major float32: s [B, s_dim], a [B, a_dim], probe [B, a_dim], t [B] ...
 *   - parameters travel as ONE flat float32 vector in the reference's state_dict order
 *     (modril/flowril/networks.py:13-21 SharedVNet, :37-44 ConditionalVNet):
 *        trunk layer l: weight [H, K_l] then bias [H]  (K_0 = a_dim + s_dim + 1, input columns = [a | s | t])
 *        n_heads == 2:  head_c.weight [a,H], head_c.bias [a], head_r.weight [a,H], head_r.bias [a]
 *        n_heads == 1:  out.weight [a,H], out.bias [a]
 */
#ifndef FLOWRIL_B200_H
#define FLOWRIL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FRL_ABI_VERSION 7
#define FRL_MAX_STEPS 256
#define FRL_MAX_DEPTH 8

typedef enum {
    FRL_OK = 0,
    FRL_ERR_INVALID = -1,     /* bad argument / unsupported dimensions */
    FRL_ERR_CUDA = -2,        /* a CUDA runtime call failed            */
    FRL_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve */
    FRL_ERR_NOMEM = -4
} frl_status;

/* arithmetic mode of the log-density / velocity kernels */
typedef enum {
    FRL_PREC_FP32 = 0,    /* FP32 FFMA on CUDA cores: the 1e-4 parity mode                               */
    FRL_PREC_BF16 = 1,    /* tcgen05 tensor cores, bf16 operands, fp32 accumulate, fp32 ODE state/trace  */
    FRL_PREC_FP16X3 = 2,  /* tcgen05 tensor cores, 3-term fp16 operand split (x_hi W_hi + x_lo W_hi + x_hi W_lo): fp32-grade
                             products, the 1e-4 parity mode on tensor cores (nets outside its plan run the FP32 kernel)  */
    FRL_PREC_BF16X3 = 2,  /* former name of FRL_PREC_FP16X3 (ABI <= 5 reserved it)                               */
    FRL_PREC_FP16 = 3     /* tcgen05 tensor cores, fp16 operands (11-bit significand), fp32 accumulate   */
} frl_precision;

/* divergence estimator of the log-density (modril/flowril/pretrain.py:88-92, :163-177) */
typedef enum {
    FRL_PROBE_SHARED = 0,   /* probe [B,a] reused at every Euler step (scrf: _rademacher is shape-deterministic) */
    FRL_PROBE_PER_STEP = 1, /* probe [n_steps,B,a], fresh per step (2fs: randint_like, pretrain.py:90)           */
    FRL_PROBE_EXACT = 2     /* exact Jacobian trace: a_dim basis tangents, probe pointer ignored                  */
} frl_probe_mode;

/* reward transform, modril/flowril/flowril.py:191-207 (--reward-type) */
typedef enum { FRL_REWARD_RAW = 0, FRL_REWARD_NORM = 1, FRL_REWARD_EXP = 2, FRL_REWARD_CLIP = 3,
               FRL_REWARD_SMOOTH = 4 } frl_reward_type;

typedef struct {
    int s_dim;     /* state dimension                                                            */
    int a_dim;     /* action dimension (the ODE state)                                           */
    int hidden;    /* trunk width H: 128 or 256                                                  */
    int depth;     /* number of hidden layers, 1..FRL_MAX_DEPTH (networks.py:12 max(1, depth))   */
    int n_heads;   /* 2 = SharedVNet (head_c + tanh(head_r)), 1 = ConditionalVNet                */
    int device;    /* CUDA device ordinal                                                        */
} frl_net_config;

typedef struct frl_net frl_net; /* opaque: packed device copies of one velocity net + workspaces */

int frl_abi_version(void);
const char* frl_last_error(void);

/* Replaces SharedVNet.__init__ / ConditionalVNet.__init__ + .to(device) (networks.py:10-21, :34-44;
 * constructed by CoupledResidualFM / FlowMatching at pretrain.py:80, :153 and flowril.py:42-46, :88-99). */
int frl_net_create(const frl_net_config* cfg, frl_net** out);
int frl_net_destroy(frl_net* net);
long long frl_net_num_params(const frl_net* net);

/* Declare the flat fp32 master copy (device pointer, state_dict order) of the net's weights, or that its contents
 * changed.  Replaces what load_state_dict / optimizer.step leave implicit in the reference (flowril.py:48-49, :330):
 * call after every change of the master weights.  Packing is lazy: each kernel family rebuilds its own layout
 * (transposed / padded fp32, 16-bit UMMA block streams, GEMM images) on its first use after this call, on the stream of
 * that use, so the buffer must stay valid (and unchanged, or re-declared) until then.  `stream` is unused. */
int frl_net_set_params(frl_net* net, const float* params_dev, void* stream);

/* SharedVNet.forward / ConditionalVNet.forward (networks.py:23-25, :46-49).
 * out0 [B,a] = v_c (n_heads 2) or v (n_heads 1); out1 [B,a] = tanh(head_r) (n_heads 2; may be NULL).
 * t_dev [B] per-row times. */
int frl_vfield(const frl_net* net, const float* s_dev, const float* a_dev, const float* t_dev,
               float* out0_dev, float* out1_dev, long long B, int precision, void* stream);

/* CoupledResidualFM.log_prob (pretrain.py:192-214) / FlowMatching.log_prob (:112-134) for one role.
 * role_sign: +1 expert (v_c + r), -1 agent (v_c - r) (pretrain.py:156-158); ignored when n_heads == 1.
 * t_mid_host [n_steps]: midpoint times 0.5*(g_k+g_{k+1}) of the caller's fp32 linspace (pretrain.py:122, :206);
 * delta = 1/n_steps.  out_logp_dev [B]. */
int frl_logprob(const frl_net* net, float role_sign, const float* s_dev, const float* a_dev,
                const float* probe_dev, int probe_mode, const float* t_mid_host, int n_steps,
                float* out_logp_dev, long long B, int precision, void* stream);

/* FlowMatchingEstimation._compute_flow_reward / _compute_disc_val (flowril.py:175-208, :352-365): both roles in
 * one launch.  net_expert scores role expert, net_agent role agent; pass the same net twice for scrf.
 * Any of out_logp_e / out_logp_a / out_reward may be NULL.  reward = transform(logp_e - logp_a). */
int frl_score(const frl_net* net_expert, const frl_net* net_agent, const float* s_dev, const float* a_dev,
              const float* probe_dev, int probe_mode, const float* t_mid_host, int n_steps, int reward_type,
              float* out_logp_e_dev, float* out_logp_a_dev, float* out_reward_dev, long long B, int precision,
              void* stream);

/* Same as frl_score with HOST buffers: chunks the batch, overlaps H2D / kernel / D2H on internal streams and
 * returns when the outputs are in host memory (the end-to-end call bench.py times as `e2e`). */
int frl_score_host(frl_net* net_expert, frl_net* net_agent, const float* s_host, const float* a_host,
                   const float* probe_host, int probe_mode, const float* t_mid_host, int n_steps, int reward_type,
                   float* out_logp_e_host, float* out_logp_a_host, float* out_reward_host, long long B,
                   int precision);

/* CoupledResidualFM.fm_loss / FlowMatching.fm_loss followed by backward (pretrain.py:179-190, :94-109;
 * flowril.py:317-322) with caller-supplied randomness: t_dev [B] in [eps, 1-eps], noise_dev [B,a] ~ N(0,1).
 * Computes L = mean_b (1-t_b)^1.5 sum_j (v_role - (noise - a0))^2, writes it to loss_dev[0] and ADDS
 * grad_scale * dL/dtheta into grad_dev (flat, state_dict order) when accumulate != 0, else overwrites.
 * (For 2fs the caller adds the dequantisation noise to a0 first, pretrain.py:98.)
 * Deterministic: fixed reduction order, no float atomics.  mean_divisor: the B of the mean (the global batch
 * when the rows are one data-parallel shard); <= 0 means B. */
int frl_cfm_loss_grad(frl_net* net, float role_sign, const float* s_dev, const float* a0_dev, const float* t_dev,
                      const float* noise_dev, long long B, long long mean_divisor, float grad_scale,
                      int accumulate, float* loss_dev, float* grad_dev, void* stream);

/* frl_cfm_loss_grad with a choice of arithmetic: FRL_PREC_FP32 (the call above) or FRL_PREC_FP16 -- the tcgen05
 * tensor-core path: fp16 operands (weights, activations and back-propagated gradients are rounded to fp16 between
 * layers; dL/dV is carried unscaled by 1/B so it stays O(1)), fp32 accumulation in TMEM, fp32 loss and gradient
 * reduction in a fixed order (deterministic).  Same arguments and semantics otherwise. */
int frl_cfm_loss_grad_p(frl_net* net, float role_sign, const float* s_dev, const float* a0_dev, const float* t_dev,
                        const float* noise_dev, long long B, long long mean_divisor, float grad_scale,
                        int accumulate, float* loss_dev, float* grad_dev, int precision, void* stream);

/* The two CFM terms of one reward-model update in ONE pass (flowril.py:244-253: loss = rate_E * fm_loss(expert batch)
 * + rate_A * fm_loss(agent batch), same net for the coupled-residual model): the two batches are stacked row-wise, so
 * every layer GEMM runs once over B_0 + B_1 rows (fewer, fuller waves than two passes).  terms[0..1] as in
 * frl_cfm_loss_grad_p (grad_scale = the term's rate); each loss_dev receives its own scalar; grad_dev receives the sum.
 * FRL_PREC_FP32 runs the two FFMA passes back to back (same results as two frl_cfm_loss_grad calls). */
typedef struct {
    float role_sign;
    const float* s_dev;
    const float* a0_dev;
    const float* t_dev;
    const float* noise_dev;
    long long B, mean_divisor;
    float grad_scale;
    float* loss_dev;
} frl_cfm_term;
int frl_cfm_loss_grad_pair(frl_net* net, const frl_cfm_term* terms, int accumulate, float* grad_dev, int precision,
                           void* stream);

/* Anti-divergence and Jacobian-stability regularisers of the coupled-residual (n_heads == 2) model with their
 * parameter gradients (flowril.py:255-264: `_hutch_div(mix_a, r, k=1).square().mean()` and
 * `_jacobian_frobenius(mix_a, v_c + r).mean()`, pretrain.py:163-177, :55-59; the reference differentiates them by
 * autograd double-backward inside loss.backward(), flowril.py:322, and GradNorm2D.update, networks.py:108-117).
 *   loss_anti   = mean_b (v^T J_r v)^2,  loss_stable = mean_b |J_{v_c+r}^T v|^2,  v = probe [N,a] (caller-supplied;
 *   the reference draws the same shape-deterministic `_rademacher` matrix for both), evaluated at (a, s, t).
 * which: bit 0 = anti, bit 1 = stable.  Each requested loss writes its scalar to loss_*_dev[0] (may be NULL) and
 * OVERWRITES grad_*_dev (flat, state_dict order, unscaled; may be NULL to skip the backward).  mean_divisor as in
 * frl_cfm_loss_grad.  Deterministic. */
int frl_reg_loss_grad(frl_net* net, const float* s_dev, const float* a_dev, const float* t_dev,
                      const float* probe_dev, long long N, long long mean_divisor, int which, float* loss_anti_dev,
                      float* loss_stable_dev, float* grad_anti_dev, float* grad_stable_dev, void* stream);

/* y += c * x over n floats with c = alpha * num / (den + eps)  (den NULL: alpha * num; num NULL: alpha), forced to 0
 * when gate_dev[0] <= 0.  num_dev / den_dev / gate_dev are optional DEVICE scalars, so the warm-up weights
 * w = L_E / (L_reg + 1e-8) * 1e-3 of flowril.py:272-284 and GradNorm2D's beta / gamma (networks.py:104-106) scale a
 * gradient without a host round trip; `gate` (the previous weight) reproduces `use_anti = args.loss_anti > 0`
 * (flowril.py:255-256).  coef_out_dev (may be NULL; TWO floats: [0] = c, [1] = scratch that must start as 0; may alias
 * gate_dev). */
int frl_grad_axpy(float* y_dev, const float* x_dev, long long n, float alpha, const float* num_dev,
                  const float* den_dev, float eps, const float* gate_dev, float* coef_out_dev, int device,
                  void* stream);

/* Return normalisation of FlowMatchingEstimation._get_reward (flowril.py:221-228) for a whole rollout in one launch:
 * for t in 0..T-1:  returns = returns * masks[t] * gamma + reward[t];  RunningMeanStd.update(returns)
 * (float32 moments, rlf/baselines/common/running_mean_std.py:3-35);  out[t] = reward[t] / sqrt(var + 1e-8).
 * reward_dev / masks_dev / out_dev are [T, N]; returns_io_dev [N] carries the discounted returns between calls
 * (returns_valid == 0 on the very first call: returns starts as reward[0], flowril.py:222-223);
 * rms_mean_var_dev [2] = {mean, var} (initially {0, 1}) and rms_count_dev [1] (double, initially 1e-4) carry the
 * running moments.  Replaces T device->host syncs per update with none. */
int frl_normalise_rewards(const float* reward_dev, const float* masks_dev, int T, int N, float gamma,
                          float* returns_io_dev, int returns_valid, float* rms_mean_var_dev, double* rms_count_dev,
                          float* out_dev, int device, void* stream);

/* L2 norm of each parameter tensor's slice of a flat gradient: out_dev[i] = |flat[offsets[i] .. offsets[i+1])|.
 * GradNorm2D.update sums `g.norm()` over the parameter tensors of each regulariser's gradient
 * (networks.py:108-117).  offsets_host has n_tensors + 1 entries (n_tensors <= 20). */
int frl_tensor_norms(const float* flat_dev, const long long* offsets_host, int n_tensors, float* out_dev, int device,
                     void* stream);

/* torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step (flowril.py:324-330; Adam defaults betas (0.9,0.999),
 * eps 1e-8, no weight decay) over up to 4 flat segments that share one global gradient norm.
 * step = 1-based step count after increment.  norm_out_dev[0] receives the pre-clip total norm (may be NULL). */
typedef struct {
    float* params;      /* [n] master weights, updated in place */
    const float* grad;  /* [n]                                  */
    float* exp_avg;     /* [n] Adam m                           */
    float* exp_avg_sq;  /* [n] Adam v                           */
    long long n;
} frl_adam_segment;
int frl_clip_adam(const frl_adam_segment* segs, int n_segs, long long step, double lr, double beta1, double beta2,
                  double eps, double max_norm, float* norm_out_dev, int device, void* stream);

/* frl_clip_adam with the step count held ON THE DEVICE (step_dev[0], int64; the number of completed steps on entry):
 * the call increments it and derives the bias corrections from it inside the kernels, so a captured CUDA graph of a
 * whole update (frl_cfm_loss_grad_p ... frl_clip_adam_dev, frl_net_set_params) replays correctly step after step. */
int frl_clip_adam_dev(const frl_adam_segment* segs, int n_segs, long long* step_dev, double lr, double beta1,
                      double beta2, double eps, double max_norm, float* norm_out_dev, int device, void* stream);

/* Diagnostics of the fused tensor-core update (frl_cfm_loss_grad_p / _pair with a gradient): its stages hand row tiles to
 * each other through flags in global memory, and every such wait is bounded (~2 s) so that a lost stage cannot hang the
 * GPU.  Returns how many waits have given up on this net so far (0 in a healthy process; > 0 means the gradients of
 * that update were garbage), or -1 on a CUDA error.  Synchronises the device. */
long long frl_train_timeouts(const frl_net* net);

/* ---- reward hook helpers ---------------------------------------------------------------------------------------------
 * frl_reward: the reward transform of _compute_flow_reward (flowril.py:191-208) on two log-density vectors that were
 * produced by separate launches (2fs: each net integrates with its own probes): reward = transform(logp_e - logp_a),
 * reward_type 0 raw, 1 norm, 2 exp, 3 clip, 4 smooth (same codes as frl_score).
 * frl_rademacher_calls: the Hutchinson probes FlowMatching._hutch_div (pretrain.py:88-92) would draw from torch's CUDA
 * generator over a whole rollout -- n_outer rollout steps x n_nets nets x n_steps Euler steps, each a
 * randint_like([numel], 0, 2) * 2 - 1 -- replayed bit for bit in one launch from the generator's (seed, offset):
 * out<net>_dev [n_steps][n_outer][numel].  The caller advances the generator offset by 4 * n_outer * n_nets * n_steps. */
int frl_reward(const float* logp_e_dev, const float* logp_a_dev, float* reward_dev, long long B, int reward_type, int device,
               void* stream);
int frl_rademacher_calls(unsigned long long seed, unsigned long long offset, int n_outer, int n_nets, int n_steps, int numel,
                         float* out0_dev, float* out1_dev, int device, void* stream);

/* ---- data-parallel gradient all-reduce over NVLink peer memory ------------------------------------------
 * Replaces what torch.distributed / DistributedDataParallel would do around the reference's update
 * (modril/flowril/flowril.py:311-330: backward -> [gradient all-reduce] -> clip_grad_norm_ -> Adam.step) by two ordinary
 * kernels on the caller's stream: graph-capturable, no host synchronisation, no collective library launch between the
 * gradient kernels and clip + Adam.  One process per GPU: every rank creates a comm (an exchange buffer of
 * max_floats fp32 values, double-buffered), passes its 64-byte handle to all ranks by any host channel
 * (torch.distributed.all_gather_object) and connects with the [world][64] array of all handles.  Within ONE process
 * driving several GPUs, frl_comm_connect_local takes the comm objects instead.
 *   frl_comm_allreduce: bufs[k][0..ns[k]) (k < n_bufs <= 8, fp32, device) <- sum over ranks, in place.  The sum runs in
 *   rank order on every rank, so all ranks hold bit-identical results (the same clip + Adam then keeps the weights
 *   identical without a broadcast).  Every rank must issue the same sequence of calls with the same sizes.
 *   frl_comm_publish / frl_comm_reduce: the two halves (publish my vector; wait for all peers and sum), for callers that
 *   place independent work between them.
 *   frl_comm_timeouts: how many waits gave up because a peer never arrived within ~10 s (0 in a healthy job; synchronises). */
typedef struct frl_comm frl_comm;
int frl_comm_create(int rank, int world, int device, long long max_floats, frl_comm** out);
int frl_comm_handle(const frl_comm* comm, void* out_handle_64_bytes);
int frl_comm_connect(frl_comm* comm, const void* all_handles_world_x_64_bytes);
int frl_comm_connect_local(frl_comm* comm, frl_comm* const* all_comms);
int frl_comm_publish(frl_comm* comm, float* const* bufs_dev, const long long* ns, int n_bufs, void* stream);
int frl_comm_reduce(frl_comm* comm, float* const* bufs_dev, const long long* ns, int n_bufs, void* stream);
int frl_comm_allreduce(frl_comm* comm, float* const* bufs_dev, const long long* ns, int n_bufs, void* stream);
long long frl_comm_timeouts(frl_comm* comm);
int frl_comm_destroy(frl_comm* comm);

/* ---- BCF flow-matching policy (behaviour cloning with a flow-matching actor; scope row f3) ------------------------
 * VectorNetwork (deps/rl-toolkit/rlf/rl/model.py:391-432): three embeddings + a two-layer velocity head,
 *   v(x, s, t) = W2 relu(W1 [relu(Ws s + bs) | relu(Wt2 relu(wt1 t + bt1) + bt2) | relu(Wa x + ba)] + b1) + b2.
 * Parameters are ONE flat FP32 vector in state_dict() order: time_embed.0.{weight [T,1], bias}, time_embed.2.{weight
 * [H,T], bias}, state_embed.0.{weight [H,S], bias}, action_embed.0.{weight [H,A], bias}, velocity_head.0.{weight
 * [H,3H], bias}, velocity_head.2.{weight [A,H], bias}.  hidden and time_dim must be multiples of 16. */
typedef struct {
    int s_dim, a_dim, hidden, time_dim; /* model.py:392 (state_dim, action_dim, hidden_dim=256, time_embed_dim=128) */
    int device;
} frl_vnet_config;
typedef struct frl_vnet frl_vnet;
int frl_vnet_create(const frl_vnet_config* cfg, frl_vnet** out);
int frl_vnet_destroy(frl_vnet* net);
long long frl_vnet_num_params(const frl_vnet* net);
/* Same contract as frl_net_set_params: params_dev stays owned by the caller, kernel-side layouts are rebuilt lazily. */
int frl_vnet_set_params(frl_vnet* net, const float* params_dev, void* stream);

/* VectorNetwork.forward(x, state, t) (model.py:422-432): x [B,A], s [B,S], t [B] -> v [B,A]. */
int frl_vnet_forward(frl_vnet* net, const float* x_dev, const float* s_dev, const float* t_dev, float* v_dev,
                     long long B, void* stream);

/* FlowPolicy.get_action's Euler sampler (rlf/policies/flow_policy.py:49-63): x <- x0 (the caller's N(0,1) draw,
 * flow_policy.py:53), then n_steps of x += v(x, s, i/n_steps) / n_steps; writes the action [B,A].  The state
 * embedding is computed once.  action_dev may alias x0_dev. */
int frl_vnet_sample(frl_vnet* net, const float* s_dev, const float* x0_dev, int n_steps, float* action_dev, long long B,
                    void* stream);

/* Loss and parameter gradient of BehaviorCloneFlowMatching._bc_step (rlf/algos/il/bcf.py:108-128):
 *   x_t = (1-t) x0 + t a;  flow = mse(v(x_t, s, t), a - x0);  act = mse(x0 + v(a, s, 1), a);  loss = flow + bc_coef act.
 * x0 [B,A] and t [B] are the caller's draws (bcf.py:109-110).  losses_dev (may be NULL) receives {loss, flow, act};
 * grad_dev (flat, state_dict order; may be NULL) receives grad_scale * dloss/dparams, overwritten or accumulated.
 * mean_divisor: rows the mean is taken over (0 = B; the global batch under data parallelism).  Deterministic. */
int frl_bcf_loss_grad(frl_vnet* net, const float* s_dev, const float* actions_dev, const float* x0_dev,
                      const float* t_dev, long long B, long long mean_divisor, float bc_coef, float grad_scale,
                      int accumulate, float* losses_dev, float* grad_dev, void* stream);

/* ---- PPO consumer of the rewards: returns / GAE and advantage normalisation (scope row f4, first part) -------------
 * RolloutStorage.compute_returns (deps/rl-toolkit/rlf/storage/rollout_storage.py:178-233) as ONE launch.
 * rewards [T,N] (value_dim copies of one reward, :179), value_preds [T+1,N,D], masks / bad_masks [T+1,N],
 * next_value [N,D], returns [T+1,N,D].  bad_masks == NULL <=> not args.use_proper_time_limits.  use_gae: writes
 * value_preds[T] = next_value (:185) and returns[0..T-1]; otherwise writes returns[T] = next_value (:201) and the
 * discounted sums.  Float32 operation order of the torch loop (gamma * gae_lambda formed in double like the python
 * floats): bit-identical results. */
int frl_compute_returns(const float* rewards_dev, float* value_preds_dev, const float* masks_dev,
                        const float* bad_masks_dev, const float* next_value_dev, int T, int N, int D, double gamma,
                        double gae_lambda, int use_gae, float* returns_dev, int device, void* stream);

/* RolloutStorage.compute_advantages (rollout_storage.py:235-239): adv = returns[:-1] - value_preds[:-1] over the first
 * n = T*N*D elements, then (adv - adv.mean()) / (adv.std() + eps) (unbiased std; the reference's eps is 1e-5).
 * scratch_dev: 256 doubles.  stats_out_dev (may be NULL) receives {mean, std}. */
int frl_normalise_advantages(const float* returns_dev, const float* value_preds_dev, long long n, float eps,
                             float* adv_dev, double* scratch_dev, float* stats_out_dev, int device, void* stream);

/* ---- PPO consumer, second part: the actor-critic policy and the clipped-PPO minibatch loss (scope row f4) ----------
 * The policy get_ppo_policy builds (deps/baselines/get_policy.py:12-23): DistActorCritic with an identity base net, a
 * critic MLPBasic(obs, hidden, layers) -> critic_head Linear(H,1) and an actor MLPBasic(obs, hidden, layers) ->
 * DiagGaussian(fc_mean Linear(H,A), state-independent logstd); tanh after EVERY trunk layer (rlf/rl/model.py:246-296).
 * Parameters are ONE flat FP32 vector in the policy's state_dict() order: critic.net.{0,2,..}.{weight,bias},
 * critic_head.{weight [1,H], bias [1]}, actor.net.{0,2,..}.{weight,bias}, dist.logstd [1,A], dist.fc_mean.{weight [A,H],
 * bias [A]}.  hidden must be a multiple of 16; 1 <= layers <= 8. */
typedef struct {
    int obs_dim, a_dim, hidden, layers; /* --ppo-hidden-dim (64), --ppo-layers (2): deps/baselines/main.py:106-107 */
    int device;
} frl_acnet_config;
typedef struct frl_acnet frl_acnet;
int frl_acnet_create(const frl_acnet_config* cfg, frl_acnet** out);
int frl_acnet_destroy(frl_acnet* net);
long long frl_acnet_num_params(const frl_acnet* net);
/* Same contract as frl_net_set_params: params_dev stays owned by the caller, kernel-side layouts are rebuilt lazily. */
int frl_acnet_set_params(frl_acnet* net, const float* params_dev, void* stream);

/* DistActorCritic.forward (rlf/policies/actor_critic/dist_actor_critic.py:63-74): obs [B,S] -> value [B] (may be NULL:
 * the critic is skipped) and the action mean [B,A] (may be NULL: the actor is skipped; get_value, base_actor_critic.py:52-56). */
int frl_acnet_forward(frl_acnet* net, const float* obs_dev, long long B, float* value_dev, float* mean_dev, void* stream);

/* get_action (dist_actor_critic.py:48-61): action = mean + exp(logstd) * noise with the caller's N(0,1) draw noise [B,A]
 * (NULL = dist.mode(), args.deterministic_policy), its log-prob [B] (FixedNormal.log_probs, rlf/rl/distributions.py:31-34)
 * and the entropy [B] (:36-37).  value / logp / ent may be NULL. */
int frl_acnet_act(frl_acnet* net, const float* obs_dev, const float* noise_dev, long long B, float* value_dev,
                  float* action_dev, float* logp_dev, float* ent_dev, void* stream);

/* evaluate_actions (dist_actor_critic.py:76-86): value [B], log-prob of the given actions [B], entropy [B]. */
int frl_acnet_evaluate(frl_acnet* net, const float* obs_dev, const float* action_dev, long long B, float* value_dev,
                       float* logp_dev, float* ent_dev, void* stream);

/* Loss and parameter gradient of one PPO.update minibatch (rlf/algos/on_policy/ppo.py:23-49):
 *   ratio = exp(logp - old_logp);  actor = -mean(min(ratio adv, clamp(ratio, 1 -+ clip) adv));
 *   value = 0.5 mean(max((v - ret)^2, (old_v + clamp(v - old_v, -+clip) - ret)^2));
 *   loss = value_loss_coef value + actor - entropy_coef mean(entropy).
 * idx_dev (int64 [B], may be NULL) gathers the minibatch rows out of the flattened rollout arrays (the BatchSampler
 * indices of RolloutStorage.feed_forward_generator, rollout_storage.py:270-323): obs [*,S], action [*,A], old_logp / adv /
 * ret / old_value [*].  losses_dev (may be NULL) receives {loss, value_loss, actor_loss, entropy}; grad_dev (flat,
 * state_dict order; may be NULL) receives grad_scale * dloss/dparams, overwritten or accumulated.  mean_divisor: rows
 * the means are taken over (0 = B; the global minibatch under data parallelism, losses are then this rank's share).
 * Ties of torch.min / torch.max split the gradient like autograd does.  Deterministic. */
int frl_ppo_loss_grad(frl_acnet* net, const float* obs_dev, const float* action_dev, const float* old_logp_dev,
                      const float* adv_dev, const float* ret_dev, const float* old_value_dev, const long long* idx_dev,
                      long long B, long long mean_divisor, float clip_param, float value_loss_coef, float entropy_coef,
                      float grad_scale, int accumulate, float* losses_dev, float* grad_dev, void* stream);

/* ---- offline pre-trainer: demonstration pre-processing (modril/flowril/pretrain.py:33-39, :269-285) -----------------
 * frl_column_moments: mean = x.mean(0), std = x.std(0) (unbiased) of x [n,d] -> mean_dev [d], std_dev [d].
 * frl_norm_vec: norm_vec(x, mean, std) = clamp((x - mean) / (std + 1e-8), -10, 10) -> out_dev [n,d] (may alias x_dev). */
int frl_column_moments(const float* x_dev, long long n, int d, float* mean_dev, float* std_dev, int device, void* stream);
int frl_norm_vec(const float* x_dev, const float* mean_dev, const float* std_dev, long long n, int d, float* out_dev,
                 int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWRIL_B200_H */
