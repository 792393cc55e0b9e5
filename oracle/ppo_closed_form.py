"""CPU oracle for the returns / GAE recursion and the advantage normalisation of the PPO consumer (scope row f4, first
part): numpy float32 restatement of ``RolloutStorage.compute_returns`` / ``compute_advantages``
(deps/rl-toolkit/rlf/storage/rollout_storage.py:178-239), one numpy op per torch op so the scan is bit-identical.

TEST INFRASTRUCTURE ONLY: nothing under ``modril_b200/`` imports this module.  Pinned to outputs of the reference's own
two methods by ``oracle/gen_golden.py::gen_gae`` (``tests/golden/gae_*.npz``).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def compute_returns(rewards, value_preds, masks, bad_masks, next_value, gamma, gae_lambda, use_gae,
                    use_proper_time_limits):
    """rewards [T,N,1], value_preds [T+1,N,D], masks / bad_masks [T+1,N,1], next_value [N,D] -> (returns [T+1,N,D],
    value_preds after the call).  Entries of ``returns`` the reference leaves untouched (returns[T] under GAE) are 0."""
    rewards, masks, bad_masks = (np.asarray(x, f32) for x in (rewards, masks, bad_masks))
    value_preds = np.array(value_preds, f32)
    T, D = rewards.shape[0], value_preds.shape[2]
    exp_rewards = np.repeat(rewards, D, axis=2)                       # :179
    returns = np.zeros_like(value_preds)
    g, gl = f32(gamma), f32(gamma * gae_lambda)
    if use_gae:
        value_preds[-1] = next_value                                   # :185 / :210
        gae = np.zeros_like(value_preds[0])
        for step in reversed(range(T)):
            delta = exp_rewards[step] + g * value_preds[step + 1] * masks[step + 1] - value_preds[step]
            gae = delta + gl * masks[step + 1] * gae
            if use_proper_time_limits:
                gae = gae * bad_masks[step + 1]                        # :198
            returns[step] = gae + value_preds[step]
    else:
        returns[-1] = next_value                                       # :201 / :226
        for step in reversed(range(T)):
            r = returns[step + 1] * g * masks[step + 1] + exp_rewards[step]
            if use_proper_time_limits:                                 # :205-211
                r = r * bad_masks[step + 1] + (f32(1) - bad_masks[step + 1]) * value_preds[step]
            returns[step] = r
    return returns, value_preds


def compute_advantages(returns, value_preds, eps=1e-5):
    """:235-239 -- (adv - mean) / (std + 1e-5) with the unbiased std; moments in float64 (torch's float32 reductions
    agree to an ulp or two, the test tolerance says so)."""
    adv = np.asarray(returns, f32)[:-1] - np.asarray(value_preds, f32)[:-1]
    mean = f32(adv.astype(np.float64).mean())
    std = f32(adv.astype(np.float64).std(ddof=1))
    return (adv - mean) / (std + f32(eps))


# ---- actor-critic policy and the clipped-PPO minibatch update (scope row f4, second part) -----------------------------
# numpy restatement (float64 by default) of the policy the reference builds for --alg flowril
# (deps/baselines/get_policy.py:12-23: DistActorCritic with MLPBasic actor / critic of args.ppo_hidden_dim x
# args.ppo_layers, tanh after EVERY layer, deps/rl-toolkit/rlf/rl/model.py:246-296; critic_head Linear(H,1),
# rlf/policies/utils.py:16-19; DiagGaussian fc_mean + state-independent logstd, rlf/rl/distributions.py:60-75) and of
# PPO.update's loss (rlf/algos/on_policy/ppo.py:17-49) with a HAND-WRITTEN backward (no autograd).  Pinned to outputs of
# the reference's own MLPBasic / DiagGaussian classes by oracle/gen_golden.py::gen_ppo (tests/golden/ppo_*.npz).
#
# Flat parameter order = state_dict order of the reference policy (critic and critic_head are registered by
# ActorCritic.init, actor and dist by DistActorCritic.init; a module's own parameters come before its children's):
#   critic.net.{0,2,..}.{weight,bias}, critic_head.{weight [1,H], bias [1]}, actor.net.{0,2,..}.{weight,bias},
#   dist.logstd [1,A], dist.fc_mean.{weight [A,H], bias [A]}
from dataclasses import dataclass

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


@dataclass(frozen=True)
class ACSpec:
    obs_dim: int
    a_dim: int
    hidden: int = 64
    layers: int = 2

    def shapes(self):
        H, S, A, L = self.hidden, self.obs_dim, self.a_dim, self.layers
        out = []
        for net in ("critic", "actor"):
            for l in range(L):
                out += [(f"{net}.net.{2 * l}.weight", (H, S if l == 0 else H)), (f"{net}.net.{2 * l}.bias", (H,))]
            if net == "critic":
                out += [("critic_head.weight", (1, H)), ("critic_head.bias", (1,))]
        out += [("dist.logstd", (1, A)), ("dist.fc_mean.weight", (A, H)), ("dist.fc_mean.bias", (A,))]
        return out

    def n_params(self):
        return sum(int(np.prod(s)) for _, s in self.shapes())


def ac_unpack(spec: ACSpec, flat):
    out, off = {}, 0
    for name, shape in spec.shapes():
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    assert off == flat.size
    return out


def ac_make_params(spec: ACSpec, seed: int, dtype=np.float32):
    """Deterministic generic weights U(+-1/sqrt(fan_in)), non-zero biases and logstd (a trained policy has neither the
    orthogonal init nor the zero biases of model.py:22-25)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts, fan = [], 1
    for name, shape in spec.shapes():
        if name == "dist.logstd":
            parts.append(rng.uniform(-0.7, 0.3, size=shape).astype(dtype).ravel())
            continue
        if len(shape) == 2:
            fan = shape[1]
        b = 1.0 / np.sqrt(fan)
        parts.append(rng.uniform(-b, b, size=shape).astype(dtype).ravel())
    return np.concatenate(parts)


def _mlp_tanh(P, net, L, x):
    acts = [x]
    for l in range(L):
        acts.append(np.tanh(acts[-1] @ P[f"{net}.net.{2 * l}.weight"].T + P[f"{net}.net.{2 * l}.bias"]))
    return acts


def ac_forward(spec: ACSpec, flat, obs, dtype=np.float64):
    """DistActorCritic.forward (rlf/policies/actor_critic/dist_actor_critic.py:63-74) with the identity base net:
    -> value [B,1], action mean [B,A], (critic activations, actor activations)."""
    P = ac_unpack(spec, np.asarray(flat, dtype))
    obs = np.asarray(obs, dtype)
    hc = _mlp_tanh(P, "critic", spec.layers, obs)
    ha = _mlp_tanh(P, "actor", spec.layers, obs)
    value = hc[-1] @ P["critic_head.weight"].T + P["critic_head.bias"]
    mean = ha[-1] @ P["dist.fc_mean.weight"].T + P["dist.fc_mean.bias"]
    return value, mean, (hc, ha)


def ac_evaluate(spec: ACSpec, flat, obs, action, dtype=np.float64):
    """evaluate_actions (dist_actor_critic.py:76-86): value [B,1], FixedNormal.log_probs [B,1] and entropy [B]
    (distributions.py:26-38; torch's Normal.log_prob / entropy formulas)."""
    P = ac_unpack(spec, np.asarray(flat, dtype))
    value, mean, _ = ac_forward(spec, flat, obs, dtype)
    logstd = P["dist.logstd"]
    var = np.exp(logstd) ** 2
    logp = (-((np.asarray(action, dtype) - mean) ** 2) / (2 * var) - logstd - LOG_SQRT_2PI).sum(-1, keepdims=True)
    ent = np.broadcast_to(0.5 + LOG_SQRT_2PI + logstd, mean.shape).sum(-1)
    return value, logp, ent


def ac_act(spec: ACSpec, flat, obs, noise, dtype=np.float64):
    """get_action (dist_actor_critic.py:48-61) with the N(0,1) draw of Normal.sample injected (noise None = dist.mode()):
    -> value, action = mean + std * noise, log_prob [B,1]."""
    P = ac_unpack(spec, np.asarray(flat, dtype))
    value, mean, _ = ac_forward(spec, flat, obs, dtype)
    std = np.exp(P["dist.logstd"])
    action = mean if noise is None else mean + std * np.asarray(noise, dtype)
    logp = (-((action - mean) ** 2) / (2 * std ** 2) - P["dist.logstd"] - LOG_SQRT_2PI).sum(-1, keepdims=True)
    return value, action, logp


def ppo_loss_and_grad(spec: ACSpec, flat, obs, action, old_logp, adv, ret, old_value, clip=0.2, value_coef=0.5,
                      ent_coef=0.01, dtype=np.float64, need_grad=True):
    """ppo.py:23-49 -> ([loss, value_loss, actor_loss, entropy], flat gradient).  torch.min / torch.max split the
    gradient evenly on ties, clamp passes it on the closed interval; the backward below follows those rules."""
    flat = np.asarray(flat, dtype)
    P = ac_unpack(spec, flat)
    obs, action, old_logp, adv, ret, old_value = (np.asarray(x, dtype) for x in (obs, action, old_logp, adv, ret, old_value))
    B, L = obs.shape[0], spec.layers
    value, mean, (hc, ha) = ac_forward(spec, flat, obs, dtype)
    logstd = P["dist.logstd"]
    var = np.exp(logstd) ** 2
    diff = action - mean
    logp = (-(diff ** 2) / (2 * var) - logstd - LOG_SQRT_2PI).sum(-1, keepdims=True)
    ent = np.broadcast_to(0.5 + LOG_SQRT_2PI + logstd, mean.shape).sum(-1)
    ratio = np.exp(logp - old_logp)                                              # :29
    surr1 = ratio * adv
    clamped = np.clip(ratio, 1.0 - clip, 1.0 + clip)
    surr2 = clamped * adv
    actor_loss = -np.minimum(surr1, surr2).mean()                                # :34
    dv = value - old_value
    vpc = old_value + np.clip(dv, -clip, clip)                                   # :37-39
    vl, vlc = (value - ret) ** 2, (vpc - ret) ** 2
    value_loss = 0.5 * np.maximum(vl, vlc).mean()                                # :43-44
    loss = value_loss * value_coef + actor_loss - ent.mean() * ent_coef          # :48-49
    losses = np.array([loss, value_loss, actor_loss, ent.mean()], dtype)
    if not need_grad:
        return losses, None

    w1 = np.where(surr1 < surr2, 1.0, np.where(surr1 > surr2, 0.0, 0.5))
    in_r = ((ratio >= 1.0 - clip) & (ratio <= 1.0 + clip)).astype(dtype)
    d_ratio = -(adv / B) * (w1 + (1.0 - w1) * in_r)
    d_logp = d_ratio * ratio                                                     # [B,1]
    wv = np.where(vl > vlc, 1.0, np.where(vl < vlc, 0.0, 0.5))
    in_v = ((dv >= -clip) & (dv <= clip)).astype(dtype)
    d_value = (0.5 * value_coef / B) * (wv * 2 * (value - ret) + (1.0 - wv) * 2 * (vpc - ret) * in_v)
    d_mean = d_logp * diff / var                                                 # [B,A]
    d_logstd = (d_logp * (diff ** 2 / var - 1.0)).sum(0, keepdims=True) - ent_coef

    G = {k: np.zeros_like(v) for k, v in P.items()}
    G["dist.logstd"] = d_logstd

    def back(net, acts, d_out, head_w, head_b):
        G[head_w] = d_out.T @ acts[-1]
        G[head_b] = d_out.sum(0)
        d_h = d_out @ P[head_w]
        for l in reversed(range(L)):
            d_z = d_h * (1.0 - acts[l + 1] ** 2)
            G[f"{net}.net.{2 * l}.weight"] = d_z.T @ acts[l]
            G[f"{net}.net.{2 * l}.bias"] = d_z.sum(0)
            d_h = d_z @ P[f"{net}.net.{2 * l}.weight"]

    back("critic", hc, d_value, "critic_head.weight", "critic_head.bias")
    back("actor", ha, d_mean, "dist.fc_mean.weight", "dist.fc_mean.bias")
    return losses, np.concatenate([G[name].ravel() for name, _ in spec.shapes()])


def ppo_step(spec: ACSpec, flat, m, v, step, batch, lr, clip=0.2, value_coef=0.5, ent_coef=0.01, max_norm=0.5,
             eps=1e-12, dtype=np.float64):
    """One minibatch of PPO.update: the loss (ppo.py:23-49) then ``_standard_step`` (rlf/algos/base_net_algo.py:84-100:
    zero_grad, backward, clip_grad_norm_(max_grad_norm 0.5), Adam.step with eps 1e-12, :130-146).  ``batch`` =
    (obs, action, old_logp, adv, ret, old_value).  Returns (flat, m, v, losses[4], pre-clip grad norm)."""
    from oracle import closed_form as cf

    losses, grad = ppo_loss_and_grad(spec, flat, *batch, clip=clip, value_coef=value_coef, ent_coef=ent_coef, dtype=dtype)
    norm = 0.0
    if max_norm > 0:
        coef, norm = cf.clip_coef(grad, max_norm)
        grad = grad * coef
    flat, m, v = cf.adam_step(np.asarray(flat, dtype), grad, m, v, step, lr, eps=eps)
    return flat, m, v, losses, norm
