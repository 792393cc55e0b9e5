"""Generate tests/golden/*.npz by running the UNMODIFIED reference estimators on CPU (build container only).

    python -m oracle.gen_golden            # writes tests/golden/

TEST INFRASTRUCTURE.  The reference (``/root/reference``) does not travel to the GPU box, so its outputs on seeded
inputs are committed as fixtures.  Inputs and weights are regenerated from seeds by ``oracle/synth.py`` and
``oracle/closed_form.make_params`` (numpy PCG64, version-stable); only the reference's OUTPUTS are stored.

What is pinned (reference call sites in brackets):
  logprob_*.npz   CoupledResidualFM.log_prob / FlowMatching.log_prob [pretrain.py:192-214, :112-134] with the
                  probe injected through ``_rademacher`` / ``randint_like`` (ref_shim.injected_randomness), for the
                  task dims of SURVEY section 8, n_steps 20 and 32, plus the exact trace built from a_dim runs of
                  the reference with basis-vector probes.
  vfield_*.npz    CoupledResidualFM.v_field / ConditionalVNet.forward [pretrain.py:156-158, networks.py:23-25,46-49]
  fmloss_*.npz    fm_loss value and autograd gradients [pretrain.py:179-190, :94-109] with t / noise injected
  reg_*.npz       the anti-divergence / Jacobian-stability regularisers [flowril.py:257-263 -> pretrain.py:163-177,
                  :55-59] and their double-backward parameter gradients; trainreg_*.npz: the update loop with them.
  bcf_*.npz       VectorNetwork.forward, the BCF loss + gradients and the Euler action sampler [rlf/rl/model.py:391-432,
                  rlf/algos/il/bcf.py:108-128, rlf/policies/flow_policy.py:49-63]; bcftrain_*.npz: _bc_step updates.
  retnorm_*.npz   FlowMatchingEstimation._get_reward [flowril.py:210-230] with the reference RunningMeanStd
                  [rlf/baselines/common/running_mean_std.py:3-35]: the return normalisation of row a10.
  gae_*.npz       RolloutStorage.compute_returns / compute_advantages [rlf/storage/rollout_storage.py:178-239]
  train_*.npz     the update loop of flowril.py:311-335 (zero_grad, backward, clip_grad_norm_, Adam.step) for
                  scrf (all params), scrf fine-tune (head_r only, flowril.py:52-63) and 2fs; loss curves + final
                  parameters.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import closed_form as cf  # noqa: E402
from oracle import ref_shim, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_ref(P, spec: cf.NetSpec, flat: np.ndarray, eps=1e-3):
    import torch

    if spec.n_heads == 2:
        model = P.CoupledResidualFM(spec.s_dim, spec.a_dim, spec.hidden, depth=spec.depth, eps=eps)
    else:
        model = P.FlowMatching(spec.s_dim, spec.a_dim, spec.hidden, depth=spec.depth, eps=eps)
    sd = {k: torch.from_numpy(v.copy()) for k, v in cf.unpack(spec, flat).items()}
    assert list(sd.keys()) == list(model.state_dict().keys()), (list(sd.keys()), list(model.state_dict().keys()))
    model.load_state_dict(sd)
    return model


def ref_log_prob(P, model, spec, s, a, role, probe, n_steps):
    """probe [B,a] shared by all steps, or [n_steps,B,a]."""
    import torch

    st, at = torch.from_numpy(s), torch.from_numpy(a)
    probes = [torch.from_numpy(probe)] if probe.ndim == 2 else [torch.from_numpy(p) for p in probe]
    with ref_shim.injected_randomness(P, probes=probes):
        if spec.n_heads == 2:
            with torch.enable_grad():
                out = model.log_prob(st, at, role, n_steps=n_steps)
        else:
            out = model.log_prob(torch.cat([st, at], dim=-1), n_steps=n_steps)
    return out.detach().numpy().astype(np.float32)


def gen_logprob(P):
    cases = []
    for task in ("sine", "maze", "hopper", "halfcheetah", "ant", "hand", "pick"):
        for n_heads in (2, 1):
            for n_steps in ((20, 32) if task in ("maze", "sine") else (32,)):
                cases.append((task, n_heads, n_steps, 256, 4, 2.0))
    cases.append(("hopper", 2, 32, 128, 4, 1.0))
    cases.append(("hopper", 2, 8, 128, 2, 2.0))
    cases.append(("maze", 1, 32, 128, 3, 2.0))
    _gen_logprob_cases(P, cases, "", 1000)


def gen_logprob_smooth(P):
    """logprob_sNN_*.npz: the same quantities on nets at the scale of the reference's default init (gain 1) for every
    task dimension -- the nets on which the single-pass 16-bit tensor-core modes are held to north_star's 2e-3 bar
    (the gain-2 nets above amplify every operand rounding through the ReLU kinks)."""
    cases = [(task, 2, 20 if task == "maze" else 32, 256, 4, 1.0)
             for task in ("sine", "maze", "hopper", "halfcheetah", "ant", "hand", "pick")]
    cases += [("maze", 1, 20, 256, 4, 1.0), ("hopper", 1, 32, 256, 4, 1.0)]
    _gen_logprob_cases(P, cases, "s", 1500)


def _gen_logprob_cases(P, cases, tag, seed0):
    for idx, (task, n_heads, n_steps, hidden, depth, gain) in enumerate(cases):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = cf.NetSpec(s_dim, a_dim, hidden, depth, n_heads)
        B = 96
        seed = seed0 + idx
        out = {"meta": np.array([s_dim, a_dim, hidden, depth, n_heads, n_steps, B, seed]), "gain": np.float32(gain)}
        s, a = synth.states_actions(seed, B, s_dim, a_dim)
        probe = synth.rademacher(seed + 1, (B, a_dim))
        probe_steps = synth.rademacher(seed + 2, (n_steps, B, a_dim))
        roles = ("expert", "agent")
        for ri, role in enumerate(roles):
            # scrf: one net, two roles.  2fs: two independent nets (flowril.py:88-99), role only picks the seed.
            flat = cf.make_params(spec, seed + (10 * ri if n_heads == 1 else 0), gain)
            model = build_ref(P, spec, flat)
            out[f"logp_{role}"] = ref_log_prob(P, model, spec, s, a, role, probe, n_steps)
            out[f"logp_{role}_stepprobe"] = ref_log_prob(P, model, spec, s, a, role, probe_steps, n_steps)
            if a_dim <= 8:
                # exact trace from the unmodified reference: probe e_i gives prior(a_1) - delta*sum_k J_ii, and
                # the primal trajectory does not depend on the probe.
                runs = []
                for i in range(a_dim):
                    e = np.zeros((B, a_dim), np.float32)
                    e[:, i] = 1.0
                    runs.append(ref_log_prob(P, model, spec, s, a, role, e, n_steps).astype(np.float64))
                zero = ref_log_prob(P, model, spec, s, a, role, np.zeros((B, a_dim), np.float32), n_steps)
                out[f"logp_{role}_exact"] = (sum(runs) - (a_dim - 1) * zero.astype(np.float64)).astype(np.float32)
        name = f"logprob_{tag}{idx:02d}_{task}_{'scrf' if n_heads == 2 else '2fs'}_h{hidden}d{depth}_n{n_steps}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, out["logp_expert"][:3])


def gen_vfield(P):
    import torch

    for idx, (task, n_heads) in enumerate((("hopper", 2), ("maze", 1), ("ant", 2), ("sine", 2))):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = cf.NetSpec(s_dim, a_dim, 256, 4, n_heads)
        seed, B = 2000 + idx, 80
        flat = cf.make_params(spec, seed, 2.0)
        model = build_ref(P, spec, flat)
        s, a = synth.states_actions(seed, B, s_dim, a_dim)
        t = synth.rng(seed + 1).random((B, 1), dtype=np.float32)
        st, at, tt = map(torch.from_numpy, (s, a, t))
        out = {"meta": np.array([s_dim, a_dim, 256, 4, n_heads, 0, B, seed]), "gain": np.float32(2.0)}
        with torch.no_grad():
            if n_heads == 2:
                out["v_expert"] = model.v_field(at, st, tt, "expert").numpy()
                out["v_agent"] = model.v_field(at, st, tt, "agent").numpy()
                vc, r = model.net(at, st, tt)
                out["v_c"], out["r"] = vc.numpy(), r.numpy()
            else:
                out["v"] = model.net(at, st, tt).numpy()
        name = f"vfield_{idx:02d}_{task}_{'scrf' if n_heads == 2 else '2fs'}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name)


def ref_fm_loss(P, model, spec, s, a, role, u, noise, dequant=None):
    """u is the raw U[0,1) draw: the reference maps it to t = u*(1-2eps)+eps itself (pretrain.py:99, :182)."""
    import torch

    st, at = torch.from_numpy(s), torch.from_numpy(a)
    kw = dict(rand=[torch.from_numpy(u)], normal=[torch.from_numpy(noise)])
    if dequant is not None:
        # reference scales randn_like by dequant_std itself (pretrain.py:98): feed unit-scale noise
        kw["randn_like"] = [torch.from_numpy(dequant)]
    with ref_shim.injected_randomness(P, **kw):
        if spec.n_heads == 2:
            return model.fm_loss(st, at, role)
        return model.fm_loss(st, at)


def flat_grads(model):
    return np.concatenate([
        (p.grad if p.grad is not None else p.new_zeros(p.shape)).detach().numpy().ravel() for p in model.parameters()
    ]).astype(np.float32)


def gen_fmloss(P):
    cases = [("hopper", 2, 128, True), ("maze", 1, 128, True), ("sine", 2, 128, True),
             ("halfcheetah", 2, 256, False), ("ant", 2, 256, False), ("hand", 1, 256, False)]
    for idx, (task, n_heads, hidden, store_grad) in enumerate(cases):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = cf.NetSpec(s_dim, a_dim, hidden, 4, n_heads)
        seed, B = 3000 + idx, 200
        flat = cf.make_params(spec, seed, 2.0)
        s, a = synth.states_actions(seed, B, s_dim, a_dim)
        out = {"meta": np.array([s_dim, a_dim, hidden, 4, n_heads, 0, B, seed]), "gain": np.float32(2.0)}
        for ri, role in enumerate(("expert", "agent")):
            model = build_ref(P, spec, flat)
            _, noise, u = synth.cfm_draws(seed + 1 + ri, B, a_dim, return_u=True)
            dq = synth.rng(seed + 5 + ri).standard_normal((B, a_dim)).astype(np.float32) if n_heads == 1 else None
            loss = ref_fm_loss(P, model, spec, s, a, role, u, noise, dq)
            loss.backward()
            g = flat_grads(model)
            out[f"loss_{role}"] = np.float32(loss.item())
            out[f"gradnorm_{role}"] = np.float32(np.sqrt((g.astype(np.float64) ** 2).sum()))
            # per-tensor norms for every case; the full gradient only for the small nets
            norms, off = [], 0
            for _, shape in spec.shapes():
                n = int(np.prod(shape))
                norms.append(np.sqrt((g[off:off + n].astype(np.float64) ** 2).sum()))
                off += n
            out[f"tensor_gradnorms_{role}"] = np.array(norms, np.float32)
            if store_grad:
                out[f"grad_{role}"] = g
        name = f"fmloss_{idx:02d}_{task}_{'scrf' if n_heads == 2 else '2fs'}_h{hidden}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, out["loss_expert"], out["loss_agent"])


def gen_train(P, only=None):
    """flowril.py:311-335 for scrf (no regularisers: configs/sine/flowril-no-as.yaml), scrf fine-tune with frozen
    trunk + head_c (flowril.py:52-63) and 2fs both nets (flowril.py:108-109).  Role swap quirk
    lives above fm_loss and is not part of this fixture: batch 'E' is fed to role expert, batch 'A' to role agent.
    ``only``: indices of the cases to (re)generate (``python -m oracle.gen_golden train:4``)."""
    import torch

    for idx, (task, kind, hidden, steps, lr) in enumerate((
            ("hopper", "scrf", 128, 1000, 1e-3),
            ("hopper", "scrf_ft", 128, 50, 1e-3),
            ("maze", "2fs", 128, 50, 1e-3),
            ("sine", "scrf", 128, 200, 1e-3),
            ("maze", "scrf", 256, 1000, 1e-3),   # the module's default width (flowril.py:562) and the benchmarked net
    )):
        if only is not None and idx not in only:
            continue
        s_dim, a_dim = synth.TASK_DIMS[task]
        n_heads = 1 if kind == "2fs" else 2
        spec = cf.NetSpec(s_dim, a_dim, hidden, 4, n_heads)
        seed, B = 4000 + idx, 128
        models = [build_ref(P, spec, cf.make_params(spec, seed, 1.0))]
        if kind == "2fs":
            models.append(build_ref(P, spec, cf.make_params(spec, seed + 10, 1.0)))
        if kind == "scrf_ft":
            for p in models[0].net.shared.parameters():
                p.requires_grad = False
            for p in models[0].net.head_c.parameters():
                p.requires_grad = False
        all_params = [p for m in models for p in m.parameters()]
        opt = torch.optim.Adam([p for p in all_params if p.requires_grad], lr=lr)
        pool_s, _ = synth.states_actions(seed + 1, 4096, s_dim, a_dim)
        pool_aE = synth.expert_actions(seed + 2, pool_s, a_dim)
        pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, a_dim)), -1, 1).astype(np.float32)
        curve = np.zeros((steps, 2), np.float32)
        norms = np.zeros(steps, np.float32)
        for it in range(steps):
            g = synth.rng(seed + 100 + it)
            iE, iA = g.integers(0, 4096, B), g.integers(0, 4096, B)
            _, nE, tE = synth.cfm_draws(seed + 10000 + 2 * it, B, a_dim, return_u=True)  # tE/tA: raw uniforms
            _, nA, tA = synth.cfm_draws(seed + 10001 + 2 * it, B, a_dim, return_u=True)
            dqE = dqA = None
            if kind == "2fs":
                dqE = synth.rng(seed + 50000 + 2 * it).standard_normal((B, a_dim)).astype(np.float32)
                dqA = synth.rng(seed + 50001 + 2 * it).standard_normal((B, a_dim)).astype(np.float32)
            lE = ref_fm_loss(P, models[0], spec, pool_s[iE], pool_aE[iE], "expert", tE, nE, dqE)
            lA = ref_fm_loss(P, models[-1], spec, pool_s[iA], pool_aA[iA], "agent", tA, nA, dqA)
            loss = 1.0 * lE + 1.0 * lA  # flowril.py:253 with the default rates
            opt.zero_grad()
            loss.backward()
            norms[it] = float(torch.nn.utils.clip_grad_norm_(all_params, 1.0))  # flowril.py:324-329, disc_grad_pen=1
            opt.step()
            curve[it] = (lE.item(), lA.item())
        out = {"meta": np.array([s_dim, a_dim, hidden, 4, n_heads, steps, B, seed]), "lr": np.float32(lr),
               "curve": curve, "gradnorm": norms}
        for mi, m in enumerate(models):
            out[f"final_params_{mi}"] = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
        name = f"train_{idx:02d}_{task}_{kind}_h{hidden}_{steps}steps.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, curve[0], curve[-1])


def ref_reg_losses(P, model, s, a, t, probe):
    """loss_anti / loss_stable exactly as flowril.py:257-263 forms them (the probe is injected through
    `_rademacher`, which both `_hutch_div` and `_jacobian_frobenius` call)."""
    import torch

    st, tt = torch.from_numpy(s), torch.from_numpy(t)
    mix_a = torch.from_numpy(a).detach().requires_grad_(True)
    with ref_shim.injected_randomness(P, probes=[torch.from_numpy(probe)]):
        v_c, r = model.net(mix_a, st, tt)
        loss_anti = model._hutch_div(mix_a, r, k=1).square().mean()
        loss_stable = P._jacobian_frobenius(mix_a, v_c + r).mean()
    return loss_anti, loss_stable


def gen_reg(P):
    """reg_*.npz: the two regulariser losses and their autograd (double-backward) parameter gradients."""
    import torch

    cases = [("hopper", 128, True), ("sine", 128, True), ("maze", 256, False), ("halfcheetah", 256, False),
             ("ant", 256, False), ("hand", 256, False)]
    for idx, (task, hidden, store_grad) in enumerate(cases):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = cf.NetSpec(s_dim, a_dim, hidden, 4, 2)
        seed, N = 5000 + idx, 192
        flat = cf.make_params(spec, seed, 2.0)
        model = build_ref(P, spec, flat)
        s, a = synth.states_actions(seed, N, s_dim, a_dim)
        t = synth.rng(seed + 1).random((N, 1), dtype=np.float32)
        probe = synth.rademacher(seed + 2, (N, a_dim))
        la, ls = ref_reg_losses(P, model, s, a, t, probe)
        params = list(model.parameters())
        out = {"meta": np.array([s_dim, a_dim, hidden, 4, 2, 0, N, seed]), "gain": np.float32(2.0),
               "loss_anti": np.float32(la.item()), "loss_stable": np.float32(ls.item())}
        for name, loss in (("anti", la), ("stable", ls)):
            gs = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
            g = np.concatenate([(g_ if g_ is not None else torch.zeros_like(p_)).numpy().ravel()
                                for g_, p_ in zip(gs, params)]).astype(np.float32)
            out[f"normsum_{name}"] = np.float32(sum(g_.norm().item() for g_ in gs if g_ is not None))  # networks.py:111
            norms, off = [], 0
            for _, shape in spec.shapes():
                n = int(np.prod(shape))
                norms.append(np.sqrt((g[off:off + n].astype(np.float64) ** 2).sum()))
                off += n
            out[f"tensor_gradnorms_{name}"] = np.array(norms, np.float32)
            if store_grad:
                out[f"grad_{name}"] = g
        name = f"reg_{idx:02d}_{task}_scrf_h{hidden}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, out["loss_anti"], out["loss_stable"])


def gen_train_reg(P):
    """trainreg_*.npz: the scrf update of flowril.py:251-284 + :311-330 WITH the regularisers in the warm-up weighting
    (the only branch the reference's control flow reaches, see modril_b200/flowril/flowril.py) and with fixed
    weights (params_autotune false, flowril.py:266-270)."""
    import torch

    for idx, (task, hidden, steps, lr, mode) in enumerate((
            ("hopper", 128, 300, 1e-3, "warmup"),
            ("sine", 128, 100, 1e-3, "fixed"),
    )):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = cf.NetSpec(s_dim, a_dim, hidden, 4, 2)
        seed, B = 6000 + idx, 128
        model = build_ref(P, spec, cf.make_params(spec, seed, 1.0))
        params = list(model.parameters())
        opt = torch.optim.Adam(params, lr=lr)
        pool_s, _ = synth.states_actions(seed + 1, 4096, s_dim, a_dim)
        pool_aE = synth.expert_actions(seed + 2, pool_s, a_dim)
        pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, a_dim)), -1, 1).astype(np.float32)
        probe = synth.rademacher(seed + 4, (2 * B, a_dim))  # the reference's probe is the same matrix every step
        curve = np.zeros((steps, 4), np.float32)
        weights = np.zeros((steps, 2), np.float32)
        norms = np.zeros(steps, np.float32)
        w_fixed = (1e-4, 1e-6)  # --loss-anti / --loss-stable defaults (flowril.py:566-567)
        for it in range(steps):
            g = synth.rng(seed + 100 + it)
            iE, iA = g.integers(0, 4096, B), g.integers(0, 4096, B)
            _, nE, tE = synth.cfm_draws(seed + 10000 + 2 * it, B, a_dim, return_u=True)
            _, nA, tA = synth.cfm_draws(seed + 10001 + 2 * it, B, a_dim, return_u=True)
            t_mix = synth.rng(seed + 70000 + it).random((2 * B, 1), dtype=np.float32)
            lE = ref_fm_loss(P, model, spec, pool_s[iE], pool_aE[iE], "expert", tE, nE)
            lA = ref_fm_loss(P, model, spec, pool_s[iA], pool_aA[iA], "agent", tA, nA)
            flow_loss = 1.0 * lE + 1.0 * lA
            mix_s = np.concatenate([pool_s[iE], pool_s[iA]])
            mix_a = np.concatenate([pool_aE[iE], pool_aA[iA]])
            l_anti, l_st = ref_reg_losses(P, model, mix_s, mix_a, t_mix, probe)
            if mode == "warmup":  # flowril.py:272-284 with eps 1e-8, global_scale 1e-3 (flowril.py:118-119)
                w_anti = lE.detach() / (l_anti.detach() + 1e-8) * 1e-3
                w_st = lE.detach() / (l_st.detach() + 1e-8) * 1e-3
            else:
                w_anti, w_st = torch.tensor(w_fixed[0]), torch.tensor(w_fixed[1])
            flow_loss = flow_loss + w_anti * l_anti + w_st * l_st
            opt.zero_grad()
            flow_loss.backward()
            norms[it] = float(torch.nn.utils.clip_grad_norm_(params, 1.0))
            opt.step()
            curve[it] = (lE.item(), lA.item(), l_anti.item(), l_st.item())
            weights[it] = (float(w_anti), float(w_st))
        out = {"meta": np.array([s_dim, a_dim, hidden, 4, 2, steps, B, seed]), "lr": np.float32(lr), "curve": curve,
               "weights": weights, "gradnorm": norms, "mode": np.array(0 if mode == "warmup" else 1),
               "final_params_0": np.concatenate([p.detach().numpy().ravel() for p in params])}
        name = f"trainreg_{idx:02d}_{task}_scrf_{mode}_h{hidden}_{steps}steps.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, curve[0], curve[-1], weights[0], weights[-1])


def gen_gradnorm():
    """gradnorm_00.npz: trajectory of the reference's GradNorm2D (networks.py:52-138; the file needs only torch and
    math, so it is loaded directly) on a toy pair of losses whose parameter gradients are known in closed form."""
    import importlib.util

    import torch

    sp = importlib.util.spec_from_file_location("ref_networks",
                                                os.path.join(ref_shim.REFERENCE_ROOT, "modril/flowril/networks.py"))
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    gn = mod.GradNorm2D(beta_init=1e-4, gamma_init=1e-6, alpha=0.1, warmup_steps=20, update_freq=5, ema_decay=0.9,
                        beta_min=1e-8, beta_max=1e1, gamma_min=1e-12, gamma_max=1e-2)
    p1 = torch.nn.Parameter(torch.tensor([1.0, -2.0, 0.5]))
    p2 = torch.nn.Parameter(torch.tensor([[0.3, 0.7]]))
    g = synth.rng(42)
    steps = 120
    ca = g.uniform(0.5, 2.0, steps).astype(np.float32)
    cb = g.uniform(0.1, 5.0, steps).astype(np.float32)
    traj, regs = np.zeros((steps, 2)), np.zeros((steps, 3))
    for k in range(steps):
        la = float(ca[k]) * ((p1 ** 2).sum() + (p2 ** 2).sum())
        lj = float(cb[k]) * ((p1 ** 4).sum() + p2.abs().sum())
        reg, beta, gamma = gn(la, lj)
        gn.update(la, lj, [p1, p2])
        traj[k] = (float(gn.log_beta), float(gn.log_gamma))
        regs[k] = (float(reg), float(beta), float(gamma))
    np.savez_compressed(os.path.join(OUT, "gradnorm_00.npz"), ca=ca, cb=cb, traj=traj, regs=regs,
                        L0=np.array([float(gn.L0_anti), float(gn.L0_jpen)]),
                        ema=np.array([float(gn.ema_g_anti), float(gn.ema_g_jpen)]),
                        step_count=np.array(int(gn.step_count)))
    print("wrote gradnorm_00.npz", traj[-1])


def load_vector_network():
    """The reference's ``VectorNetwork`` class object.  ``rlf.rl.model`` imports the whole toolkit (gym, ...), which is
    not installed here, so the class and the two init helpers it uses are compiled from the reference source file as it
    lies under /root/reference (nothing is copied into the repo)."""
    import ast

    import torch
    import torch.nn as nn

    path = os.path.join(ref_shim.REFERENCE_ROOT, "deps", "rl-toolkit", "rlf", "rl", "model.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body
            if (isinstance(n, ast.ClassDef) and n.name == "VectorNetwork")
            or (isinstance(n, ast.FunctionDef) and n.name in ("weight_init", "no_bias_weight_init"))]
    assert len(keep) == 3, [getattr(n, "name", None) for n in keep]
    ns = {"torch": torch, "nn": nn, "np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["VectorNetwork"]


def build_vnet(VN, spec, flat, dtype=None):
    import torch

    m = VN(spec.s_dim, spec.a_dim, spec.hidden, spec.time_dim)
    sd = m.state_dict()
    assert [k for k in sd] == [n for n, _ in spec.shapes()], list(sd)
    off = 0
    for (name, shape) in spec.shapes():
        n = int(np.prod(shape))
        sd[name].copy_(torch.from_numpy(np.asarray(flat[off:off + n], np.float32).reshape(shape)))
        off += n
    return m.double() if dtype == "f64" else m


def ref_bcf_loss(model, s, a, x0, t, coef):
    """bcf.py:108-128 with the draws (x0, t) injected."""
    import torch
    import torch.nn.functional as F

    dt = next(model.parameters()).dtype
    states, actions, x0, t = (torch.from_numpy(np.asarray(v)).to(dt) for v in (s, a, x0, t.reshape(-1, 1)))
    x_t = (1 - t) * x0 + t * actions
    v_pred = model(x_t, states, t)
    flow_loss = F.mse_loss(v_pred, actions - x0)
    v_end = model(actions, states, torch.ones_like(t))
    action_loss = F.mse_loss(x0 + v_end, actions)
    return flow_loss + coef * action_loss, flow_loss, action_loss


def ref_sample(model, s, x0, n_steps):
    """flow_policy.py:49-63 with the initial draw injected."""
    import torch

    dt_ = next(model.parameters()).dtype
    state, x = torch.from_numpy(s).to(dt_), torch.from_numpy(x0).to(dt_)
    dt = 1.0 / n_steps
    with torch.no_grad():
        for i in range(n_steps):
            t = torch.full((x.shape[0], 1), fill_value=i * dt).to(dt_)
            x = x + model(x, state, t) * dt
    return x.numpy()


BCF_CASES = (  # (task, hidden, time_dim, B, coef, n_steps)
    ("maze", 256, 128, 128, 1.0, 32),
    ("hopper", 128, 128, 128, 1.0, 32),
    ("sine", 64, 32, 100, 0.5, 20),
    ("ant", 128, 64, 37, 2.0, 7),
    ("hand", 192, 48, 64, 1.0, 32),
)


def bcf_inputs(seed, B, s_dim, a_dim):
    s, _ = synth.states_actions(seed + 1, B, s_dim, a_dim)
    a = synth.expert_actions(seed + 2, s, a_dim)
    g = synth.rng(seed + 3)
    x0 = g.standard_normal((B, a_dim)).astype(np.float32)
    t = g.random(B).astype(np.float32)
    return s, a, x0, t


def gen_bcf():
    """bcf_*.npz: VectorNetwork.forward, the bcf loss + autograd gradients (fp32 as the reference runs, and an fp64
    copy of the same module for a tight pin of the oracle) and the Euler sampler; bcftrain_*.npz: 200 updates of
    _bc_step + _standard_step (clip 0.5, Adam eps 1e-12, lr 1e-3)."""
    import torch

    from oracle import bcf_closed_form as bcf

    VN = load_vector_network()
    for idx, (task, hidden, tdim, B, coef, n_steps) in enumerate(BCF_CASES):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = bcf.VNetSpec(s_dim, a_dim, hidden, tdim)
        seed = 7000 + idx
        flat = bcf.make_params(spec, seed, 1.0)
        s, a, x0, t = bcf_inputs(seed, B, s_dim, a_dim)
        out = {"meta": np.array([s_dim, a_dim, hidden, tdim, B, n_steps, seed]), "coef": np.float32(coef)}
        for tag in ("f32", "f64"):
            m = build_vnet(VN, spec, flat, tag)
            dt = next(m.parameters()).dtype
            with torch.no_grad():
                out[f"v_{tag}"] = m(torch.from_numpy(x0).to(dt), torch.from_numpy(s).to(dt),
                                    torch.from_numpy(t).to(dt)).numpy()     # 1-D t: model.py:423-424
            loss, fl, al = ref_bcf_loss(m, s, a, x0, t, coef)
            loss.backward()
            out[f"loss_{tag}"] = np.array([loss.item(), fl.item(), al.item()])
            g = np.concatenate([p.grad.numpy().ravel() for p in m.parameters()])
            out[f"grad_{tag}"] = g if g.size < 100_000 else g.astype(np.float32)   # keep the fixtures small
            out[f"sample_{tag}"] = ref_sample(m, s, x0, n_steps)
        name = f"bcf_{idx:02d}_{task}_h{hidden}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, out["loss_f32"])

    for idx, (task, hidden, steps) in enumerate((("hopper", 128, 200), ("maze", 256, 100))):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = bcf.VNetSpec(s_dim, a_dim, hidden, 128)
        seed, B, lr = 7500 + idx, 128, 1e-3
        m = build_vnet(VN, spec, bcf.make_params(spec, seed, 1.0))
        opt = torch.optim.Adam(m.parameters(), lr=lr, eps=1e-12)
        pool_s, _ = synth.states_actions(seed + 1, 4096, s_dim, a_dim)
        pool_a = synth.expert_actions(seed + 2, pool_s, a_dim)
        curve = np.zeros((steps, 3), np.float32)
        norms = np.zeros(steps, np.float32)
        for it in range(steps):
            g = synth.rng(seed + 100 + it)
            i = g.integers(0, 4096, B)
            x0 = g.standard_normal((B, a_dim)).astype(np.float32)
            t = g.random(B).astype(np.float32)
            loss, fl, al = ref_bcf_loss(m, pool_s[i], pool_a[i], x0, t, 1.0)
            opt.zero_grad()
            loss.backward()
            norms[it] = float(torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5))
            opt.step()
            curve[it] = (loss.item(), fl.item(), al.item())
        name = f"bcftrain_{idx:02d}_{task}_h{hidden}_{steps}steps.npz"
        np.savez_compressed(os.path.join(OUT, name), meta=np.array([s_dim, a_dim, hidden, 128, steps, B, seed]),
                            lr=np.float32(lr), curve=curve, gradnorm=norms,
                            final_params=np.concatenate([p.detach().numpy().ravel() for p in m.parameters()]))
        print("wrote", name, curve[0], curve[-1])


GAE_CASES = (  # (T, N, D, gamma, lambda, use_gae, proper_time_limits)
    (128, 32, 1, 0.99, 0.95, True, False),      # the reference's PPO defaults (num_steps 128, 32 processes)
    (128, 32, 1, 0.99, 0.95, True, True),
    (64, 7, 1, 0.97, 0.9, False, False),
    (64, 7, 2, 0.97, 0.9, False, True),
    (5, 3, 3, 0.9, 1.0, True, True),
)


def gae_inputs(seed, T, N, D):
    g = synth.rng(seed)
    rewards = g.standard_normal((T, N, 1)).astype(np.float32)
    value_preds = g.standard_normal((T + 1, N, D)).astype(np.float32)
    masks = (g.random((T + 1, N, 1)) > 0.05).astype(np.float32)       # ~5% episode ends
    bad_masks = (g.random((T + 1, N, 1)) > 0.03).astype(np.float32)   # ~3% time-limit truncations
    next_value = g.standard_normal((N, D)).astype(np.float32)
    return rewards, value_preds, masks, bad_masks, next_value


def gen_gae():
    """gae_*.npz: RolloutStorage.compute_returns / compute_advantages [rollout_storage.py:178-239] run as they are
    (the two methods are compiled from the reference file and bound to a stand-in object carrying the tensors; the
    storage class itself needs gym and the whole toolkit)."""
    import ast
    import types

    import torch

    path = os.path.join(ref_shim.REFERENCE_ROOT, "deps", "rl-toolkit", "rlf", "storage", "rollout_storage.py")
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RolloutStorage"][0]
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("compute_returns", "compute_advantages")]
    assert len(keep) == 2
    ns = {"torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    for idx, (T, N, D, gamma, lam, use_gae, proper) in enumerate(GAE_CASES):
        seed = 8000 + idx
        rewards, value_preds, masks, bad_masks, next_value = gae_inputs(seed, T, N, D)
        st = types.SimpleNamespace(
            rewards=torch.from_numpy(rewards), value_preds=torch.from_numpy(value_preds.copy()),
            masks=torch.from_numpy(masks), bad_masks=torch.from_numpy(bad_masks), value_dim=D,
            returns=torch.zeros(T + 1, N, D),
            args=types.SimpleNamespace(gamma=gamma, gae_lambda=lam, use_gae=use_gae, use_proper_time_limits=proper))
        ns["compute_returns"](st, torch.from_numpy(next_value))
        adv = ns["compute_advantages"](st)
        name = f"gae_{idx:02d}_T{T}_N{N}_D{D}_{'gae' if use_gae else 'disc'}{'_ptl' if proper else ''}.npz"
        np.savez_compressed(os.path.join(OUT, name), meta=np.array([T, N, D, int(use_gae), int(proper), seed]),
                            gamma=np.float64(gamma), gae_lambda=np.float64(lam), returns=st.returns.numpy(),
                            value_preds_after=st.value_preds.numpy(), advantages=adv.numpy())
        print("wrote", name, float(st.returns.abs().mean()))


RETNORM_CASES = [  # (T, N, gamma, rollouts, reward scale)
    (16, 8, 0.99, 1, 1.0),
    (128, 32, 0.99, 3, 5.0),     # the production shape (--num-steps 128 x --num-processes 32), state carried over rollouts
    (7, 1, 0.9, 2, 0.01),
]


def retnorm_inputs(seed, T, N, rollouts, scale):
    g = synth.rng(seed)
    rewards = (g.standard_normal((rollouts, T, N, 1)) * scale - 0.3 * scale).astype(np.float32)
    masks = (g.random((rollouts, T, N, 1)) > 0.04).astype(np.float32)
    return rewards, masks


def gen_retnorm():
    """retnorm_*.npz: FlowMatchingEstimation._get_reward [modril/flowril/flowril.py:210-230] run as it is -- the method is
    compiled from the reference file and bound to a stand-in object whose ``_compute_flow_reward`` returns the injected
    reward -- with the reference's own ``RunningMeanStd`` [rlf/baselines/common/running_mean_std.py:3-35] imported by path
    (numpy only).  Pins row a10: the discounted-return recursion, the float32 running moments, the scaling."""
    import ast
    import importlib.util
    import types

    import torch

    rms_path = os.path.join(ref_shim.REFERENCE_ROOT, "deps", "rl-toolkit", "rlf", "baselines", "common", "running_mean_std.py")
    spec_ = importlib.util.spec_from_file_location("ref_running_mean_std", rms_path)
    rms_mod = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(rms_mod)
    path = os.path.join(ref_shim.REFERENCE_ROOT, "modril", "flowril", "flowril.py")
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FlowMatchingEstimation"][0]
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_get_reward"]
    assert len(keep) == 1
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    net = types.SimpleNamespace(eval=lambda: None)
    for idx, (T, N, gamma, rollouts, scale) in enumerate(RETNORM_CASES):
        seed = 8800 + idx
        rewards, masks = retnorm_inputs(seed, T, N, rollouts, scale)
        obj = types.SimpleNamespace(
            args=types.SimpleNamespace(option="scrf", flowril_reward_norm=True, gamma=gamma), flow_net=net, returns=None,
            ret_rms=rms_mod.RunningMeanStd(shape=()))
        out = np.empty_like(rewards)
        state = []
        for ro in range(rollouts):
            storage = types.SimpleNamespace(masks=torch.from_numpy(masks[ro]))
            for step in range(T):
                obj._compute_flow_reward = lambda storage, step, add_info, ro=ro: torch.from_numpy(rewards[ro, step].copy())
                r, _ = ns["_get_reward"](obj, step, storage, {})
                out[ro, step] = r.numpy()
            state.append((np.float32(obj.ret_rms.mean), np.float32(obj.ret_rms.var), np.float64(obj.ret_rms.count)))
        name = f"retnorm_{idx:02d}_T{T}_N{N}_R{rollouts}.npz"
        np.savez_compressed(os.path.join(OUT, name), meta=np.array([T, N, rollouts, seed]), gamma=np.float64(gamma),
                            scale=np.float64(scale), normalised=out, final_returns=obj.returns.numpy(),
                            rms_mean=np.array([s[0] for s in state]), rms_var=np.array([s[1] for s in state]),
                            rms_count=np.array([s[2] for s in state]))
        print("wrote", name, float(np.abs(out).mean()))


def load_policy_classes():
    """The reference's ``MLPBasic`` (rlf/rl/model.py:246-296) and ``DiagGaussian`` (rlf/rl/distributions.py:26-75) class
    objects plus ``def_mlp_weight_init``, compiled from the reference sources as they lie under /root/reference (the
    ``rlf`` package itself imports gym).  NOTE: distributions.py patches ``torch.distributions.Normal`` in place
    (log_probs / entropy summed over the last axis / mode); this generator runs last for that reason."""
    import ast

    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    base = os.path.join(ref_shim.REFERENCE_ROOT, "deps", "rl-toolkit", "rlf", "rl")
    ns = {"torch": torch, "nn": nn, "np": np, "F": F, "math": __import__("math")}
    path = os.path.join(base, "model.py")
    tree = ast.parse(open(path).read())
    names = ("weight_init", "def_mlp_weight_init", "BaseNet", "MLPBase", "MLPBasic")
    keep = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    assert len(keep) == len(names), [n.name for n in keep]
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    path = os.path.join(base, "distributions.py")
    tree = ast.parse(open(path).read())
    keep = []
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name == "DiagGaussian":
            keep.append(n)
        elif isinstance(n, ast.Assign) and "Normal" in ast.unparse(n) or (isinstance(n, ast.Assign) and "normal" in ast.unparse(n)):
            keep.append(n)
    assert sum(isinstance(n, ast.ClassDef) for n in keep) == 1 and len(keep) == 7, [ast.unparse(n)[:40] for n in keep]
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["MLPBasic"], ns["DiagGaussian"], ns["def_mlp_weight_init"]


def build_policy(classes, spec, flat, dtype=None):
    """The module tree DistActorCritic.init builds for get_ppo_policy (deps/baselines/get_policy.py:16-23): same
    sub-module names and registration order (critic, critic_head: base_actor_critic.py:42-44; actor, dist:
    dist_actor_critic.py:43-47), so state_dict order is the reference's."""
    import torch
    import torch.nn as nn

    MLPBasic, DiagGaussian, def_init = classes

    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.critic = MLPBasic(spec.obs_dim, hidden_size=spec.hidden, num_layers=spec.layers)
            self.critic_head = def_init(nn.Linear(spec.hidden, 1))
            self.actor = MLPBasic(spec.obs_dim, hidden_size=spec.hidden, num_layers=spec.layers)
            self.dist = DiagGaussian(spec.hidden, spec.a_dim)

        def forward(self, state):                                   # dist_actor_critic.py:63-74, identity base net
            value = self.critic_head(self.critic(state, None, None)[0])
            return self.dist(self.actor(state, None, None)[0]), value

    m = Holder()
    sd = m.state_dict()
    assert [k for k in sd] == [n for n, _ in spec.shapes()], list(sd)
    off = 0
    for name, shape in spec.shapes():
        n = int(np.prod(shape))
        sd[name].copy_(torch.from_numpy(np.asarray(flat[off:off + n], np.float32).reshape(shape)))
        off += n
    return m.double() if dtype == "f64" else m


def ref_ppo_loss(model, obs, action, old_logp, adv, ret, old_value, clip, vcoef, ecoef):
    """evaluate_actions (dist_actor_critic.py:76-86) + the loss statements of PPO.update (ppo.py:29-49)."""
    import torch

    dt = next(model.parameters()).dtype
    obs, action, old_logp, adv, ret, old_value = (torch.from_numpy(np.asarray(v)).to(dt)
                                                  for v in (obs, action, old_logp, adv, ret, old_value))
    dist, value = model(obs)
    log_prob, ent = dist.log_probs(action), dist.entropy()
    ratio = torch.exp(log_prob - old_logp)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * adv
    actor_loss = -torch.min(surr1, surr2).mean(0)
    value_pred_clipped = old_value + (value - old_value).clamp(-clip, clip)
    value_losses = (value - ret).pow(2)
    value_losses_clipped = (value_pred_clipped - ret).pow(2)
    value_loss = 0.5 * torch.max(value_losses, value_losses_clipped).mean()
    loss = (value_loss * vcoef + actor_loss - ent.mean() * ecoef)
    return loss.sum(), value_loss, actor_loss.sum(), ent.mean(), (value, log_prob, ent, dist.mean)


PPO_CASES = (  # (task, hidden, layers, B, clip, value_coef, entropy_coef)
    ("hopper", 64, 2, 1024, 0.2, 0.5, 0.01),       # the reference's defaults: 128 steps x 32 processes / 4 minibatches
    ("maze", 256, 3, 128, 0.2, 0.5, 0.01),
    ("ant", 64, 2, 200, 0.1, 1.0, 0.0),
    ("sine", 16, 1, 37, 0.3, 0.5, 0.05),
    ("hand", 128, 2, 64, 0.2, 0.25, 0.001),
)


def ppo_inputs(spec, flat, seed, B):
    """A synthetic minibatch whose 'old' quantities come from a slightly different policy, so that a fair share of the
    rows hits the ratio clip and the value clip."""
    from oracle import ppo_closed_form as pcf

    g = synth.rng(seed)
    obs, _ = synth.states_actions(seed + 1, B, spec.obs_dim, spec.a_dim)
    old = (flat.astype(np.float64) + 0.02 * g.standard_normal(flat.size)).astype(np.float32)
    noise = g.standard_normal((B, spec.a_dim)).astype(np.float32)
    old_value, action, old_logp = (x.astype(np.float32) for x in pcf.ac_act(spec, old, obs, noise))
    adv = g.standard_normal((B, 1)).astype(np.float32)
    ret = (old_value + 0.5 * g.standard_normal((B, 1))).astype(np.float32)
    return obs, action, old_logp, adv, ret, old_value, noise


def gen_ppo():
    """ppo_*.npz: value / action mean / log-prob / entropy of the reference's policy classes, the clipped-PPO loss and
    its autograd gradients (fp32 as the reference runs + an fp64 copy of the same modules); ppotrain_*.npz: minibatch
    updates with _standard_step (clip_grad_norm_ 0.5, Adam lr 1e-3 eps 1e-12: base_net_algo.py:84-146)."""
    import torch

    from oracle import ppo_closed_form as pcf

    classes = load_policy_classes()
    for idx, (task, hidden, layers, B, clip, vcoef, ecoef) in enumerate(PPO_CASES):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = pcf.ACSpec(s_dim, a_dim, hidden, layers)
        seed = 9000 + idx
        flat = pcf.ac_make_params(spec, seed)
        obs, action, old_logp, adv, ret, old_value, noise = ppo_inputs(spec, flat, seed, B)
        out = {"meta": np.array([s_dim, a_dim, hidden, layers, B, seed]), "coefs": np.array([clip, vcoef, ecoef])}
        for tag in ("f32", "f64"):
            m = build_policy(classes, spec, flat, tag)
            loss, vl, al, ent, (value, logp, ents, mean) = ref_ppo_loss(m, obs, action, old_logp, adv, ret, old_value,
                                                                        clip, vcoef, ecoef)
            loss.backward()
            out[f"loss_{tag}"] = np.array([loss.item(), vl.item(), al.item(), ent.item()])
            g = np.concatenate([p.grad.numpy().ravel() for p in m.parameters()])
            out[f"grad_{tag}"] = g if g.size < 100_000 else g.astype(np.float32)   # keep the fixtures small
            out[f"value_{tag}"], out[f"logp_{tag}"] = value.detach().numpy(), logp.detach().numpy()
            out[f"ent_{tag}"], out[f"mean_{tag}"] = ents.detach().numpy(), mean.detach().numpy()
            if tag == "f32":
                ratio = np.exp(out["logp_f32"] - old_logp)
                out["clip_frac"] = np.array([np.mean(np.abs(ratio - 1) > clip), np.mean(np.abs(out["value_f32"] - old_value) > clip)])
        name = f"ppo_{idx:02d}_{task}_h{hidden}x{layers}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, out["loss_f32"], "clipped (ratio, value):", out["clip_frac"])

    for idx, (task, hidden, layers, rows, nmb, epochs) in enumerate((("hopper", 64, 2, 1024, 4, 10), ("maze", 128, 3, 512, 4, 8))):
        s_dim, a_dim = synth.TASK_DIMS[task]
        spec = pcf.ACSpec(s_dim, a_dim, hidden, layers)
        seed, lr = 9500 + idx, 1e-3
        flat = pcf.ac_make_params(spec, seed)
        obs, action, old_logp, adv, ret, old_value, _ = ppo_inputs(spec, flat, seed, rows)
        m = build_policy(classes, spec, flat)
        opt = torch.optim.Adam(m.parameters(), lr=lr, eps=1e-12)
        steps = epochs * nmb
        curve, norms = np.zeros((steps, 4), np.float32), np.zeros(steps, np.float32)
        mb = rows // nmb
        for e in range(epochs):
            perm = synth.rng(seed + 100 + e).permutation(rows)       # stands in for SubsetRandomSampler (:270-272)
            for k in range(nmb):
                i = perm[k * mb:(k + 1) * mb]
                loss, vl, al, ent, _ = ref_ppo_loss(m, obs[i], action[i], old_logp[i], adv[i], ret[i], old_value[i],
                                                    0.2, 0.5, 0.01)
                opt.zero_grad()
                loss.backward()
                norms[e * nmb + k] = float(torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5))
                opt.step()
                curve[e * nmb + k] = (loss.item(), vl.item(), al.item(), ent.item())
        name = f"ppotrain_{idx:02d}_{task}_h{hidden}x{layers}_{steps}steps.npz"
        np.savez_compressed(os.path.join(OUT, name), meta=np.array([s_dim, a_dim, hidden, layers, rows, nmb, epochs, seed]),
                            lr=np.float32(lr), curve=curve, gradnorm=norms,
                            final_params=np.concatenate([p.detach().numpy().ravel() for p in m.parameters()]))
        print("wrote", name, curve[0], curve[-1])


PRETRAIN_CASES = (  # (task, rows, hidden, epochs, lr, mode)   mode: autotune (warm-up weights) / fixed weights / 2fs
    ("hopper", 1001, 128, 3, 1e-3, "autotune"),
    ("maze", 600, 128, 2, 1e-3, "fixed"),
    ("sine", 500, 128, 2, 1e-3, "2fs"),
)


def pretrain_dataset(seed, task, rows):
    """Un-normalised synthetic demonstrations (observations with per-column offsets / scales so that norm_vec matters)."""
    s_dim, a_dim = synth.TASK_DIMS[task]
    g = synth.rng(seed)
    s, _ = synth.states_actions(seed + 1, rows, s_dim, a_dim)
    a = synth.expert_actions(seed + 2, s, a_dim)
    scale = g.uniform(0.5, 4.0, s_dim).astype(np.float32)
    shift = g.uniform(-2.0, 2.0, s_dim).astype(np.float32)
    return (s * scale + shift).astype(np.float32), (a * np.float32(1.7) + np.float32(0.3)).astype(np.float32)


def pretrain_draws(seed, epoch, k, B, a_dim, eps):
    """The draws of batch k of an epoch in the order the reference makes them (pretrain.py:368-378)."""
    g = synth.rng(seed + 1000 * (epoch + 1) + k)
    d = {}
    for tag in ("E", "A"):
        d["u" + tag] = g.random((B, 1), dtype=np.float32)
        d["t" + tag] = (d["u" + tag] * np.float32(1 - 2 * eps) + np.float32(eps)).astype(np.float32)
        d["n" + tag] = g.standard_normal((B, a_dim)).astype(np.float32)
        if tag == "E":
            d["z"] = g.standard_normal((B, a_dim)).astype(np.float32)       # randn_like(a) between the two fm_loss calls
    d["t_mix"] = g.random((B, 1), dtype=np.float32)
    d["probe"] = synth.rademacher(seed + 7, (B, a_dim))                     # shape-deterministic in the reference
    return d


def gen_pretrain(P):
    """pretrain_*.npz: the loop of the reference's offline pre-trainer (pretrain.py:269-440) on a small synthetic
    demonstration set: norm_vec with the data's own moments, the odd-row rule, shuffled 128-row batches (ragged last
    batch), expert + perturbed-agent CFM terms, regularisers with warm-up / fixed weights, clip 1, Adam."""
    import torch

    eps, B = 1e-2, 128
    for idx, (task, rows, hidden, epochs, lr, mode) in enumerate(PRETRAIN_CASES):
        s_dim, a_dim = synth.TASK_DIMS[task]
        seed = 11000 + 100 * idx
        obs_raw, act_raw = pretrain_dataset(seed, task, rows)
        obs, actions = torch.from_numpy(obs_raw), torch.from_numpy(act_raw)
        do_norm = task != "sine"                                              # pretrain.py:258-259
        moments = {}
        if do_norm:                                                           # :273-285, the reference's own norm_vec
            moments = {"obs_mean": obs.mean(0), "obs_std": obs.std(0), "act_mean": actions.mean(0), "act_std": actions.std(0)}
            obs = P.norm_vec(obs, moments["obs_mean"], moments["obs_std"])
            actions = P.norm_vec(actions, moments["act_mean"], moments["act_std"])
        dataset = torch.cat((obs, actions), 1)
        if dataset.size(0) % 2 == 1:                                          # :289-290
            dataset = dataset[1:]
        dataset = dataset.float()
        n = dataset.size(0)
        n_heads = 1 if mode == "2fs" else 2
        spec = cf.NetSpec(s_dim, a_dim, hidden, 4, n_heads)
        flat0 = cf.make_params(spec, seed + 5, 1.0)
        model = build_ref(P, spec, flat0, eps=eps)
        params = list(model.parameters())
        opt = torch.optim.Adam(params, lr=lr)
        w_anti, w_st = 1e-4, 1e-6                                             # --loss-anti / --loss-stable defaults
        curve, weights, norms = [], [], []
        for e in range(epochs):
            perm = synth.rng(seed + 50 + e).permutation(n)
            for k, lo in enumerate(range(0, n, B)):
                i = perm[lo:lo + B]
                batch = dataset[torch.from_numpy(i)]
                s, a = batch[:, :s_dim].numpy(), batch[:, s_dim:].numpy()
                d = pretrain_draws(seed, e, k, len(i), a_dim, eps)
                if mode == "2fs":
                    loss = ref_fm_loss(P, model, spec, s, a, None, d["uE"], d["nE"], dequant=d["z"])   # :414-415
                    l_e, l_anti, l_st = loss, torch.zeros(()), torch.zeros(())
                else:
                    l_e = ref_fm_loss(P, model, spec, s, a, "expert", d["uE"], d["nE"])
                    a_noise = (torch.from_numpy(a) + 0.1 * torch.from_numpy(d["z"])).numpy()            # :370-371
                    l_a = ref_fm_loss(P, model, spec, s, a_noise, "agent", d["uA"], d["nA"])
                    loss = l_e + l_a
                    l_anti, l_st = ref_reg_losses(P, model, s, a_noise, d["t_mix"], d["probe"])
                    if mode == "autotune":                                    # warm-up branch (:391-402)
                        w_anti = l_e.detach() / (l_anti.detach() + 1e-8) * 1e-3
                        w_st = l_e.detach() / (l_st.detach() + 1e-8) * 1e-3
                    loss = loss + w_anti * l_anti + w_st * l_st
                opt.zero_grad()
                loss.backward()
                norms.append(float(torch.nn.utils.clip_grad_norm_(params, 1)))                          # :417-420
                opt.step()
                curve.append((loss.item(), l_e.item(), l_anti.item(), l_st.item()))
                weights.append((float(w_anti), float(w_st)))
        out = {"meta": np.array([s_dim, a_dim, hidden, rows, epochs, seed, n_heads, int(do_norm)]), "lr": np.float32(lr),
               "mode": np.array(["autotune", "fixed", "2fs"].index(mode)), "curve": np.array(curve, np.float32),
               "weights": np.array(weights, np.float32), "gradnorm": np.array(norms, np.float32),
               "dataset": dataset.numpy(), "final_params": np.concatenate([p.detach().numpy().ravel() for p in params])}
        out.update({k: v.numpy() for k, v in moments.items()})
        name = f"pretrain_{idx:02d}_{task}_{mode}_h{hidden}_{epochs}epochs.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print("wrote", name, len(curve), "steps", curve[0], curve[-1])


def gen_yaml_keys():
    """flowril_yaml_keys.json: the sweep parameters of the reference's flowril configs (configs/sine/flowril*.yaml: the only
    flowril YAMLs the reference ships), so that the flag-compatibility test runs where /root/reference is absent."""
    import glob
    import json

    import yaml

    out = {}
    for f in sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "configs", "sine", "flowril*.yaml"))):
        y = yaml.safe_load(open(f))
        out[os.path.basename(f)] = {k: v for k, v in y.get("parameters", {}).items()}
    with open(os.path.join(OUT, "flowril_yaml_keys.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote flowril_yaml_keys.json", {k: len(v) for k, v in out.items()})


def main():
    import torch

    assert ref_shim.available(), "reference checkout not found"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    P = ref_shim.load()
    which = sys.argv[1:] or ["logprob", "vfield", "fmloss", "train", "reg", "trainreg", "gradnorm", "pretrain", "bcf", "gae", "yaml", "ppo", "retnorm"]
    if "logprob" in which:
        gen_logprob(P)
    if "logprob_smooth" in which or "logprob" in which:
        gen_logprob_smooth(P)
    if "vfield" in which:
        gen_vfield(P)
    if "fmloss" in which:
        gen_fmloss(P)
    if "train" in which:
        gen_train(P)
    for w in which:
        if w.startswith("train:"):
            gen_train(P, only=[int(x) for x in w.split(":", 1)[1].split(",")])
    if "reg" in which:
        gen_reg(P)
    if "trainreg" in which:
        gen_train_reg(P)
    if "gradnorm" in which:
        gen_gradnorm()
    if "pretrain" in which:
        gen_pretrain(P)
    if "bcf" in which:
        gen_bcf()
    if "gae" in which:
        gen_gae()
    if "yaml" in which:
        gen_yaml_keys()
    if "ppo" in which:
        gen_ppo()
    if "retnorm" in which:
        gen_retnorm()


if __name__ == "__main__":
    main()
