"""GPU parity of the CFM loss/gradient kernels and the fused clip+Adam step (through the C ABI)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from tests.gpu_common import assert_close, build_model, dev, spec_of

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _grad_close(grad, ref, what):
    """Gradient parity: relative L2 error <= 1e-4 and max abs error <= 5e-4 * max|ref|.  (A sample on a ReLU kink
    flips one mask between fp32 and the float64 oracle and moves a few entries by ~1/B of its contribution; the
    L2 bar is the parity bar, the max bar bounds those entries.)"""
    ref = np.asarray(ref, np.float64)
    d = np.asarray(grad, np.float64) - ref
    l2 = np.sqrt((d ** 2).sum()) / np.sqrt((ref ** 2).sum())
    mx = np.abs(d).max() / np.abs(ref).max()
    assert l2 <= 1e-4 and mx <= 5e-4, f"{what}: rel L2 {l2:.3e}, max {mx:.3e}"


def _flat_grad(model):
    net = model.net
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                      for p, _, _ in net.param_slices()]).cpu().numpy()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "fmloss_*.npz"))), ids=os.path.basename)
def test_fmloss_and_autograd_grads_vs_reference(path):
    g = np.load(path)
    spec = spec_of(g["meta"])
    B, seed = int(g["meta"][6]), int(g["meta"][7])
    flat = cf.make_params(spec, seed, float(g["gain"]))
    s, a = synth.states_actions(seed, B, spec.s_dim, spec.a_dim)
    for ri, role in enumerate(("expert", "agent")):
        m = build_model(spec, flat)
        t, noise = synth.cfm_draws(seed + 1 + ri, B, spec.a_dim)
        if spec.n_heads == 2:
            loss = m.fm_loss(dev(s), dev(a), role, t=dev(t), noise=dev(noise))
        else:
            dq = synth.rng(seed + 5 + ri).standard_normal((B, spec.a_dim)).astype(np.float32)
            loss = m.fm_loss(dev(s), dev(a), dequant=dev(dq), t=dev(t), noise=dev(noise))
        (2.0 * loss).backward()  # exercises grad_output scaling
        grad = _flat_grad(m) / 2.0
        assert_close(np.array(loss.item()), g[f"loss_{role}"], 1e-5, f"loss {role}")
        gn = np.sqrt((grad.astype(np.float64) ** 2).sum())
        assert_close(np.array(gn), g[f"gradnorm_{role}"], 1e-4, f"grad norm {role}")
        norms, off = [], 0
        for _, shape in spec.shapes():
            n = int(np.prod(shape))
            norms.append(np.sqrt((grad[off:off + n].astype(np.float64) ** 2).sum()))
            off += n
        assert np.allclose(norms, g[f"tensor_gradnorms_{role}"], rtol=2e-4, atol=1e-6)
        if f"grad_{role}" in g:
            ref = g[f"grad_{role}"]
            assert np.abs(grad - ref).max() <= 1e-4 * np.abs(ref).max()
        # and against the float64 oracle for every case
        dq64 = None if spec.n_heads == 2 else dq * np.float32(0.02)
        ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, role, t, noise, dequant=dq64)
        _grad_close(grad, ref_grad, "vs float64 oracle")


@pytest.mark.parametrize("B", [1, 7, 64, 129, 1000, 5000])
def test_fmloss_ragged_vs_oracle(B):
    spec = cf.NetSpec(17, 6, 256, 4, 2)
    flat = cf.make_params(spec, 91, 2.0)
    m = build_model(spec, flat)
    s, a = synth.states_actions(300 + B, B, 17, 6)
    t, noise = synth.cfm_draws(400 + B, B, 6)
    loss = m.fm_loss(dev(s), dev(a), "agent", t=dev(t), noise=dev(noise))
    loss.backward()
    ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, "agent", t, noise)
    assert abs(loss.item() - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    grad = _flat_grad(m)
    _grad_close(grad, ref_grad, f"B={B}")


def test_fmloss_deterministic_and_no_grad():
    spec = cf.NetSpec(11, 3, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 5, 2.0))
    B = 4096
    s, a = synth.states_actions(1, B, 11, 3)
    t, noise = synth.cfm_draws(2, B, 3)
    grads = []
    for _ in range(2):
        m.zero_grad()
        m.fm_loss(dev(s), dev(a), "expert", t=dev(t), noise=dev(noise)).backward()
        grads.append(_flat_grad(m))
    assert np.array_equal(grads[0], grads[1])
    with torch.no_grad():
        l = m.fm_loss(dev(s), dev(a), "expert", t=dev(t), noise=dev(noise))
    assert not l.requires_grad


def test_clip_adam_vs_oracle_and_torch():
    """frl_clip_adam == clip_grad_norm_ + torch.optim.Adam on the same numbers (flowril.py:324-330)."""
    from modril_b200.flowril import FusedAdam
    spec = cf.NetSpec(11, 3, 128, 4, 2)
    flat0 = cf.make_params(spec, 8, 1.0)
    m = build_model(spec, flat0)
    opt = FusedAdam([m.net], lr=1e-3)
    ref_p = [torch.nn.Parameter(p.detach().clone()) for p, _, _ in m.net.param_slices()]
    ref_opt = torch.optim.Adam(ref_p, lr=1e-3)
    rng = synth.rng(3)
    p64, m64, v64 = flat0.astype(np.float64), np.zeros(flat0.size), np.zeros(flat0.size)
    for step in range(1, 6):
        g = (rng.standard_normal(flat0.size) * (0.05 if step % 2 else 5e-4)).astype(np.float32)
        gt = dev(g)
        norm = opt.step_flat([gt.clone()], max_norm=1.0)
        off = 0
        for p in ref_p:
            p.grad = gt[off:off + p.numel()].view(p.shape).clone()
            off += p.numel()
        ref_norm = torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        ref_opt.step()
        assert abs(norm.item() - ref_norm.item()) <= 1e-6 * ref_norm.item()
        coef, _ = cf.clip_coef(g.astype(np.float64), 1.0)
        p64, m64, v64 = cf.adam_step(p64, g.astype(np.float64) * coef, m64, v64, step, 1e-3)
    got = m.net.flat_params().cpu().numpy()
    ref = torch.cat([p.detach().reshape(-1) for p in ref_p]).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-7, np.abs(got - ref).max()
    assert np.abs(got - p64).max() <= 1e-6
    # optimizer state in torch.optim.Adam layout
    sd = opt.state_dict()
    rsd = ref_opt.state_dict()
    assert set(sd["state"].keys()) == set(rsd["state"].keys())
    for k in rsd["state"]:
        assert torch.allclose(sd["state"][k]["exp_avg"], rsd["state"][k]["exp_avg"], atol=1e-7)
        assert float(sd["state"][k]["step"]) == float(rsd["state"][k]["step"])


def _train_case(path):
    g = np.load(path)
    spec = spec_of(g["meta"])
    steps, B, seed = (int(x) for x in g["meta"][5:8])
    kind = os.path.basename(path).split("_")[3]
    if "scrf_ft" in path:
        kind = "scrf_ft"
    return g, spec, steps, B, seed, kind


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "train_*.npz"))), ids=os.path.basename)
def test_training_curves_vs_reference(path):
    """The update loop of flowril.py:311-335 (loss E + loss A, clip_grad_norm_ 1.0, Adam) on the all-native route.
    Stated tolerance: per-step losses within 2e-3*max(1,|ref|) over the first 50 steps (fp32 round-off has not
    yet been amplified), and the 50-step moving average within 3e-2 relative up to the last step (1k for
    train_00)."""
    from modril_b200.flowril import FusedAdam, train_step
    g, spec, steps, B, seed, kind = _train_case(path)
    lr = float(g["lr"])
    nets = [build_model(spec, cf.make_params(spec, seed, 1.0))]
    if kind == "2fs":
        nets.append(build_model(spec, cf.make_params(spec, seed + 10, 1.0)))
    if kind == "scrf_ft":
        for p in nets[0].net.shared.parameters():
            p.requires_grad = False
        for p in nets[0].net.head_c.parameters():
            p.requires_grad = False
    opt = FusedAdam([m.net for m in nets], lr=lr)
    pool_s, _ = synth.states_actions(seed + 1, 4096, spec.s_dim, spec.a_dim)
    pool_aE = synth.expert_actions(seed + 2, pool_s, spec.a_dim)
    pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, spec.a_dim)), -1, 1).astype(np.float32)
    d_s, d_aE, d_aA = dev(pool_s), dev(pool_aE), dev(pool_aA)
    curve = np.zeros((steps, 2), np.float32)
    loss_log = []
    for it in range(steps):
        r = synth.rng(seed + 100 + it)
        iE, iA = r.integers(0, 4096, B), r.integers(0, 4096, B)
        tE, nE = synth.cfm_draws(seed + 10000 + 2 * it, B, spec.a_dim)
        tA, nA = synth.cfm_draws(seed + 10001 + 2 * it, B, spec.a_dim)
        aE, aA = d_aE[dev(iE)], d_aA[dev(iA)]
        if kind == "2fs":
            dqE = synth.rng(seed + 50000 + 2 * it).standard_normal((B, spec.a_dim)).astype(np.float32)
            dqA = synth.rng(seed + 50001 + 2 * it).standard_normal((B, spec.a_dim)).astype(np.float32)
            aE = aE + dev(dqE) * 0.02
            aA = aA + dev(dqA) * 0.02
        terms = [dict(net=nets[0].net, sign=1.0, s=d_s[dev(iE)], a0=aE, t=dev(tE), noise=dev(nE), rate=1.0),
                 dict(net=nets[-1].net, sign=-1.0, s=d_s[dev(iA)], a0=aA, t=dev(tA), noise=dev(nA), rate=1.0)]
        loss_log.append(train_step(terms, opt, max_norm=1.0))
    curve = np.array([[float(a), float(b)] for a, b in loss_log], np.float32)
    ref = g["curve"]
    head = min(50, steps)
    assert_close(curve[:head], ref[:head], 2e-3, "first steps")
    k = min(50, steps)
    ma = lambda x: np.stack([np.convolve(x[:, j], np.ones(k) / k, mode="valid") for j in range(2)], 1)
    rel = np.abs(ma(curve) - ma(ref)) / np.abs(ma(ref))
    assert rel.max() <= 3e-2, rel.max()
    if steps <= 50:
        for mi, m in enumerate(nets):
            got = m.net.flat_params().cpu().numpy()
            refp = g[f"final_params_{mi}"]
            assert np.abs(got - refp).max() <= 2e-3 * np.abs(refp).max(), np.abs(got - refp).max()
    if kind == "scrf_ft":  # frozen trunk + head_c untouched (flowril.py:52-58)
        init = cf.make_params(spec, seed, 1.0)
        got = nets[0].net.flat_params().cpu().numpy()
        n_frozen = sum(int(np.prod(s)) for name, s in spec.shapes() if "head_r" not in name)
        assert np.array_equal(got[:n_frozen], init[:n_frozen])
        assert not np.array_equal(got[n_frozen:], init[n_frozen:])


def test_torch_optimizer_route_matches_native_route():
    """fm_loss(...).backward(); clip_grad_norm_; FusedAdam.step()  ==  train_step(...)."""
    from modril_b200.flowril import FusedAdam, train_step
    spec = cf.NetSpec(6, 2, 128, 4, 2)
    flat = cf.make_params(spec, 12, 1.0)
    B = 128
    s, a = synth.states_actions(5, B, 6, 2)
    t, noise = synth.cfm_draws(6, B, 2)
    t2, noise2 = synth.cfm_draws(7, B, 2)
    m1, m2 = build_model(spec, flat), build_model(spec, flat)
    o1, o2 = FusedAdam([m1.net], lr=1e-3), FusedAdam([m2.net], lr=1e-3)
    for _ in range(3):
        o1.zero_grad()
        loss = m1.fm_loss(dev(s), dev(a), "expert", t=dev(t), noise=dev(noise)) + \
            0.5 * m1.fm_loss(dev(s), dev(a), "agent", t=dev(t2), noise=dev(noise2))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m1.parameters(), 1.0)
        o1.step()
        train_step([dict(net=m2.net, sign=1.0, s=dev(s), a0=dev(a), t=dev(t), noise=dev(noise), rate=1.0),
                    dict(net=m2.net, sign=-1.0, s=dev(s), a0=dev(a), t=dev(t2), noise=dev(noise2), rate=0.5)],
                   o2, max_norm=1.0)
    p1, p2 = m1.net.flat_params().cpu().numpy(), m2.net.flat_params().cpu().numpy()
    assert np.abs(p1 - p2).max() <= 1e-6


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_train_step_phases_equal_the_whole_step(precision):
    """train_step(phase='grad') -> 'reduce' -> 'apply' (the split a data-parallel loop captures into two CUDA graphs
    around the NCCL call) gives bit-identical losses and weights to the one-call update."""
    from modril_b200.flowril import FusedAdam, train_step

    spec = cf.NetSpec(11, 3, 256, 4, 2)
    flat = cf.make_params(spec, 31, 1.0)
    B = 300
    s, a = synth.states_actions(71, B, 11, 3)
    draws = [synth.cfm_draws(72 + k, B, 3) for k in range(2)]
    models = [build_model(spec, flat) for _ in range(2)]
    opts = [FusedAdam([m.net], lr=1e-3) for m in models]

    def terms(m):
        return [dict(net=m.net, sign=sg, s=dev(s), a0=dev(a), t=dev(t), noise=dev(z), rate=r)
                for (t, z), sg, r in zip(draws, (1.0, -1.0), (1.0, 0.5))]

    for _ in range(3):
        whole = train_step(terms(models[0]), opts[0], 1.0, precision=precision)
        part = train_step(terms(models[1]), opts[1], 1.0, precision=precision, phase="grad")
        part = train_step(terms(models[1]), opts[1], 1.0, precision=precision, phase="reduce", losses=part)
        part = train_step(terms(models[1]), opts[1], 1.0, precision=precision, phase="apply", losses=part)
        assert all(torch.equal(x, y) for x, y in zip(whole, part))
        assert torch.equal(models[0].net.flat_params(), models[1].net.flat_params())
    with pytest.raises(ValueError):
        train_step(terms(models[0]), opts[0], 1.0, phase="bogus")
