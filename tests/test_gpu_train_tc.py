"""GPU parity of the tcgen05 tensor-core CFM loss/gradient path (frl_cfm_loss_grad_p, FRL_PREC_FP16) through the
C ABI.  fp16 operands bound the accuracy (scripts/tc_train_err.py, B200): the loss is within 1e-4 of the float64
oracle (bar 2e-3); the gradient's relative L2 error is 2e-3..7e-3 at B = 4096 and 3e-3..2e-2 at B = 200 -- operand
rounding (2^-11) flips ReLU masks of units whose pre-activation is within rounding of 0, and with few rows a single
flipped unit is a visible share of the mean.  Bars: 3e-2 overall / 6e-2 per tensor on the 200-row stress nets, 1e-2 on
the ragged batches, and the 1000-step loss curve of the reference within the stated tolerance of the fp16 mode."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from tests.gpu_common import assert_close, build_model, dev, spec_of

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_stage_timeouts(monkeypatch):
    """Every net a test of this file trains must finish without a time-out in the fused kernel's stage hand-offs."""
    from modril_b200.flowril import networks

    made = []
    orig = networks._FlatVNet._finish_init

    def tracking(self, *a, **k):
        orig(self, *a, **k)
        made.append(self)

    monkeypatch.setattr(networks._FlatVNet, "_finish_init", tracking)
    yield
    for net in made:
        assert net.train_timeouts() == 0


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _flat_grad(model):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                      for p, _, _ in model.net.param_slices()]).cpu().numpy()


def _rel_l2(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.sqrt(((got - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum())


def _tensor_rel_l2(spec, got, ref):
    out, off = {}, 0
    for name, shape in spec.shapes():
        n = int(np.prod(shape))
        out[name] = _rel_l2(got[off:off + n], ref[off:off + n])
        off += n
    return out


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "fmloss_*.npz"))), ids=os.path.basename)
def test_tc_fmloss_and_grads_vs_reference_and_oracle(path):
    g = np.load(path)
    spec = spec_of(g["meta"])
    if spec.n_heads * spec.a_dim > 64:
        pytest.skip("a_dim too large for the tensor-core training path")
    B, seed = int(g["meta"][6]), int(g["meta"][7])
    flat = cf.make_params(spec, seed, float(g["gain"]))
    s, a = synth.states_actions(seed, B, spec.s_dim, spec.a_dim)
    for ri, role in enumerate(("expert", "agent")):
        m = build_model(spec, flat)
        t, noise = synth.cfm_draws(seed + 1 + ri, B, spec.a_dim)
        if spec.n_heads == 2:
            loss = m.fm_loss(dev(s), dev(a), role, t=dev(t), noise=dev(noise), precision="fp16")
            dq64 = None
        else:
            dq = synth.rng(seed + 5 + ri).standard_normal((B, spec.a_dim)).astype(np.float32)
            loss = m.fm_loss(dev(s), dev(a), dequant=dev(dq), t=dev(t), noise=dev(noise), precision="fp16")
            dq64 = dq * np.float32(0.02)
        loss.backward()
        grad = _flat_grad(m)
        assert_close(np.array(loss.item()), g[f"loss_{role}"], 2e-3, f"loss {role}")
        ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, role, t, noise, dequant=dq64)
        assert _rel_l2(grad, ref_grad) <= 3e-2, _rel_l2(grad, ref_grad)
        worst = max(_tensor_rel_l2(spec, grad, ref_grad).values())
        assert worst <= 6e-2, _tensor_rel_l2(spec, grad, ref_grad)
        if f"grad_{role}" in g:
            assert _rel_l2(grad, g[f"grad_{role}"]) <= 3e-2


@pytest.mark.parametrize("B", [1, 7, 128, 129, 1000, 5000, 40000])
def test_tc_fmloss_ragged_vs_fp32_kernels(B):
    spec = cf.NetSpec(17, 6, 256, 4, 2)
    flat = cf.make_params(spec, 91, 2.0)
    s, a = synth.states_actions(300 + B, B, 17, 6)
    t, noise = synth.cfm_draws(400 + B, B, 6)
    out = {}
    for prec in ("fp32", "fp16"):
        m = build_model(spec, flat)
        loss = m.fm_loss(dev(s), dev(a), "agent", t=dev(t), noise=dev(noise), precision=prec)
        loss.backward()
        out[prec] = (loss.item(), _flat_grad(m))
    assert abs(out["fp16"][0] - out["fp32"][0]) <= 2e-3 * max(1.0, abs(out["fp32"][0]))
    assert _rel_l2(out["fp16"][1], out["fp32"][1]) <= 1e-2, _rel_l2(out["fp16"][1], out["fp32"][1])
    worst = max(_tensor_rel_l2(spec, out["fp16"][1], out["fp32"][1]).values())
    assert worst <= 6e-2, _tensor_rel_l2(spec, out["fp16"][1], out["fp32"][1])


@pytest.mark.parametrize("dims", [(1, 1, 128, 4, 2), (6, 2, 128, 3, 1), (11, 3, 256, 2, 2), (132, 8, 256, 4, 2),
                                  (68, 20, 256, 4, 2), (16, 4, 128, 1, 1)])
def test_tc_other_shapes_deterministic(dims):
    spec = cf.NetSpec(*dims)
    flat = cf.make_params(spec, 3, 2.0)
    B = 700
    s, a = synth.states_actions(8, B, spec.s_dim, spec.a_dim)
    t, noise = synth.cfm_draws(9, B, spec.a_dim)
    grads = []
    for rep in range(2):
        m = build_model(spec, flat)
        if spec.n_heads == 2:
            loss = m.fm_loss(dev(s), dev(a), "expert", t=dev(t), noise=dev(noise), precision="fp16")
        else:
            loss = m.fm_loss(dev(s), dev(a), dequant=torch.zeros(B, spec.a_dim, device="cuda"), t=dev(t), noise=dev(noise),
                             precision="fp16")
        loss.backward()
        grads.append(_flat_grad(m))
    assert np.array_equal(grads[0], grads[1])
    ref_loss, ref_grad = cf.fm_loss_and_grad(spec, flat, s, a, "expert", t, noise)
    assert abs(loss.item() - ref_loss) <= 2e-3 * max(1.0, abs(ref_loss))
    assert _rel_l2(grads[0], ref_grad) <= 3e-2, _rel_l2(grads[0], ref_grad)


def test_tc_grad_scale_accumulate_and_mean_divisor():
    """train_step semantics on the tensor-core path: rate-weighted accumulation of two terms into one flat gradient
    equals the FP32 kernels' result to fp16 accuracy; loss-only call leaves no gradient."""
    from modril_b200.flowril import FusedAdam, train_step
    spec = cf.NetSpec(11, 3, 256, 4, 2)
    flat = cf.make_params(spec, 12, 1.0)
    B = 3000
    s, a = synth.states_actions(5, B, 11, 3)
    t, noise = synth.cfm_draws(6, B, 3)
    t2, noise2 = synth.cfm_draws(7, B, 3)
    res = {}
    for prec in ("fp32", "fp16"):
        m = build_model(spec, flat)
        opt = FusedAdam([m.net], lr=1e-3)
        terms = [dict(net=m.net, sign=1.0, s=dev(s), a0=dev(a), t=dev(t), noise=dev(noise), rate=1.0),
                 dict(net=m.net, sign=-1.0, s=dev(s), a0=dev(a), t=dev(t2), noise=dev(noise2), rate=0.5)]
        losses = train_step(terms, opt, max_norm=1e9, precision=prec)
        res[prec] = ([float(l) for l in losses], m.net.flat_grad().cpu().numpy().copy())
    for l16, l32 in zip(res["fp16"][0], res["fp32"][0]):
        assert abs(l16 - l32) <= 2e-3 * max(1.0, abs(l32))
    assert _rel_l2(res["fp16"][1], res["fp32"][1]) <= 1e-2
    with torch.no_grad():
        m = build_model(spec, flat)
        l = m.fm_loss(dev(s), dev(a), "expert", t=dev(t), noise=dev(noise), precision="fp16")
    assert abs(l.item() - res["fp32"][0][0]) <= 2e-3 * max(1.0, abs(res["fp32"][0][0]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "train_0[04]_*1000steps.npz"))), ids=os.path.basename)
def test_tc_training_curve_vs_reference(path):
    """1000 regular update steps of flowril.py:311-335 with the tensor-core gradient kernels: the loss curve tracks
    the reference's (tests/golden/train_00: H = 128; train_04: H = 256, the module default and the benchmarked net) --
    per-step 1.5e-2 over the first 50 steps, 50-step moving average within 5 % to the end (stated tolerance of the fp16
    mode; measured on a B200 with scripts/tc_curve_dev.py: first 50 steps 4.0e-3 (H = 128) / 1.23e-2 (H = 256), moving
    average 2.7 % / 2.1 %; the FP32 kernels on the same goldens: 1.8e-7 and 0.7 % / 1.1 % -- beyond ~100 steps any
    round-off is amplified chaotically, the reference against itself at another thread count included)."""
    from modril_b200.flowril import FusedAdam, train_step
    g = np.load(path)
    spec = spec_of(g["meta"])
    steps, B, seed = (int(x) for x in g["meta"][5:8])
    m = build_model(spec, cf.make_params(spec, seed, 1.0))
    opt = FusedAdam([m.net], lr=float(g["lr"]))
    pool_s, _ = synth.states_actions(seed + 1, 4096, spec.s_dim, spec.a_dim)
    pool_aE = synth.expert_actions(seed + 2, pool_s, spec.a_dim)
    pool_aA = np.clip(synth.rng(seed + 3).standard_normal((4096, spec.a_dim)), -1, 1).astype(np.float32)
    d_s, d_aE, d_aA = dev(pool_s), dev(pool_aE), dev(pool_aA)
    log = []
    for it in range(steps):
        r = synth.rng(seed + 100 + it)
        iE, iA = dev(r.integers(0, 4096, B)), dev(r.integers(0, 4096, B))
        tE, nE = synth.cfm_draws(seed + 10000 + 2 * it, B, spec.a_dim)
        tA, nA = synth.cfm_draws(seed + 10001 + 2 * it, B, spec.a_dim)
        terms = [dict(net=m.net, sign=1.0, s=d_s[iE], a0=d_aE[iE], t=dev(tE), noise=dev(nE), rate=1.0),
                 dict(net=m.net, sign=-1.0, s=d_s[iA], a0=d_aA[iA], t=dev(tA), noise=dev(nA), rate=1.0)]
        log.append(torch.stack(train_step(terms, opt, max_norm=1.0, precision="fp16")))
    curve = torch.stack(log).cpu().numpy()
    ref = g["curve"]
    assert_close(curve[:50], ref[:50], 1.5e-2, "first steps", kink_cap=1.5e-2)
    k = 50
    ma = lambda x: np.stack([np.convolve(x[:, j], np.ones(k) / k, mode="valid") for j in range(2)], 1)
    rel = np.abs(ma(curve) - ma(ref)) / np.abs(ma(ref))
    assert rel.max() <= 5e-2, rel.max()


@pytest.mark.parametrize("B0,B1,rates", [(1000, 300, (1.0, 0.5)), (128, 128, (1.0, 1.0)), (4096, 5000, (0.0, 2.0))])
def test_stacked_pair_pass_equals_two_single_passes(B0, B1, rates):
    """frl_cfm_loss_grad_pair (both roles' batches stacked row-wise through the net in one pass) against two
    frl_cfm_loss_grad_p calls.  With equal per-row scales the two routes round identically and differ only in the split
    of the batch reductions (<= 1e-5 relative); with unequal scales (ragged batches, unequal rates, a zero rate) the second
    term's dV is rounded to fp16 after its relative scale instead of before, which is fp16 noise (bar 3e-3)."""
    import ctypes as C

    from modril_b200 import _native
    from modril_b200.flowril.networks import _current_stream, _ptr

    spec = cf.NetSpec(11, 3, 256, 4, 2)
    m = build_model(spec, cf.make_params(spec, 77, 1.0))
    net = m.net
    h = net.native()
    L = _native.lib()
    ins, single, losses1 = [], None, []
    g = torch.zeros_like(net.flat_params())
    for k, (B, sign) in enumerate(((B0, 1.0), (B1, -1.0))):
        s, a = synth.states_actions(500 + k, B, 11, 3)
        t, noise = synth.cfm_draws(600 + k, B, 3)
        d = tuple(dev(x) for x in (s, a, t.reshape(-1), noise))
        ins.append(d)
        loss = torch.empty(1, device="cuda:0")
        _native.check(L.frl_cfm_loss_grad_p(h, sign, _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _ptr(d[3]), B, 0, rates[k], k, _ptr(loss),
                                            _ptr(g), _native.PREC_FP16, _current_stream(g.device)), "single")
        losses1.append(loss)
    single = g.clone()
    arr = (_native.CfmTerm * 2)()
    losses2 = [torch.empty(1, device="cuda:0") for _ in range(2)]
    for k, (B, sign) in enumerate(((B0, 1.0), (B1, -1.0))):
        d = ins[k]
        arr[k].role_sign, arr[k].B, arr[k].mean_divisor, arr[k].grad_scale = sign, B, 0, rates[k]
        arr[k].s, arr[k].a0, arr[k].t, arr[k].noise = (x.data_ptr() for x in d)
        arr[k].loss = losses2[k].data_ptr()
    g2 = torch.full_like(g, 7.0)    # overwritten (accumulate = 0)
    _native.check(L.frl_cfm_loss_grad_pair(h, arr, 0, _ptr(g2), _native.PREC_FP16, _current_stream(g.device)), "pair")
    for l1, l2 in zip(losses1, losses2):
        assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1))
    tol = 1e-5 if (B0 == B1 and rates[0] == rates[1]) else 3e-3
    assert _rel_l2(g2.cpu().numpy(), single.cpu().numpy()) <= tol
    # accumulate = 1 adds on top; FP32 precision runs the two FFMA passes
    _native.check(L.frl_cfm_loss_grad_pair(h, arr, 1, _ptr(g2), _native.PREC_FP16, _current_stream(g.device)), "pair acc")
    assert _rel_l2(g2.cpu().numpy(), 2 * single.cpu().numpy()) <= tol
    g3 = torch.zeros_like(g)
    _native.check(L.frl_cfm_loss_grad_pair(h, arr, 0, _ptr(g3), _native.PREC_FP32, _current_stream(g.device)), "pair fp32")
    assert _rel_l2(g3.cpu().numpy(), single.cpu().numpy()) <= 3e-2      # fp32 vs fp16 operands


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_train_step_graph_equals_eager_steps_and_scoring_sees_the_new_weights(prec):
    """A whole update replayed as ONE CUDA graph (TrainStepGraph) == the same updates launched eagerly, bit for bit; building
    the graph leaves weights and optimizer state untouched; and scoring after a replay uses the UPDATED weights (the
    wrapper marks the nets dirty -- raw-pointer writes do not bump torch's version counters)."""
    from modril_b200.flowril import FusedAdam, TrainStepGraph, train_step

    spec = cf.NetSpec(11, 3, 256, 4, 2)
    flat = cf.make_params(spec, 5, 1.0)
    B = 700
    s, a = synth.states_actions(3, B, 11, 3)
    draws = [synth.cfm_draws(10 + i, B, 3) for i in range(2)]
    probe = dev(synth.rademacher(4, (B, 3)))

    def make():
        m = build_model(spec, flat)
        terms = [dict(net=m.net, sign=sg, s=dev(s), a0=dev(a), t=dev(t), noise=dev(z), rate=1.0)
                 for sg, (t, z) in zip((1.0, -1.0), draws)]
        return m, terms, FusedAdam([m.net], lr=1e-3, capturable=True)

    m_e, terms_e, opt_e = make()
    m_g, terms_g, opt_g = make()
    before = m_g.net.flat_params().clone()
    score0 = m_g.score(dev(s), dev(a), n_steps=4, probe=probe, precision="fp32")[0].clone()
    graph = TrainStepGraph(terms_g, opt_g, 1.0, precision=prec)
    assert torch.equal(m_g.net.flat_params(), before) and opt_g.step_count() == 0
    for it in range(4):
        l_e = train_step(terms_e, opt_e, 1.0, precision=prec)
        l_g = graph()
        assert all(torch.equal(x, y) for x, y in zip(l_e, l_g))
        assert torch.equal(m_e.net.flat_params(), m_g.net.flat_params()), it
        if it == 1:   # new batch contents go in through the static input tensors
            for te, tg in zip(terms_e, terms_g):
                for k in ("t", "noise"):
                    te[k].copy_(te[k].flip(0))
                    tg[k].copy_(tg[k].flip(0))
    assert opt_g.step_count() == 4
    score_g = m_g.score(dev(s), dev(a), n_steps=4, probe=probe, precision="fp32")[0]
    score_e = m_e.score(dev(s), dev(a), n_steps=4, probe=probe, precision="fp32")[0]
    assert torch.equal(score_g, score_e) and not torch.equal(score_g, score0)


def test_fused_dataflow_kernel_matches_the_per_layer_launches(monkeypatch):
    """The dataflow launches -- FRL_TRAIN_FUSED=1: one persistent kernel, every GEMM stage resident on its own SMs, tiles
    handed over through L2 rings; =2: a forward-chain and a dgrad-chain kernel + the stand-alone wgrad kernel -- compute the
    same stacked-pair gradient as the per-layer launches (=0): identical MMAs per tile, identical fixed-order reduction of
    per-tile partials; only the grouping of row tiles into wgrad partial sums may differ (fp32 round-off)."""
    from modril_b200.flowril import FusedAdam, train_step

    spec = cf.NetSpec(17, 6, 256, 4, 2)
    flat = cf.make_params(spec, 8, 1.0)
    grads, losses = {}, {}
    for fused in ("0", "1", "2"):
        monkeypatch.setenv("FRL_TRAIN_FUSED", fused)
        m = build_model(spec, flat)
        opt = FusedAdam([m.net], lr=0.0)   # the weights stay put: three launches on identical weights
        terms = []
        for i, (sg, B) in enumerate(((1.0, 5000), (-1.0, 3333))):
            s, a = synth.states_actions(20 + i, B, 17, 6)
            t, z = synth.cfm_draws(30 + i, B, 6)
            terms.append(dict(net=m.net, sign=sg, s=dev(s), a0=dev(a), t=dev(t), noise=dev(z), rate=1.0))
        for _ in range(3):   # ring slots and flags are reused across launches
            losses[fused] = [float(x) for x in train_step(terms, opt, 1.0, precision="fp16")]
        grads[fused] = m.net.flat_grad().double().cpu().numpy()
        assert m.net.train_timeouts() == 0
    for mode in ("1", "2"):
        assert losses["0"] == pytest.approx(losses[mode], rel=1e-6)
        assert _rel_l2(grads[mode], grads["0"]) <= 2e-6
