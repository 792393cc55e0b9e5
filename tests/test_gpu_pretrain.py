"""GPU parity of the offline pre-trainer (modril_b200.flowril.pretrain.OfflinePretrainer: frl_column_moments / frl_norm_vec
+ the fused CFM / regulariser / clip+Adam kernels through ``train_step``) against the reference's own pre-training loop run
on the same demonstrations and draws (tests/golden/pretrain_*.npz, oracle/gen_golden.py::gen_pretrain)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import synth
from oracle.gen_golden import pretrain_dataset, pretrain_draws
from tests.gpu_common import dev

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "pretrain_*.npz")))


class ReplayDraws:
    """Feeds the recorded draws of one epoch to the trainer in the order it asks for them."""

    def __init__(self, seed, epoch, n, a_dim, eps, two_fs):
        self.seed, self.epoch, self.n, self.a_dim, self.eps, self.two_fs = seed, epoch, n, a_dim, eps, two_fs
        self.k, self.cur, self.calls = -1, None, 0

    def perm(self, n):
        assert n == self.n
        return torch.from_numpy(synth.rng(self.seed + 50 + self.epoch).permutation(n)).cuda()

    def _batch(self, B, first):
        if first:
            self.k += 1
            self.cur = pretrain_draws(self.seed, self.epoch, self.k, B, self.a_dim, self.eps)
            self.calls = 0
        return self.cur

    def cfm(self, B):
        d = self._batch(B, first=not self.two_fs and self.calls == 0)
        tag = "E" if self.calls == 0 or self.two_fs else "A"
        self.calls += 1
        return dev(d["t" + tag]), dev(d["n" + tag])

    def normal(self, B):
        return dev(self._batch(B, first=self.two_fs)["z"])   # 2fs draws the dequantisation first

    def t_mix(self, B):
        self.calls = 0
        return dev(self.cur["t_mix"])


def _rel_l2(got, ref):
    ref = np.asarray(ref, np.float64)
    return np.sqrt(((np.asarray(got, np.float64) - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum())


def test_fixtures_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize("path", CASES, ids=os.path.basename)
def test_pretrain_loop_follows_reference(path, monkeypatch):
    from modril_b200.flowril import pretrain as pt

    g = np.load(path)
    s_dim, a_dim, hidden, rows, epochs, seed, n_heads, do_norm = (int(x) for x in g["meta"])
    mode = ["autotune", "fixed", "2fs"][int(g["mode"])]
    task = os.path.basename(path).split("_")[2]
    obs_raw, act_raw = pretrain_dataset(seed, task, rows)
    tr = pt.OfflinePretrainer(obs_raw, act_raw, option="2fs" if mode == "2fs" else "scrf", hidden_dim=hidden, depth=4, eps=1e-2,
                              lr=float(g["lr"]), norm=bool(do_norm), autotune=mode == "autotune")
    # data preparation: the reference's norm_vec with the data's own moments, then the odd-row rule
    data = torch.cat([tr.obs, tr.actions], 1).cpu().numpy()
    assert data.shape == g["dataset"].shape
    assert np.abs(data - g["dataset"]).max() <= 2e-5
    if do_norm:
        assert np.allclose(tr.obs_moments[0].cpu().numpy(), g["obs_mean"], rtol=1e-5, atol=1e-6)
        assert np.allclose(tr.obs_moments[1].cpu().numpy(), g["obs_std"], rtol=1e-5)
        assert np.allclose(tr.action_moments[1].cpu().numpy(), g["act_std"], rtol=1e-5)
    with torch.no_grad():                                       # continue from the reference's (normalised) rows bit for bit
        tr.obs.copy_(dev(g["dataset"][:, :s_dim]))
        tr.actions.copy_(dev(g["dataset"][:, s_dim:]))
        spec = cf.NetSpec(s_dim, a_dim, hidden, 4, n_heads)
        tr.estimator.net.flat_params().copy_(torch.from_numpy(cf.make_params(spec, seed + 5, 1.0)))
        tr.estimator.net.mark_dirty()
    probes = {}
    monkeypatch.setattr(pt, "_rademacher", lambda shape, device=None: probes.setdefault(
        shape[0], dev(synth.rademacher(seed + 7, tuple(shape)))))
    n = tr.obs.shape[0]
    curve, weights = [], []
    for e in range(epochs):
        out = tr.epoch(ReplayDraws(seed, e, n, a_dim, 1e-2, mode == "2fs"))
        curve.append(out.cpu().numpy())
    curve = np.concatenate(curve)
    ref = g["curve"]
    assert curve.shape[0] == ref.shape[0] == tr.steps
    k = min(8, len(ref))
    assert np.allclose(curve[:k, 0], ref[:k, 0], rtol=2e-3), np.abs(curve[:k, 0] / ref[:k, 0] - 1).max()
    assert np.allclose(curve[:k, 1], ref[:k, 1], rtol=2e-3)
    assert np.allclose(curve[:, 1], ref[:, 1], rtol=3e-2), np.abs(curve[:, 1] / ref[:, 1] - 1).max()
    assert _rel_l2(tr.estimator.net.flat_params().cpu().numpy(), g["final_params"]) <= 1e-2
    if mode != "2fs":   # the weights the last step used (args.loss_anti / args.loss_stable of the reference)
        w = np.array([float(tr._w_anti[0]), float(tr._w_stable[0])])
        assert np.allclose(w, g["weights"][-1], rtol=5e-2), (w, g["weights"][-1])
    assert abs(tr.train_loss_list[0] - ref[:len(ref) // epochs, 0].mean()) <= 2e-2 * abs(ref[:len(ref) // epochs, 0].mean())


def test_norm_vec_and_cli_roundtrip(tmp_path):
    """norm_vec clamps at +-10; the command line trains from a reference-format .pt file and writes the reference's files."""
    from modril_b200.flowril import pretrain as pt

    x = torch.tensor([[0.0, 100.0], [1.0, -100.0], [2.0, 0.0], [3.0, 0.0]], device="cuda")
    mean, std = pt.column_moments(x)
    assert torch.allclose(mean.cpu(), x.cpu().mean(0)) and torch.allclose(std.cpu(), x.cpu().std(0))
    y = pt.norm_vec(x, torch.zeros(2, device="cuda"), torch.full((2,), 1.0, device="cuda"))
    assert float(y.max()) == 10.0 and float(y.min()) == -10.0 and torch.equal(y[:, 0], x[:, 0])
    obs, act = pretrain_dataset(1, "hopper", 301)
    src = tmp_path / "hopper.pt"
    torch.save({"obs": torch.from_numpy(obs), "actions": torch.from_numpy(act)}, src)
    pt.main(["--traj-load-path", str(src), "--num-epoch", "2", "--hidden-dim", "128", "--save-path", str(tmp_path / "pre"),
             "--enable-loss-stable", "false"])
    d = tmp_path / "pre" / "hopper" / "scrf_no_s" / "trained_models"
    sd = torch.load(d / "hopper_fm.pt")
    assert list(sd)[:2] == ["net.shared.0.weight", "net.shared.0.bias"] and len(torch.load(d / "hopper_train_loss.pt")) == 2
    with pytest.raises(RuntimeError):
        pt.OfflinePretrainer(obs, act, device="cpu")


def test_pretrained_checkpoint_feeds_the_reward_module(tmp_path):
    """The file the pre-trainer saves is what ``--flow-path`` loads (reference flowril.py:48-58): the scrf reward module
    starts from it, freezes the shared trunk and head_c and fine-tunes head_r only."""
    from modril_b200.flowril import pretrain as pt
    from tests.test_gpu_module import FakeStorage, make_args, make_module

    obs, act = pretrain_dataset(3, "hopper", 400)
    tr = pt.OfflinePretrainer(obs, act, hidden_dim=128, lr=1e-3, norm=True)
    tr.fit(2)
    d = tr.save(str(tmp_path), "hopper", "scrf")
    args = make_args(flow_path=f"{d}/hopper_fm.pt", flowril_reward_norm=False)
    mod = make_module(args)
    net = mod.flow_net.net
    assert torch.equal(net.flat_params(), tr.estimator.net.flat_params())
    before = {k: v.clone() for k, v in net.state_dict().items()}
    storage = FakeStorage(args.num_steps, args.num_processes, 11, 3, 12, args.device)
    logs = mod.update(storage)
    after = net.state_dict()
    for k in before:
        same = torch.equal(before[k], after[k])
        assert same != k.startswith("head_r"), k
    assert np.isfinite(logs["flow_loss"]) and torch.isfinite(storage.rewards).all()
