"""Import shim for the *reference* estimators (build container only).

TEST INFRASTRUCTURE ONLY -- used by ``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by
``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is absent, e.g. on the GPU box).

``modril/flowril/pretrain.py`` imports matplotlib (absent here) and importing the ``modril`` package runs
``modril/__init__.py:1`` which imports gym (absent).  Both are side-stepped by registering stub modules and
namespace packages before the import.  No reference source is copied.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FLOWRIL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modril", "flowril", "pretrain.py"))


def load():
    """Returns the reference module ``modril.flowril.pretrain`` (CoupledResidualFM, FlowMatching, _rademacher)."""
    if "modril.flowril.pretrain" in sys.modules:
        return sys.modules["modril.flowril.pretrain"]
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    for name, sub in (("modril", "modril"), ("modril.flowril", "modril/flowril"), ("modril.utils", "modril/utils")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, sub)]
            sys.modules[name] = m
    import importlib

    return importlib.import_module("modril.flowril.pretrain")


@contextlib.contextmanager
def injected_randomness(P, *, probes=None, rand=None, normal=None, randn_like=None):
    """Make the reference consume caller-supplied randomness.

    probes:     list of tensors returned in order by P._rademacher (scrf) and torch.randint_like (2fs; the
                list holds the +-1 probes, the patched randint_like returns (probe+1)/2 so that the reference's
                ``.float().mul_(2).sub_(1)`` at pretrain.py:90 reproduces it).
    rand:       list of tensors returned by torch.rand (t draws, pretrain.py:99, :182).
    normal:     list of tensors returned by torch.distributions.Normal.sample (pretrain.py:101, :183).
    randn_like: list returned by torch.randn_like (2fs dequantisation, pretrain.py:98).
    """
    import torch

    saved = (P._rademacher, torch.randint_like, torch.rand, torch.randn_like, torch.distributions.Normal.sample)
    probes = list(probes or [])
    rand = list(rand or [])
    normal = list(normal or [])
    randn_like = list(randn_like or [])
    cyc = {"p": 0}

    def next_probe():
        v = probes[cyc["p"] % len(probes)]
        cyc["p"] += 1
        return v

    if probes:
        P._rademacher = lambda shape_or_tensor, device=None: next_probe().clone()
        torch.randint_like = lambda y, lo, hi, **kw: ((next_probe() + 1) / 2).to(y.dtype)
    if rand:
        torch.rand = lambda *a, **kw: rand.pop(0).clone()
    if randn_like:
        torch.randn_like = lambda x, **kw: randn_like.pop(0).clone()
    if normal:
        torch.distributions.Normal.sample = lambda self, shape=(): normal.pop(0).clone()
    try:
        yield
    finally:
        (P._rademacher, torch.randint_like, torch.rand, torch.randn_like, torch.distributions.Normal.sample) = saved
