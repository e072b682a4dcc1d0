"""TEST INFRASTRUCTURE ONLY -- never imported by the product package (gpmdm_b200/).

Imports the UNMODIFIED reference (`/root/reference/gpmdm/{gpmdm,gpmdm_pf}.py`) in the build
container so that (a) the oracle restatement (`oracle/gpmdm_oracle.py`) can be pinned against
it and (b) golden vectors can be generated (`oracle/make_golden.py`).  The reference tree does
not exist on the GPU box, so nothing under `tests/ -m gpu`, `bench.py` or `smoke()` may call
`load_reference()`; CPU tests that need it skip when `/root/reference` is absent.

The reference imports two annotation/printing-only packages that are not installed here
(`torchtyping` -- gpmdm.py:9, gpmdm_pf.py:2; `termcolor` -- gpmdm.py:14).  Two stub modules are
registered in `sys.modules` before the import; the reference files themselves are not touched.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GPMDM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gpmdm", "gpmdm_pf.py"))


def _install_stubs() -> None:
    if "torchtyping" not in sys.modules:
        m = types.ModuleType("torchtyping")

        class TensorType:  # only ever used as `TensorType["N", "d"]` inside annotations
            def __class_getitem__(cls, item):
                return cls

        m.TensorType = TensorType
        sys.modules["torchtyping"] = m
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.cprint = lambda *a, **k: None
        m.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = m


def load_reference():
    """Return the reference's `gpmdm` package (classes `GPMDM`, `GPMDM_PF`)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    pkg = importlib.import_module("gpmdm")
    if not os.path.realpath(pkg.__file__).startswith(os.path.realpath(REFERENCE_ROOT)):
        raise RuntimeError(f"`import gpmdm` resolved to {pkg.__file__}, not the reference")
    return pkg


class InjectedDraws(contextlib.AbstractContextManager):
    """Patch, as seen from the reference's `gpmdm_pf` module, the four RNG entry points so that
    they consume caller-supplied raw draws with the exact semantics torch 2.11 CPU has
    (each identity is asserted against the true torch stream in
    tests/test_rng_identities.py):

      torch.multinomial(dist[P,C], 1, True)   gpmdm_pf.py:150 -> argmax_j dist[p,j] / E[p,j]
      torch.normal(mean, std)                 gpmdm_pf.py:168 -> eps * std + mean (class order)
      torch.multinomial(w[P], P, True)        gpmdm_pf.py:211 -> first j with cdf_j >= u
      torch.randint(0, n, (k,))               gpmdm_pf.py:113 -> supplied indices

    `E` [P,C] are Exp(1) draws, `eps` [P,d] standard normals indexed by PARTICLE (row p is used
    for particle p whatever class it lands in), `u` [P] uniforms, `init_idx` a list of index
    tensors (one per class) for `_init_particles`.
    """

    def __init__(self, pf_module, E=None, eps=None, u=None, init_idx=None):
        import torch

        self._torch = torch
        self._mod = pf_module
        self.E, self.eps, self.u = E, eps, u
        self.init_idx = list(init_idx) if init_idx is not None else None
        self._class_of_call = None
        self.record = {}

    # the reference module does `import torch` and calls torch.<fn>; give it a proxy module
    def __enter__(self):
        torch = self._torch
        outer = self

        class _TorchProxy(types.ModuleType):
            def __getattr__(self, name):
                return getattr(torch, name)

        proxy = _TorchProxy("torch_proxy")

        def multinomial(inp, num_samples, replacement=False, **kw):
            if inp.dim() == 2:  # class transition
                assert num_samples == 1 and replacement
                q = inp / outer.E
                res = torch.argmax(q, dim=-1, keepdim=True)
                outer.record["new_classes"] = res.squeeze(-1).clone()
                outer._new_classes = res.squeeze(-1)
                return res
            assert replacement and num_samples == inp.numel()
            cdf = sequential_cdf(inp)
            anc = torch.searchsorted(cdf, outer.u, right=False)
            outer.record["ancestors"] = anc.clone()
            outer.record["cdf"] = cdf.clone()
            return anc

        def normal(mean, std, **kw):
            # called once per class, in class order, rows = particles of that class ascending
            cls = outer.record.setdefault("_normal_calls", 0)
            outer.record["_normal_calls"] = cls + 1
            rows = torch.nonzero(outer._new_classes == cls).squeeze(-1)
            z = outer.eps[rows]
            outer.record.setdefault("dyn_mean", {})[cls] = (rows.clone(), mean.clone())
            outer.record.setdefault("dyn_std", {})[cls] = (rows.clone(), std.clone())
            return z * std + mean

        def randint(low, high, size, **kw):
            idx = outer.init_idx.pop(0)
            assert tuple(idx.shape) == tuple(size) and (idx.numel() == 0 or int(idx.max()) < high)
            return idx

        if self.E is not None:
            proxy.multinomial = multinomial
        elif self.u is not None:
            proxy.multinomial = multinomial
        if self.eps is not None:
            proxy.normal = normal
        if self.init_idx is not None:
            proxy.randint = randint
        self._saved = self._mod.torch
        self._mod.torch = proxy
        return self

    def __exit__(self, *exc):
        self._mod.torch = self._saved
        return False


def sequential_cdf(w):
    """cdf the CPU multinomial kernel builds (ATen MultinomialKernel.cpp): running sum in index
    order, divide by the total, force the last entry to 1."""
    import torch

    c = torch.cumsum(w.to(torch.float64), 0)  # 1-D CPU cumsum is a sequential loop
    c = c / c[-1]
    c[-1] = 1.0
    return c
