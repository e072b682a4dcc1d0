"""Seeded synthetic GPMDM workloads.

The reference ships no data (`mocap/` is git-ignored) so every test, golden vector and benchmark in
this repository runs on the generator below.  It mimics what `notebooks/train_gpmdm.ipynb` cell 1
feeds to `GPMDM.add_data`: per class a family of float32 sequences of D joint angles sampled at
30 fps.  Class c follows a smooth periodic trajectory

    Y_c(t) = B_c(t) @ A_c + 0.02 * N(0, 1),   B_c(t) = [sin w_c t, cos w_c t, sin 2 w_c t, cos 2 w_c t]

with A_c ~ N(0,1)^{4 x D}, a class frequency w_c and a random phase per sequence.  numpy's PCG64 is
used so the streams are identical on every machine.
"""
from __future__ import annotations

import dataclasses

import numpy as np

FPS = 30.0


@dataclasses.dataclass
class SyntheticWorkload:
    sequences: list  # list over classes of list of float32 arrays [T, D]
    test_trials: list  # list of (class_index, float32 array [T_test, D])
    D: int
    n_classes: int


def class_frequency(c: int) -> float:
    return 2.0 * np.pi * (0.8 + 0.25 * c)


def make_sequences(n_classes: int, D: int, seqs_per_class: int, frames: int, seed: int = 0,
                   n_test_trials: int = 0, test_frames: int = 150, noise: float = 0.02) -> SyntheticWorkload:
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n_classes, 4, D))

    def one(c, T):
        phase = rng.uniform(0.0, 2.0 * np.pi)
        t = np.arange(T) / FPS
        w = class_frequency(c)
        B = np.stack([np.sin(w * t + phase), np.cos(w * t + phase),
                      np.sin(2 * w * t + 2 * phase), np.cos(2 * w * t + 2 * phase)], 1)
        return (B @ A[c] + noise * rng.standard_normal((T, D))).astype(np.float32)

    seqs = [[one(c, frames) for _ in range(seqs_per_class)] for c in range(n_classes)]
    tests = [(i % n_classes, one(i % n_classes, test_frames)) for i in range(n_test_trials)]
    return SyntheticWorkload(seqs, tests, D, n_classes)


def notebook_hyperparameters(D: int, d: int, sigma_n: float = 1e-2) -> dict:
    """Initial hyper-parameters of `notebooks/train_gpmdm.ipynb` cell 2 (all-ones lambdas,
    lengthscales and linear coefficients; noise std 1e-2), as ctor kwargs."""
    return dict(
        y_lambdas_init=np.ones(D), y_lengthscales_init=np.ones(d), y_sigma_n_init=sigma_n,
        x_lambdas_init=np.ones(d), x_lengthscales_init=np.ones(d), x_sigma_n_init=sigma_n,
        x_lin_coeff_init=np.ones(d + 1),
    )


def markov_matrix(n_classes: int, stay: float = 0.9):
    """`test_gpmdm_pf.ipynb` cell 3 generalised to C classes; float32 like the notebook, so that
    after the reference's cast (gpmdm_pf.py:71) rows hold 0.8999999761581421 etc."""
    import torch

    T = torch.full((n_classes, n_classes), (1.0 - stay) / max(n_classes - 1, 1), dtype=torch.float32)
    T.fill_diagonal_(stay if n_classes > 1 else 1.0)
    return T


def raw_draws(P: int, C: int, d: int, seed: int):
    """Host draws in the reference's consumption order: Exp(1) [P,C] for the
    class transition, N(0,1) [P,d] for the dynamics draw, U(0,1) [P] for resampling.  CPU generator,
    so the same arrays feed the oracle and (after H2D) the CUDA path."""
    import torch

    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    U = torch.rand(P, C, dtype=torch.float64, generator=g)
    E = -torch.log1p(-U)
    eps = torch.randn(P, d, dtype=torch.float64, generator=g)
    u = torch.rand(P, dtype=torch.float64, generator=g)
    return E, eps, u
