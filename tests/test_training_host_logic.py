"""CPU: the host-side linear algebra of the training step (gpmdm_b200/gpmdm.py) -- the closed-form backward of the
NLL's log det / trace node against autograd through the factorisation (what the reference does, gpmdm.py:576-589,
:617-628), and the GEMM-shaped block recursions that build K^-1 from the Cholesky factor.  Pure torch: runs without a
GPU and without the CUDA library."""
import pytest
import torch

from gpmdm_b200 import gpmdm as G

F64 = torch.float64


def spd(n, gen):
    R = torch.randn(n, n, dtype=F64, generator=gen)
    return R @ R.t() + n * torch.eye(n, dtype=F64)


def reference_scalars(K, T):
    """gpmdm.py:576-589: U = chol_upper(K); logdet = 2 sum log diag U; tr(K^-1 T T^T) via U^-1."""
    U = torch.linalg.cholesky(K, upper=True)
    Z = torch.linalg.solve_triangular(U.t(), T, upper=False)
    return 2 * torch.log(torch.diagonal(U)).sum(), (Z * Z).sum()


@pytest.mark.parametrize("offsets", [None, [0, 17, 17, 40, 64]])
def test_logdet_trace_closed_form_backward_matches_autograd(offsets, monkeypatch):
    monkeypatch.setattr(G, "_INV_LEAF", 16)  # exercise the recursions at this size
    gen = torch.Generator().manual_seed(0)
    n = 64
    K = torch.zeros(n, n, dtype=F64)
    for a, b in ([(0, n)] if offsets is None else zip(offsets[:-1], offsets[1:])):
        if b > a:
            K[a:b, a:b] = spd(b - a, gen)
    T = torch.randn(n, 5, dtype=F64, generator=gen)
    K1, T1 = K.clone().requires_grad_(), T.clone().requires_grad_()
    K2, T2 = K.clone().requires_grad_(), T.clone().requires_grad_()
    l1, t1 = G._LogdetTrace.apply(K1, T1, offsets)
    (31.0 * l1 + 0.5 * t1).backward()
    l2, t2 = reference_scalars(K2, T2)
    (31.0 * l2 + 0.5 * t2).backward()
    l1, t1, l2, t2 = (float(v.detach()) for v in (l1, t1, l2, t2))
    assert abs(l1 - l2) < 1e-11 * abs(l2) and abs(t1 - t2) < 1e-11 * abs(t2)
    mask = torch.ones_like(K) if offsets is None else (K != 0).to(F64)  # off-block gradients are never read
    assert float(((K1.grad - K2.grad) * mask).abs().max()) < 1e-12 * float(K2.grad.abs().max())
    assert float((T1.grad - T2.grad).abs().max()) < 1e-12 * float(T2.grad.abs().max())


@pytest.mark.parametrize("n", [1, 15, 16, 100, 257])
def test_spd_inverse_from_cholesky_block_recursion(n, monkeypatch):
    monkeypatch.setattr(G, "_INV_LEAF", 16)
    gen = torch.Generator().manual_seed(n)
    K = spd(n, gen)
    L = torch.linalg.cholesky(K)
    garbage = torch.triu(torch.randn(n, n, dtype=F64, generator=gen), 1)  # the strict upper part must not be read
    Kinv = G.spd_inverse_from_cholesky(L + garbage)
    ref = torch.cholesky_inverse(L)
    assert float((Kinv - ref).abs().max()) < 1e-13 * float(ref.abs().max())
    assert torch.equal(Kinv, Kinv.t())
    assert float((Kinv @ K - torch.eye(n, dtype=F64)).abs().max()) < 1e-10
