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
    monkeypatch.setattr(G, "_TRMM_LEAF", 8)
    gen = torch.Generator().manual_seed(n)
    K = spd(n, gen)
    L = torch.linalg.cholesky(K)
    garbage = torch.triu(torch.randn(n, n, dtype=F64, generator=gen), 1)  # the strict upper part must not be read
    Kinv = G.spd_inverse_from_cholesky(L + garbage)
    ref = torch.cholesky_inverse(L)
    assert float((Kinv - ref).abs().max()) < 1e-13 * float(ref.abs().max())
    assert torch.equal(Kinv, Kinv.t())
    assert float((Kinv @ K - torch.eye(n, dtype=F64)).abs().max()) < 1e-10


@pytest.mark.parametrize("n", [1, 9, 40, 131])
def test_products_with_a_triangular_operand_recurse_on_the_triangle(n, monkeypatch):
    """X A, C X (in place and into a strided output) and X^T X as block recursions that skip the structural zeros: equal
    to the dense products to rounding, for sizes below, at and far above the leaf."""
    monkeypatch.setattr(G, "_TRMM_LEAF", 8)
    gen = torch.Generator().manual_seed(100 + n)
    A = torch.tril(torch.randn(n, n, dtype=F64, generator=gen))
    X = torch.randn(23, n, dtype=F64, generator=gen)
    Z = torch.randn(n, 17, dtype=F64, generator=gen)
    tol = 1e-13 * n * float(A.abs().max())
    Y = X.clone()
    G._mm_right_lower_(Y, A)
    assert float((Y - X @ A).abs().max()) < tol * float(X.abs().max())
    Y = Z.clone()
    G._mm_left_lower_(A, Y)
    assert float((Y - A @ Z).abs().max()) < tol * float(Z.abs().max())
    big = torch.zeros(30, 2 * n + 3, dtype=F64)
    G._mm_right_lower_into(X.t().contiguous().t(), A, big[4:27, 2:2 + n])  # column-major input, strided output
    assert float((big[4:27, 2:2 + n] - X @ A).abs().max()) < tol * float(X.abs().max())
    assert float(big[:4].abs().max()) == 0.0 and float(big[:, 2 + n:].abs().max()) == 0.0
    out = torch.eye(n, dtype=F64)
    G._syrk_t_add_(X, out)
    assert float((out - (torch.eye(n, dtype=F64) + X.t() @ X)).abs().max()) < 1e-13 * 23 * float(X.abs().max()) ** 2
    assert torch.equal(out, out.t())


# ---- factor precompute without dense intermediates (gpmdm.py:1284-1305 -> GPMDM._factor_block) -------------------------
@pytest.mark.parametrize("n", [1, 7, 256, 300, 700])
def test_tril_inverse_in_place(n, monkeypatch):
    monkeypatch.setattr(G, "_TRINV_LEAF", 64)
    monkeypatch.setattr(G, "_TRMM_LEAF", 16)
    gen = torch.Generator().manual_seed(n)
    L = torch.linalg.cholesky(spd(n, gen))
    out = G.tril_inverse_inplace(L.clone())
    ref = torch.linalg.solve_triangular(L, torch.eye(n, dtype=F64), upper=False)
    assert float((out - ref).abs().max()) < 1e-13 * float(ref.abs().max())
    assert float(torch.triu(out, 1).abs().max()) == 0.0 if n > 1 else True


@pytest.mark.parametrize("tri", [True, False])
@pytest.mark.parametrize("n", [5, 256, 300, 777])
def test_quadform_panels_straight_from_the_triangular_inverse(n, tri, monkeypatch):
    """Panels built by one GEMM per column panel from L^-1 hold exactly the layout gpmdm_pack_quadform_f64 writes from a
    dense inverse (include/gpmdm_b200.h: gpmdm_gp_block), and k^T Q k == k^T K^-1 k."""
    monkeypatch.setattr(G, "_TRINV_LEAF", 64)
    monkeypatch.setattr(G, "_TRMM_LEAF", 16)
    monkeypatch.setattr(G, "_PANEL_ROW_BLOCK", 96)  # several row blocks below a panel's diagonal block
    gen = torch.Generator().manual_seed(n + 1)
    K = spd(n, gen) / n
    Kinv = torch.linalg.inv(K)
    Linv = G.tril_inverse_inplace(torch.linalg.cholesky(K))
    n_pad = (n + 255) // 256 * 256
    panels = G.quadform_panels_from_tril_inverse(Linv, n_pad, tri)
    assert panels.numel() == G.quadform_panel_elems(n_pad, tri)
    # the layout, from a numpy-style construction of the same panels out of the dense inverse
    Q = Kinv.clone() if not tri else 2 * torch.tril(Kinv, -1) + torch.diag(torch.diagonal(Kinv))
    Qp = torch.zeros(n_pad, n_pad, dtype=F64)
    Qp[:n, :n] = Q
    want = torch.cat([torch.cat([Qp[(256 * t if tri else 0):, 256 * t:256 * t + 256],
                                 torch.zeros(n_pad - (256 * t if tri else 0), 4, dtype=F64)], 1)
                      for t in range(n_pad // 256)], 0).reshape(-1)
    assert float((panels - want).abs().max()) < 1e-12 * float(Kinv.abs().max())
    assert bool((panels[want == 0] == 0).all())  # padding, pitch columns and the upper part of diagonal blocks
    back = G.dense_from_quadform_panels(panels, n, n_pad, tri)
    assert float((back - Kinv).abs().max()) < 1e-12 * float(Kinv.abs().max())
    k = torch.rand(n, dtype=F64, generator=gen)
    assert abs(float(k @ Q @ k) - float(k @ Kinv @ k)) < 1e-10 * abs(float(k @ Kinv @ k))


def test_factor_recipe_reproduces_the_reference_inverse_on_the_goldens():
    """K_y^-1 from (lower Cholesky, in-place triangular inverse, panel GEMMs) against the `Ky_inv` the unmodified reference
    computed for the golden models (gpmdm.py:1287-1289: upper Cholesky, torch.inverse, U^-1 U^-T): 1e-9 of max |K^-1|."""
    from oracle import gpmdm_oracle as orc
    from tests.helpers import GOLDEN_CASES, Golden, t64

    for name in GOLDEN_CASES:
        g = Golden(name)
        K = orc.y_kernel(g.spec, g.spec.X, g.spec.X)
        Linv = G.tril_inverse_inplace(torch.linalg.cholesky(K))
        n = K.shape[0]
        n_pad = (n + 255) // 256 * 256
        own = G.dense_from_quadform_panels(G.quadform_panels_from_tril_inverse(Linv, n_pad, True), n, n_pad, True)
        ref = t64(g.z["Ky_inv"])
        assert float((own - ref).abs().max()) < 1e-9 * float(ref.abs().max()), name
