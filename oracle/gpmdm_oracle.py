"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GPMDM particle-filter path.

A block-structured CPU restatement (torch CPU fp64 ops, the same ATen/MKL calls the reference
makes) of the reference algorithm in `/root/reference/gpmdm/gpmdm_pf.py` and the GP machinery of
`/root/reference/gpmdm/gpmdm.py`.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product
package `gpmdm_b200/` never does (it has no CPU path at all).

PARITY PIN: the reference ships no tests or golden vectors ("parity unpinned" by the reference's
own suite).  This restatement is pinned instead against the UNMODIFIED reference executed in the
build container (`oracle/ref_shim.py`, `oracle/make_golden.py` -> `tests/golden/*.npz`, checked by
`tests/test_oracle_vs_golden.py` everywhere and `tests/test_oracle_vs_reference.py` where
`/root/reference` exists).

Differences from the reference, all value-preserving (checked by the tests above):
  * the per-class dynamics factors are the N_c x N_c diagonal blocks; the reference stores dense
    Nx x Nx matrices whose off-class part is exactly 1e6*I and multiplies it by an exactly-zero
    masked cross-kernel (gpmdm.py:1061, 1299-1305);
  * `alpha = K^-1 Y` is precomputed where the reference recomputes `Y^T K^-1` every call
    (gpmdm.py:957, 1064);
  * the per-particle Python loop for the log-likelihood (gpmdm_pf.py:188-192) also exists in a
    vectorised closed form (`log_likelihoods_fused`), equal to the loop to ~5e-16 relative.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional

import numpy as np
import torch

F64 = torch.float64

# gpmdm_pf.py:5 -- evaluated in float32 by the reference (torch.tensor(2*pi) defaults to fp32)
LOG_2PI_F32 = torch.log(torch.tensor(2 * 3.14159265358979323846))


def c32_constant(D: int) -> float:
    """`0.5 * self._gpmdm.D * _LOG_2PI` (gpmdm_pf.py:191): python float * fp32 0-d tensor -> fp32."""
    return float(0.5 * D * LOG_2PI_F32)


# ------------------------------------------------------------------------------------------------
# model description
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class ModelSpec:
    """Everything the filter needs from a trained `GPMDM` (gpmdm.py:96-237, 281-309)."""

    X: torch.Tensor  # [N, d] latent coords, class-major / sequence-major / time (gpmdm.py:301-309)
    Y: torch.Tensor  # [N, D] observations (meanY = 0, gpmdm.py:791)
    seq_lengths: List[List[int]]  # per class, per sequence
    y_log_lengthscales: torch.Tensor  # [d]
    y_log_lambdas: torch.Tensor  # [D]
    y_log_sigma_n: torch.Tensor  # []
    x_log_lengthscales: torch.Tensor  # [d]
    x_log_lambdas: torch.Tensor  # [d]
    x_log_sigma_n: torch.Tensor  # []
    x_log_lin_coeff: torch.Tensor  # [d+1]
    sigma_n_num_Y: float = 0.0
    sigma_n_num_X: float = 0.0

    @property
    def n_classes(self):
        return len(self.seq_lengths)

    @property
    def d(self):
        return self.X.shape[1]

    @property
    def D(self):
        return self.Y.shape[1]

    @property
    def N(self):
        return self.X.shape[0]

    def class_row_ranges(self):
        """[start, end) rows of X per class -- `get_X_for_class` (gpmdm.py:906-921)."""
        out, s = [], 0
        for lens in self.seq_lengths:
            n = sum(lens)
            out.append((s, s + n))
            s += n
        return out

    def class_pair_ranges(self):
        """[start, end) rows of Xin/Xout per class -- block offsets of `get_M` (gpmdm.py:311-340)."""
        out, s = [], 0
        for lens in self.seq_lengths:
            n = sum(l - 1 for l in lens)
            out.append((s, s + n))
            s += n
        return out

    @staticmethod
    def from_reference(model) -> "ModelSpec":
        """Snapshot an (unmodified) reference `GPMDM` or the product `GPMDM` (same attributes)."""
        lens = [[int(len(s)) for s in cls] for cls in model.class_aware_observations_list]
        Y = torch.tensor(np.concatenate(model.observations_list, 0), dtype=F64)
        g = lambda p: p.detach().to("cpu", F64).clone()
        return ModelSpec(
            X=g(model.X), Y=Y, seq_lengths=lens,
            y_log_lengthscales=g(model.y_log_lengthscales), y_log_lambdas=g(model.y_log_lambdas),
            y_log_sigma_n=g(model.y_log_sigma_n), x_log_lengthscales=g(model.x_log_lengthscales),
            x_log_lambdas=g(model.x_log_lambdas), x_log_sigma_n=g(model.x_log_sigma_n),
            x_log_lin_coeff=g(model.x_log_lin_coeff),
            sigma_n_num_Y=float(model.sigma_n_num_Y), sigma_n_num_X=float(model.sigma_n_num_X))


# ------------------------------------------------------------------------------------------------
# kernels (gpmdm.py:381-548)
# ------------------------------------------------------------------------------------------------
def weighted_distances(X1, X2, log_lengthscales):
    """gpmdm.py:483-517: ||a||^2 + ||b||^2 - 2 a.b on lengthscale-divided inputs (no 1/2)."""
    ls = torch.exp(log_lengthscales)
    A = X1 / ls
    A2 = torch.sum(A.mul(A), dim=1, keepdim=True)
    B = X2 / ls
    B2 = torch.sum(B.mul(B), dim=1, keepdim=True)
    return A2 + B2.transpose(0, 1) - 2 * torch.matmul(A, B.transpose(0, 1))


def rbf_kernel(X1, X2, log_lengthscales, log_sigma_n, sigma_n_num=0.0, flg_noise=True):
    """gpmdm.py:436-481."""
    K = torch.exp(-weighted_distances(X1, X2, log_lengthscales))
    if flg_noise:
        n = X1.shape[0]
        K = K + torch.exp(log_sigma_n) ** 2 * torch.eye(n, dtype=X1.dtype) \
            + sigma_n_num ** 2 * torch.eye(n, dtype=X1.dtype)
    return K


def lin_kernel(X1, X2, log_lin_coeff):
    """gpmdm.py:520-548: [a,1] diag(c^2) [b,1]^T."""
    Sigma = torch.diag(torch.exp(log_lin_coeff) ** 2)
    A = torch.cat([X1, torch.ones(X1.shape[0], 1, dtype=X1.dtype)], 1)
    B = torch.cat([X2, torch.ones(X2.shape[0], 1, dtype=X2.dtype)], 1)
    return torch.matmul(A, torch.matmul(Sigma, B.transpose(0, 1)))


def y_kernel(m: ModelSpec, X1, X2, flg_noise=True):
    """gpmdm.py:381-406."""
    return rbf_kernel(X1, X2, m.y_log_lengthscales, m.y_log_sigma_n, m.sigma_n_num_Y, flg_noise)


def x_kernel(m: ModelSpec, X1, X2, flg_noise=True):
    """gpmdm.py:408-434."""
    return rbf_kernel(X1, X2, m.x_log_lengthscales, m.x_log_sigma_n, m.sigma_n_num_X, flg_noise) \
        + lin_kernel(X1, X2, m.x_log_lin_coeff)


def x_diag_kernel(m: ModelSpec, X, flg_noise=False):
    """gpmdm.py:1070-1101."""
    Sigma = torch.diag(torch.exp(m.x_log_lin_coeff) ** 2)
    Xa = torch.cat([X, torch.ones(X.shape[0], 1, dtype=X.dtype)], 1)
    out = torch.ones(X.shape[0], dtype=X.dtype) + torch.sum(torch.matmul(Xa, Sigma) * Xa, dim=1)
    if flg_noise:
        out = out + torch.exp(m.x_log_sigma_n) ** 2 + m.sigma_n_num_X ** 2
    return out


def xin_xout(m: ModelSpec):
    """`get_Xin_Xout_matrices` for target='full', back_step=1 (gpmdm.py:668-686): drop the last /
    first frame of every sequence.  The filter is only meaningful for this mode."""
    ins, outs, s = [], [], 0
    for lens in m.seq_lengths:
        for L in lens:
            ins.append(m.X[s:s + L - 1])
            outs.append(m.X[s + 1:s + L])
            s += L
    return torch.cat(ins, 0), torch.cat(outs, 0)


def inverse_via_upper_cholesky(K):
    """The reference's inverse recipe (gpmdm.py:1287-1289): U = chol_upper(K); U^-1 U^-T."""
    U, _info = torch.linalg.cholesky_ex(K, upper=True)
    U_inv = torch.inverse(U)
    return torch.matmul(U_inv, U_inv.t())


# ------------------------------------------------------------------------------------------------
# precomputed factors (gpmdm.py:1284-1305), block form
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Factors:
    Ky_inv: torch.Tensor  # [N, N]
    alpha_y: torch.Tensor  # [N, D] = Ky_inv @ Y
    Kx_inv_blocks: List[torch.Tensor]  # per class [N_c, N_c]
    alpha_x: List[torch.Tensor]  # per class [N_c, d] = block @ Xout_c
    Xin: torch.Tensor
    Xout: torch.Tensor


def precompute_factors(m: ModelSpec, Ky_inv: Optional[torch.Tensor] = None,
                       Kx_inv_blocks: Optional[List[torch.Tensor]] = None) -> Factors:
    """Block restatement of `_precompute_kernel_inverses` (gpmdm.py:1284-1305).  Inverses taken
    from the reference (dense `Ky_inv`, diagonal blocks of `Kx_inv_class[c]`) may be injected."""
    Xin, Xout = xin_xout(m)
    if Ky_inv is None:
        Ky_inv = inverse_via_upper_cholesky(y_kernel(m, m.X, m.X))
    blocks = []
    for c, (a, b) in enumerate(m.class_pair_ranges()):
        if Kx_inv_blocks is not None:
            blocks.append(Kx_inv_blocks[c])
            continue
        Kc = x_kernel(m, Xin[a:b], Xin[a:b]) + 1e-6 * torch.eye(b - a, dtype=F64)  # :1301-1302
        blocks.append(inverse_via_upper_cholesky(Kc))
    alpha_x = [torch.matmul(blocks[c], Xout[a:b]) for c, (a, b) in enumerate(m.class_pair_ranges())]
    return Factors(Ky_inv=Ky_inv, alpha_y=torch.matmul(Ky_inv, m.Y), Kx_inv_blocks=blocks,
                   alpha_x=alpha_x, Xin=Xin, Xout=Xout)


def precompute_factors_on(m: ModelSpec, device) -> Factors:
    """`precompute_factors` with the O(N^3) linear algebra (the reference recipe gpmdm.py:1287-1289, 1301-1305:
    upper Cholesky, triangular inverse, U^-1 U^-T) evaluated by plain torch on `device`, returned on the CPU.
    For N = 20 000 the CPU recipe takes minutes; tests and bench.py use this to SET UP the oracle at the benchmark
    sizes (never inside a timed region).  The cross-kernel and every prediction stay on the CPU."""
    def inv(K):
        U, _ = torch.linalg.cholesky_ex(K, upper=True)
        Ui = torch.linalg.solve_triangular(U, torch.eye(K.shape[0], dtype=K.dtype, device=K.device), upper=True)
        del U
        return Ui @ Ui.t()

    X = m.X.to(device)
    Ky = torch.exp(-weighted_distances(X, X, m.y_log_lengthscales.to(device)))
    Ky.diagonal().add_(float(torch.exp(m.y_log_sigma_n) ** 2 + m.sigma_n_num_Y ** 2))
    Ky_inv = inv(Ky).cpu()
    del Ky
    Xin, _ = xin_xout(m)
    blocks = []
    for a, b in m.class_pair_ranges():
        Kc = x_kernel(m, Xin[a:b], Xin[a:b]).to(device)
        Kc.diagonal().add_(1e-6)
        blocks.append(inv(Kc).cpu())
    return precompute_factors(m, Ky_inv=Ky_inv, Kx_inv_blocks=blocks)


# ------------------------------------------------------------------------------------------------
# GP prediction (gpmdm.py:923-963, 1032-1068)
# ------------------------------------------------------------------------------------------------
def map_x_dynamics_for_class(m: ModelSpec, f: Factors, Xstar, c: int):
    """gpmdm.py:1032-1068 on the class-c block.  Returns (mean [P,d], var [P,d], q [P], prior [P])."""
    a, b = m.class_pair_ranges()[c]
    Ks = x_kernel(m, f.Xin[a:b], Xstar, False)  # [N_c, P]
    prior = x_diag_kernel(m, Xstar, False)
    mean = torch.matmul(Ks.t(), f.alpha_x[c])
    q = torch.sum(torch.matmul(Ks.t(), f.Kx_inv_blocks[c]) * Ks.t(), dim=1)
    common = prior - q
    lam = torch.exp(m.x_log_lambdas) ** -2
    return mean, common.unsqueeze(1) * lam.unsqueeze(0), q, prior


def map_x_to_y(m: ModelSpec, f: Factors, Xstar):
    """gpmdm.py:923-963.  Returns (mean [P,D], var [P,D], v [P]) with var = v * lambda^-2."""
    Ks = y_kernel(m, m.X, Xstar, False)  # [N, P]
    mean = torch.matmul(Ks.t(), f.alpha_y)
    v = torch.ones(Xstar.shape[0], dtype=F64) - torch.sum(torch.matmul(Ks.t(), f.Ky_inv) * Ks.t(), dim=1)
    lam = torch.exp(m.y_log_lambdas) ** -2
    return mean, v.unsqueeze(1) * lam.unsqueeze(0), v


def log_likelihoods_loop(mean, var, z, D: int):
    """Verbatim per-particle loop of gpmdm_pf.py:188-192 (small P only)."""
    out = torch.zeros(mean.shape[0], dtype=F64)
    for i in range(mean.shape[0]):
        mu_term = -0.5 * torch.sum((z - mean[i]) ** 2 / var[i] + torch.log(var[i]))
        sigma_term = torch.sum(-torch.log(torch.sqrt(var[i])))
        out[i] = mu_term + sigma_term - 0.5 * D * LOG_2PI_F32
    return out


def log_likelihoods_fused(mean, v, z, y_log_lambdas):
    """Closed form of the loop above, the form the CUDA epilogue uses:
    ll = -(1/(2v)) sum_j lam_j^2 (z_j-mu_j)^2 - D log v + 2 sum_j log lam_j - c32."""
    D = mean.shape[1]
    lam2 = torch.exp(y_log_lambdas) ** 2
    S = torch.sum(lam2.unsqueeze(0) * (z.unsqueeze(0) - mean) ** 2, dim=1)
    return -0.5 * S / v - D * torch.log(v) + 2.0 * torch.sum(y_log_lambdas) - c32_constant(D)


# ------------------------------------------------------------------------------------------------
# filter stages (gpmdm_pf.py:137-262)
# ------------------------------------------------------------------------------------------------
def transition(classes, T, E):
    """gpmdm_pf.py:137-151 with `torch.multinomial(dist, 1)` == argmax(dist / Exp(1)) (App. B)."""
    dist = T[classes]  # == one_hot(classes) @ T exactly (rows with a single 1.0)
    return torch.argmax(dist / E, dim=-1)


def dynamics_draw(m: ModelSpec, f: Factors, states, new_classes, eps):
    """gpmdm_pf.py:153-168.  eps [P,d] is indexed by particle.  Returns x', mean, var (all [P,d])."""
    x = states.clone()
    mean_all = torch.zeros_like(states)
    var_all = torch.zeros_like(states)
    for c in range(m.n_classes):
        rows = torch.nonzero(new_classes == c).squeeze(-1)
        if rows.numel() == 0:
            continue
        mean, var, _, _ = map_x_dynamics_for_class(m, f, states[rows], c)
        mean_all[rows], var_all[rows] = mean, var
        x[rows] = eps[rows] * torch.sqrt(var) + mean  # torch.normal == randn*std + mean (mul, add)
    return x, mean_all, var_all


def normalize(ll):
    """gpmdm_pf.py:200-204."""
    lw = ll - torch.max(ll)
    w = torch.exp(lw)
    return lw, w / torch.sum(w)


def sequential_cdf(w):
    """ATen CPU multinomial: running sum in index order, / total, last := 1 (App. B)."""
    c = torch.cumsum(w.to(F64), 0)
    c = c / c[-1]
    c[-1] = 1.0
    return c


def resample(w, u):
    """gpmdm_pf.py:206-213: ancestors a_s = first j with cdf_j >= u_s."""
    return torch.searchsorted(sequential_cdf(w), u, right=False)


def class_probabilities(ll, lw, classes_post, C: int):
    """gpmdm_pf.py:224-248 (post-resample classes, pre-resample ll+lw)."""
    g = ll + lw
    g = g - torch.max(g)
    e = torch.exp(g)
    L = torch.zeros(C, dtype=F64)
    for i in range(C):
        L[i] = torch.sum(e[classes_post == i])
    return L / torch.sum(L)


def weighted_log_sum(ll, lw):
    """`_weighted_sum_from_log_space` (gpmdm_pf.py:302-312) as used by `log_likelihood` (:215)."""
    g = lw + ll
    return torch.sum(torch.exp(g - torch.max(g)))


def current_state_mean(states_post, w):
    """gpmdm_pf.py:256-262."""
    return torch.sum(states_post * w.unsqueeze(-1), dim=0)


def divide_into_n_parts(x: int, n: int):
    """gpmdm_pf.py:287-292 (first x mod n parts get one extra)."""
    g, r = divmod(x, n)
    return [g + (1 if i < r else 0) for i in range(n)]


class FilterOracle:
    """Stateful restatement of `GPMDM_PF` driven by injected raw draws (E, eps, u)."""

    def __init__(self, m: ModelSpec, T, num_particles: int, init_idx, factors: Optional[Factors] = None):
        self.m = m
        self.f = factors if factors is not None else precompute_factors(m)
        self.T = T.to(F64)  # gpmdm_pf.py:71
        if self.T.shape[0] != m.n_classes:
            raise ValueError("Number of classes in the GPMDM model and the Markov model do not match")
        self.P = num_particles
        parts = divide_into_n_parts(num_particles, m.n_classes)
        xs, cs = [], []
        for c, (a, b) in enumerate(m.class_row_ranges()):
            xs.append(m.X[a:b][init_idx[c]].clone())
            cs += [c] * parts[c]
        self.states = torch.cat(xs, 0)
        self.classes = torch.tensor(cs, dtype=torch.int64)
        self.ll = torch.zeros(num_particles, dtype=F64)
        self.lw = torch.zeros(num_particles, dtype=F64)
        self.w = torch.ones(num_particles, dtype=F64) / num_particles
        self.trace = {}

    def update(self, z, E, eps, u, loop_ll: bool = False):
        m, f = self.m, self.f
        z = torch.as_tensor(np.asarray(z), dtype=F64)  # gpmdm_pf.py:123
        c_new = transition(self.classes, self.T, E)
        x_new, dmean, dvar = dynamics_draw(m, f, self.states, c_new, eps)
        mu, var, v = map_x_to_y(m, f, x_new)
        ll = log_likelihoods_loop(mu, var, z, m.D) if loop_ll else log_likelihoods_fused(mu, v, z, m.y_log_lambdas)
        lw, w = normalize(ll)
        anc = resample(w, u)
        self.trace = dict(c_new=c_new, dyn_mean=dmean, dyn_var=dvar, x_new=x_new, mu=mu, v=v, ll=ll,
                          lw=lw, w=w, anc=anc)
        self.states, self.classes = x_new[anc], c_new[anc]
        self.ll, self.lw, self.w = ll, lw, w

    def class_probabilities(self):
        return class_probabilities(self.ll, self.lw, self.classes, self.m.n_classes)

    def get_most_likely_class(self):
        return int(torch.argmax(self.class_probabilities()))

    def current_state_mean(self):
        return current_state_mean(self.states, self.w)

    def log_likelihood(self):
        return float(weighted_log_sum(self.ll, self.lw))


# ------------------------------------------------------------------------------------------------
# training side (gpmdm.py:550-628): kernel build + closed-form gradient terms
# ------------------------------------------------------------------------------------------------
def class_mask(m: ModelSpec):
    """Dense 0/1 `M` of `get_M` (gpmdm.py:311-340) -- small N only."""
    n = m.class_pair_ranges()[-1][1]
    M = torch.zeros(n, n, dtype=F64)
    for a, b in m.class_pair_ranges():
        M[a:b, a:b] = 1.0
    return M


def y_neg_log_likelihood(m: ModelSpec, X=None):
    """gpmdm.py:550-589."""
    X = m.X if X is None else X
    K = y_kernel(m, X, X)
    U, _ = torch.linalg.cholesky_ex(K, upper=True)
    U_inv = torch.inverse(U)
    Kinv = torch.matmul(U_inv, U_inv.t())
    logdet = 2 * torch.sum(torch.log(torch.diag(U)))
    W2 = torch.diag(torch.exp(m.y_log_lambdas) ** 2)
    YWY = torch.linalg.multi_dot([m.Y, W2, m.Y.t()])
    return m.D / 2 * logdet + 0.5 * torch.trace(torch.mm(Kinv, YWY)) - m.N * 2 * torch.sum(m.y_log_lambdas)


def x_neg_log_likelihood(m: ModelSpec, Xin, Xout):
    """gpmdm.py:591-628."""
    K = x_kernel(m, Xin, Xin) * class_mask(m)
    U, _ = torch.linalg.cholesky_ex(K, upper=True)
    U_inv = torch.inverse(U)
    Kinv = torch.matmul(U_inv, U_inv.t())
    logdet = 2 * torch.sum(torch.log(torch.diag(U)))
    W2 = torch.diag(torch.exp(m.x_log_lambdas) ** 2)
    return m.d / 2 * logdet + 0.5 * torch.trace(torch.linalg.multi_dot([Kinv, Xout, W2, Xout.t()])) \
        - Xin.shape[0] * 2 * torch.sum(m.x_log_lambdas)


# ------------------------------------------------------------------------------------------------
# extended-precision arbiter (layer L2)
# ------------------------------------------------------------------------------------------------
def dynamics_truth_longdouble(m: ModelSpec, f: Factors, Xstar, c: int):
    """mean / var of `map_x_dynamics_for_class` with the cross-kernel, the contraction against the SAME fp64
    `Kx_inv_blocks[c]` / `alpha_x[c]` and the cancellation `prior - q` carried out in numpy longdouble
    (80-bit on x86).  Arbitrates between the oracle and the CUDA path: both approximate this value."""
    ld = np.longdouble
    a, b = m.class_pair_ranges()[c]
    Xin = f.Xin[a:b].numpy().astype(ld)
    Xs = Xstar.numpy().astype(ld)
    ls = np.exp(m.x_log_lengthscales.numpy().astype(ld))
    c2 = np.exp(m.x_log_lin_coeff.numpy().astype(ld)) ** 2
    A, B = Xin / ls, Xs / ls
    dist = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
    K = np.exp(-dist) + (Xin * c2[:-1]) @ Xs.T + c2[-1]  # [N_c, P]
    Kinv = f.Kx_inv_blocks[c].numpy().astype(ld)
    q = np.einsum("ip,ip->p", K, Kinv @ K)
    prior = 1 + (Xs * c2[:-1] * Xs).sum(1) + c2[-1]
    lam = np.exp(m.x_log_lambdas.numpy().astype(ld)) ** -2
    mean = K.T @ f.alpha_x[c].numpy().astype(ld)
    var = (prior - q)[:, None] * lam[None, :]
    return mean, var, prior


def observation_truth_longdouble(m: ModelSpec, f: Factors, Xstar):
    """mean / v of `map_x_to_y` in numpy longdouble against the same fp64 `Ky_inv` / `alpha_y`."""
    ld = np.longdouble
    X = m.X.numpy().astype(ld)
    Xs = Xstar.numpy().astype(ld)
    ls = np.exp(m.y_log_lengthscales.numpy().astype(ld))
    A, B = X / ls, Xs / ls
    K = np.exp(-((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
    q = np.einsum("ip,ip->p", K, f.Ky_inv.numpy().astype(ld) @ K)
    return K.T @ f.alpha_y.numpy().astype(ld), 1 - q
