"""`GPMDM` -- the model class of the reference (`gpmdm/gpmdm.py`) with the same constructor, data,
training, prediction and save/load API, whose arithmetic on the filter / training hot paths runs in
libgpmdm_sm100a.so (hand-written CUDA for B200) instead of dense torch-CPU expressions.

What is kept verbatim from the reference surface (file:line of the reference):
  ctor :96-237 | set_evaluation_mode :239 | set_training_mode :247 | add_data :281 |
  observations_list :301 | get_y_kernel / get_x_kernel / get_rbf_kernel / get_lin_kernel :381-548 |
  get_y_neg_log_likelihood :550 | get_x_neg_log_likelihood :591 | get_Xin_Xout_matrices :630 |
  gpdm_loss :721 | init_X :762 | get_Y :779 | train_adam :817 | get_latent_sequences :887 |
  get_X_for_class :906 | map_x_to_y :923 | map_x_dynamics_for_class :1032 | save :1307 | load :1350

What differs by design (B200-first):
  * tensors live on the CUDA device; there is no CPU path;
  * the dense 0/1 masks `M`, `M_class[c]` (C+1 dense Nx x Nx matrices, :311-378) are never built --
    class structure is carried as row offsets; `Kx_inv_class[c]` holds the N_c x N_c diagonal block
    of the reference's dense matrix (whose off-class part is exactly 1e6*I and is multiplied by an
    exactly-zero masked cross-kernel, :1061);
  * prediction never materialises the P x N cross-covariance (fused kernels, csrc/gp_predict.cu);
  * `Xin/Xout`, `Y`, `alpha = K^-1 targets` are computed once per precompute, not per call.
"""
from __future__ import annotations

import ctypes
import os
import time
from pathlib import Path
from typing import List, Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import TILE_N, TILE_P, GpModel, GpModelTf32, check, ptr, stream


F64 = torch.float64  # working precision of every kernel buffer and factor (the public dtype may be float32)


def to_tensor(input_array, dtype, device):
    if isinstance(input_array, torch.Tensor):
        return input_array.to(dtype=dtype, device=device)
    return torch.tensor(np.asarray(input_array), dtype=dtype, device=device).clone().detach()


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("gpmdm_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# CUDA kernel-matrix build with hand-written backward (csrc/train_kernels.cu)
# ------------------------------------------------------------------------------------------------
class _KernelBuild(torch.autograd.Function):
    """K = rbf(X, X; l) [+ lin(X, X; c)] + (sigma_n^2 + sigma_num^2) I, optionally * class mask.
    forward: gpmdm_kernel_build_f64; backward: gpmdm_kernel_grad_f64 (closed forms)."""

    @staticmethod
    def forward(ctx, X, log_ls, log_sigma_n, log_lin_coeff, sigma_n_num, class_offsets, flg_noise):
        lib = _cabi.lib()
        # fp32 models: parameters are up-cast here, the kernel matrix and the loss are fp64, gradients go back as fp32
        ctx.in_dtypes = (X.dtype, log_ls.dtype, log_sigma_n.dtype, None if log_lin_coeff is None else log_lin_coeff.dtype)
        X = X.to(F64).contiguous()
        n, d = X.shape
        kind = 0 if log_lin_coeff is None else 1
        ls = torch.exp(log_ls.to(F64)).contiguous()
        c2 = (torch.exp(log_lin_coeff.to(F64)) ** 2).contiguous() if kind else None
        sigma2 = float(torch.exp(log_sigma_n.detach().to(F64)).reshape(-1)[0] ** 2) if flg_noise else 0.0
        noise2 = sigma2 + (float(sigma_n_num) ** 2 if flg_noise else 0.0)
        K = torch.empty(n, n, dtype=F64, device=X.device)
        ncls = 0 if class_offsets is None else class_offsets.numel() - 1
        check(lib.gpmdm_kernel_build_f64(ptr(X), n, d, kind, ptr(ls), ptr(c2), noise2, ptr(class_offsets), ncls,
                                         ptr(K), stream()), "gpmdm_kernel_build_f64")
        ctx.save_for_backward(X, ls, c2 if kind else X.new_empty(0), class_offsets if ncls else X.new_empty(0))
        ctx.meta = (kind, sigma2, ncls, tuple(log_sigma_n.shape))
        return K

    @staticmethod
    def backward(ctx, G):
        lib = _cabi.lib()
        X, ls, c2, offs = ctx.saved_tensors
        kind, sigma2, ncls, sig_shape = ctx.meta
        n, d = X.shape
        G = G.contiguous()
        gX = torch.empty_like(X)
        g_ls = torch.empty(d, dtype=F64, device=X.device)
        g_sig = torch.empty((), dtype=F64, device=X.device)
        g_c = torch.empty(d + 1, dtype=F64, device=X.device) if kind else None
        ws = torch.empty(lib.gpmdm_kernel_grad_workspace_bytes(n, d) // 8 + 1, dtype=torch.float64, device=X.device)
        check(lib.gpmdm_kernel_grad_f64(ptr(X), ptr(G), n, d, kind, ptr(ls), ptr(c2) if kind else None, sigma2,
                                        ptr(offs) if ncls else None, ncls, ptr(gX), ptr(g_ls), ptr(g_sig),
                                        ptr(g_c) if kind else None, ptr(ws), stream()), "gpmdm_kernel_grad_f64")
        tx, tl, ts, tc = ctx.in_dtypes
        return (gX.to(tx), g_ls.to(tl), g_sig.reshape(sig_shape).to(ts), g_c.to(tc) if kind else None, None, None, None)


# ------------------------------------------------------------------------------------------------
# K^-1 from a lower Cholesky factor as GEMM-shaped block recursions (cuBLAS DGEMM runs at ~36 TF/s on B200,
# cuSOLVER's potri at ~10: 534 ms at N = 20 k); the leaves are plain library calls.  Products with a triangular operand
# recurse on the triangle (`_mm_right_lower_`, `_mm_left_lower_`, `_syrk_t_add_`), so that only the leaf blocks multiply
# structural zeros: ~N^3/3 flops for the triangular inverse and for L^-T L^-1 each, instead of 2/3 N^3 with dense GEMMs.
# ------------------------------------------------------------------------------------------------
_INV_LEAF = 1280   # (tools/leaf_sweep.py at N = 20 000: 2560 / 1024 -> 199 ms, 1280 / 512 -> 191 ms for K^-1 from L)
_TRMM_LEAF = 512


def _split(n: int) -> int:
    h = (n // 2 + 127) // 128 * 128
    return h if h < n else n // 2


def _mm_right_lower_(X, A):
    """X <- X A in place; A lower triangular with an explicit zero upper part."""
    k = A.shape[0]
    if k <= _TRMM_LEAF:
        X.copy_(torch.mm(X, A))
        return
    h = _split(k)
    X1, X2 = X[:, :h], X[:, h:]
    _mm_right_lower_(X1, A[:h, :h])
    X1.addmm_(X2, A[h:, :h])  # the original X2
    _mm_right_lower_(X2, A[h:, h:])


def _mm_left_lower_(C, X):
    """X <- C X in place; C lower triangular with an explicit zero upper part."""
    k = C.shape[0]
    if k <= _TRMM_LEAF:
        X.copy_(torch.mm(C, X))
        return
    h = _split(k)
    X1, X2 = X[:h], X[h:]
    _mm_left_lower_(C[h:, h:], X2)
    X2.addmm_(C[h:, :h], X1)  # the original X1
    _mm_left_lower_(C[:h, :h], X1)


def _mm_right_lower_into(Y, A, out):
    """out <- Y A; A lower triangular with an explicit zero upper part (Y, out may be strided views)."""
    k = A.shape[0]
    if k <= _TRMM_LEAF:
        torch.mm(Y, A, out=out)
        return
    h = _split(k)
    _mm_right_lower_into(Y[:, :h], A[:h, :h], out[:, :h])
    out[:, :h].addmm_(Y[:, h:], A[h:, :h])
    _mm_right_lower_into(Y[:, h:], A[h:, h:], out[:, h:])


def _syrk_t_add_(X, out):
    """out += X^T X (out square, symmetric): the off-diagonal blocks are computed once and mirrored."""
    k = X.shape[1]
    if k <= 2 * _TRMM_LEAF:
        out.addmm_(X.t(), X)
        return
    h = _split(k)
    X1, X2 = X[:, :h], X[:, h:]
    _syrk_t_add_(X1, out[:h, :h])
    _syrk_t_add_(X2, out[h:, h:])
    T = torch.mm(X2.t(), X1)
    out[h:, :h].add_(T)
    out[:h, h:].add_(T.t())


def _tril_inverse_into(L, out):
    """out <- L^-1 for lower-triangular L; `out` arrives zero-filled and its strict upper part stays zero."""
    n = L.shape[0]
    if n <= _INV_LEAF:
        out.copy_(torch.linalg.solve_triangular(L, torch.eye(n, dtype=L.dtype, device=L.device), upper=False))
        return
    h = _split(n)
    _tril_inverse_into(L[:h, :h], out[:h, :h])
    _tril_inverse_into(L[h:, h:], out[h:, h:])
    # [[A, 0], [B, C]]^-1 = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]
    B = out[h:, :h]
    B.copy_(L[h:, :h])
    _mm_right_lower_(B, out[:h, :h])
    _mm_left_lower_(out[h:, h:], B)
    B.neg_()


def _gram_of_tril_into(Li, out):
    """out <- Li^T Li for lower-triangular Li (stored with an explicit zero upper part)."""
    n = Li.shape[0]
    if n <= _INV_LEAF:
        torch.mm(Li.t(), Li, out=out)
        return
    h = _split(n)
    A, X, C = Li[:h, :h], Li[h:, :h], Li[h:, h:]
    _gram_of_tril_into(A, out[:h, :h])
    _syrk_t_add_(X, out[:h, :h])
    _gram_of_tril_into(C, out[h:, h:])
    _mm_right_lower_into(X.t(), C, out[:h, h:])
    out[h:, :h].copy_(out[:h, h:].t())


def spd_inverse_from_cholesky(L):
    """K^-1 = L^-T L^-1 given the lower factor of K = L L^T (only the lower triangle of L is read)."""
    Li = torch.zeros_like(L)
    _tril_inverse_into(L, Li)
    out = torch.empty_like(L)
    _gram_of_tril_into(Li, out)
    return out


# ------------------------------------------------------------------------------------------------
# Factor precompute without dense intermediates (replaces the recipe of gpmdm.py:1284-1305: upper Cholesky,
# torch.inverse(U), U^-1 U^-T -- five live N x N arrays in a direct transcription).  Here:
#     K  --cholesky_ex-->  L (lower; K is released)  --in place-->  L^-1  --GEMM per column panel-->  Q panels
# with K^-1 = L^-T L^-1 never stored densely: column panel J of the quadratic-form matrix the predict kernels stream
# (include/gpmdm_b200.h: gpmdm_gp_block) is one GEMM   L^-1[c0:, r0:]^T  L^-1[c0:, c0:c0+256]   written straight into
# the padded panel.  Peak memory: 2 N^2 doubles (K and L during the factorisation), 1.5 N^2 afterwards.
# Pure torch (cuSOLVER potrf + cuBLAS DGEMM on the device); device agnostic, so the CPU suite checks the algebra.
# ------------------------------------------------------------------------------------------------
_TRINV_LEAF = 1024
_PANEL_ROW_BLOCK = 2048  # row blocks of a column panel's GEMM below its diagonal block
PANEL_LD = 260  # GPMDM_PANEL_LD


def tril_inverse_inplace(L):
    """L <- L^-1 for a lower-triangular L whose strict upper part is zero (stays zero).
    [[A, 0], [B, C]]^-1 = [[A^-1, 0], [-C^-1 B A^-1, C^-1]], recursively; the two products recurse on their triangular
    operand in place (temporaries: one leaf-wide strip), GEMM-shaped throughout: DGEMM runs at ~36 TF/s on B200, TRSM /
    cuSOLVER's trtri well below."""
    n = L.shape[0]
    if n <= _TRINV_LEAF:
        L.copy_(torch.linalg.solve_triangular(L, torch.eye(n, dtype=L.dtype, device=L.device), upper=False))
        return L
    h = _split(n)
    A, B, C = L[:h, :h], L[h:, :h], L[h:, h:]
    tril_inverse_inplace(A)
    tril_inverse_inplace(C)
    _mm_right_lower_(B, A)
    _mm_left_lower_(C, B)
    B.neg_()
    return L


def panel_row_offset(t: int, n_pad: int, tri: bool) -> int:
    """First row of column panel t in the packed quadratic-form array (csrc/common.cuh: panel_row_offset)."""
    return t * n_pad - (TILE_N // 2) * t * (t - 1) if tri else t * n_pad


def quadform_panel_elems(n_pad: int, tri: bool) -> int:
    return panel_row_offset(n_pad // TILE_N, n_pad, tri) * PANEL_LD


def quadform_panels_from_tril_inverse(Linv, n_pad: int, tri: bool, out=None):
    """Column panels of Q (k^T K^-1 k == k^T Q k; tri: Q = 2 K^-1 strictly below the diagonal, K^-1 on it, 0 above;
    dense: Q = K^-1) for K^-1 = Linv^T Linv, one GEMM per 256-column panel, written in place into the layout the
    predict kernels stream.  Linv [n, n] lower triangular with an explicit zero upper part."""
    n = Linv.shape[0]
    if out is None:
        out = torch.empty(quadform_panel_elems(n_pad, tri), dtype=Linv.dtype, device=Linv.device)
    out.zero_()
    for J in range(n_pad // TILE_N):
        c0 = J * TILE_N
        if c0 >= n:
            break
        w = min(TILE_N, n - c0)
        r0 = c0 if tri else 0
        off = panel_row_offset(J, n_pad, tri) * PANEL_LD
        dst = out[off:off + (n_pad - r0) * PANEL_LD].view(n_pad - r0, PANEL_LD)[:n - r0, :w]
        # K^-1[i, j] = sum_{k >= max(i, j)} Linv[k, i] Linv[k, j];  j >= c0 here, so k runs over rows c0.. only -- and for
        # the rows i >= r of a row block over k >= r only (below the panel's diagonal block Linv[k, i] = 0 for k < i)
        if tri:
            r = c0
            while r < n:
                r1 = min(n, r + (w if r == c0 else _PANEL_ROW_BLOCK))
                torch.mm(Linv[r:, r:r1].t(), Linv[r:, c0:c0 + w], out=dst[r - r0:r1 - r0])
                r = r1
        else:
            torch.mm(Linv[c0:, r0:].t(), Linv[c0:, c0:c0 + w], out=dst)
        if tri:
            dst[w:].mul_(2.0)
            diag = dst[:w]
            diag.copy_(2.0 * torch.tril(diag, -1) + torch.diag(torch.diagonal(diag)))
    return out


def dense_from_quadform_panels(panels, n: int, n_pad: int, tri: bool):
    """The dense symmetric K^-1 [n, n] back from the packed panels (API parity: `Ky_inv`, `Kx_inv_class[c]`)."""
    out = torch.zeros(n, n, dtype=panels.dtype, device=panels.device)
    for J in range(n_pad // TILE_N):
        c0 = J * TILE_N
        if c0 >= n:
            break
        w = min(TILE_N, n - c0)
        r0 = c0 if tri else 0
        off = panel_row_offset(J, n_pad, tri) * PANEL_LD
        out[r0:, c0:c0 + w] = panels[off:off + (n_pad - r0) * PANEL_LD].view(n_pad - r0, PANEL_LD)[:n - r0, :w]
    if tri:  # halving is exact
        low = torch.tril(out, -1) * 0.5
        out = low + low.t() + torch.diag(torch.diagonal(out))
    return out


class _LogdetTrace(torch.autograd.Function):
    """(log det K, tr(K^-1 T T^T)) of a symmetric positive-definite K that is block diagonal over `offsets`
    (one block when None) -- the two matrix scalars of the NLL (gpmdm.py:576-589, :617-628).

    forward: per block, one Cholesky factor (torch.linalg / cuSOLVER), A = K^-1 T by two triangular solves.
    backward: closed form instead of autograd through the factorisation --
        d logdet / dK = K^-1,   d tr / dK = -A A^T,   d tr / dT = 2 A,
    so a training step costs one potrf + one potri per block (N^3 flops) and the gradient matrix is written straight
    into the buffer of K^-1.  The reference factors the masked dynamics matrix densely (Nx^3/3); the class blocks are
    factored independently here (sum N_c^3/3) with identical results: the Cholesky factor of a block-diagonal matrix
    is block diagonal."""

    @staticmethod
    def forward(ctx, K, T, offsets):
        n = K.shape[0]
        bounds = [(0, n)] if offsets is None else [(a, b) for a, b in zip(offsets[:-1], offsets[1:]) if b > a]
        logdet = K.new_zeros(())
        tr = K.new_zeros(())
        saved = []
        for a, b in bounds:
            Kb = K if (a, b) == (0, n) else K[a:b, a:b]
            U, _info = torch.linalg.cholesky_ex(Kb, upper=False)  # lower: cuSOLVER's faster side (88 vs 121 ms at N = 20 k)
            Tb = T[a:b]
            A = torch.cholesky_solve(Tb, U, upper=False)
            logdet = logdet + 2 * torch.sum(torch.log(torch.diagonal(U)))
            tr = tr + torch.sum(Tb * A)
            saved += [U, A]
        ctx.bounds, ctx.n = bounds, n
        ctx.save_for_backward(*saved)
        return logdet, tr

    @staticmethod
    def backward(ctx, g_logdet, g_tr):
        saved, n, bounds = ctx.saved_tensors, ctx.n, ctx.bounds
        single = len(bounds) == 1
        G = None if single else saved[0].new_zeros(n, n)
        gT = []
        for k, (a, b) in enumerate(bounds):
            U, A = saved[2 * k], saved[2 * k + 1]
            Gb = spd_inverse_from_cholesky(U)
            Gb.mul_(g_logdet).addmm_(A, A.t(), alpha=-1.0 * g_tr)
            gT.append(2 * g_tr * A)
            if single:
                G = Gb
            else:
                G[a:b, a:b] = Gb
        return G, (gT[0] if single else torch.cat(gT, 0)), None


class GPMDM(torch.nn.Module):
    """Gaussian Process Multi-Dynamical Model (reference gpmdm.py:18), B200-native."""

    def __init__(self, D, d, n_classes, dyn_target, dyn_back_step,
                 y_lambdas_init, y_lengthscales_init, y_sigma_n_init,
                 x_lambdas_init, x_lengthscales_init, x_sigma_n_init, x_lin_coeff_init,
                 flg_train_y_lambdas=True, flg_train_y_lengthscales=True, flg_train_y_sigma_n=True,
                 flg_train_x_lambdas=True, flg_train_x_lengthscales=True,
                 flg_train_x_sigma_n=True, flg_train_x_lin_coeff=True,
                 sigma_n_num_Y=0., sigma_n_num_X=0.,
                 dtype=torch.float64, device=None):
        super().__init__()
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("dtype must be torch.float64 or torch.float32")
        # `dtype` is the PUBLIC dtype, as in the reference: parameters, latents and returned predictions.  All kernel
        # arithmetic and every factor is fp64 whatever it is (the fp32 reference inverts K in fp32; here fp32 parameters
        # are up-cast first).  A float32 model makes the particle filter default to the tf32 tensor-core variant of the
        # observation GP (precision="tf32": north_star's 1e-4 variant).
        self.dtype = dtype
        self.device = torch.device(device) if device is not None else _default_device()
        if self.device.type != "cuda":
            raise RuntimeError("gpmdm_b200 needs a CUDA device; there is no CPU path")
        self.D, self.d, self.n_classes = D, d, n_classes
        self.dyn_target, self.dyn_back_step = dyn_target, dyn_back_step

        def par(init, flag):
            return torch.nn.Parameter(torch.log(to_tensor(init, F64, self.device)).to(self.dtype), requires_grad=flag)

        self.y_log_lengthscales = par(y_lengthscales_init, flg_train_y_lengthscales)
        self.y_log_lambdas = par(y_lambdas_init, flg_train_y_lambdas)
        self.y_log_sigma_n = par(y_sigma_n_init, flg_train_y_sigma_n)
        self.x_log_lengthscales = par(x_lengthscales_init, flg_train_x_lengthscales)
        self.x_log_lambdas = par(x_lambdas_init, flg_train_x_lambdas)
        self.x_log_sigma_n = par(x_sigma_n_init, flg_train_x_sigma_n)
        self.x_log_lin_coeff = par(x_lin_coeff_init, flg_train_x_lin_coeff)
        self.sigma_n_num_Y = sigma_n_num_Y
        self.sigma_n_num_X = sigma_n_num_X
        self.class_aware_observations_list = [[] for _ in range(self.n_classes)]
        self.meanY = 0
        self._Y_dev = None
        self._factors_version = 0

    def _d(self, name):
        """A parameter as the kernels see it: detached, fp64 (a no-op view for float64 models)."""
        return getattr(self, name).detach().to(F64)

    # ---- modes (gpmdm.py:239-279) ------------------------------------------------------------------
    def set_evaluation_mode(self):
        self.flg_trainable_list = []
        for p in self.parameters():
            p.requires_grad = False

    def set_training_mode(self, model='all'):
        y = (self.y_log_lengthscales, self.y_log_lambdas, self.y_log_sigma_n)
        x = (self.x_log_lengthscales, self.x_log_lambdas, self.x_log_sigma_n, self.x_log_lin_coeff)
        if model == 'all':
            for p in self.parameters():
                p.requires_grad = True
        elif model == 'latent':
            for p in y:
                p.requires_grad = True
            for p in x:
                p.requires_grad = False
        elif model == 'dynamics':
            for p in y:
                p.requires_grad = False
            for p in x:
                p.requires_grad = True
        else:
            raise ValueError('model must be \'all\', \'latent\' or \'dynamics\'')

    # ---- data (gpmdm.py:281-309, 779-815) ----------------------------------------------------------
    def add_data(self, Y, class_index: int):
        if Y.shape[1] != self.D:
            raise ValueError('Y must be a N x D matrix collecting observation data!')
        self.class_aware_observations_list[class_index].append(Y)
        self._Y_dev = None

    @property
    def observations_list(self):
        return [seq for class_seqs in self.class_aware_observations_list for seq in class_seqs]

    def get_Y(self) -> np.ndarray:
        observation = np.concatenate(self.observations_list, 0)
        self.meanY = 0
        return observation - self.meanY

    def get_Y_for_class(self, class_index: int) -> np.ndarray:
        observation = np.concatenate(self.class_aware_observations_list[class_index], 0)
        self.meanY = 0
        return observation - self.meanY

    def _Y_device(self):
        if self._Y_dev is None:
            self._Y_dev = torch.tensor(self.get_Y(), dtype=F64, device=self.device)
        return self._Y_dev

    # ---- class structure as offsets (replaces get_M / get_M_for_class, gpmdm.py:311-378) -------------
    def class_frame_offsets(self) -> List[int]:
        out = [0]
        for cls in self.class_aware_observations_list:
            out.append(out[-1] + sum(len(s) for s in cls))
        return out

    def class_pair_offsets(self, back_step: Optional[int] = None) -> List[int]:
        b = self.dyn_back_step if back_step is None else back_step
        out = [0]
        for cls in self.class_aware_observations_list:
            out.append(out[-1] + sum(len(s) - b for s in cls))
        return out

    def _pair_offsets_dev(self):
        return torch.tensor(self.class_pair_offsets(), dtype=torch.int64, device=self.device)

    def get_M(self):
        """Dense class mask (gpmdm.py:311-340).  Provided for API parity / small N; unused internally."""
        offs = self.class_pair_offsets()
        M = torch.zeros(offs[-1], offs[-1], dtype=F64, device=self.device)
        for a, b in zip(offs[:-1], offs[1:]):
            M[a:b, a:b] = 1
        return M

    def get_M_for_class(self, class_index: int):
        offs = self.class_pair_offsets()
        M = torch.zeros(offs[-1], offs[-1], dtype=F64, device=self.device)
        a, b = offs[class_index], offs[class_index + 1]
        M[a:b, a:b] = 1
        return M

    # ---- kernels (gpmdm.py:381-548) ------------------------------------------------------------------
    def get_y_kernel(self, X1, X2, flg_noise=True):
        return self.get_rbf_kernel(X1, X2, self.y_log_lengthscales, self.y_log_sigma_n, self.sigma_n_num_Y, flg_noise)

    def get_x_kernel(self, X1, X2, flg_noise=True):
        if X1 is X2:
            return _KernelBuild.apply(X1, self.x_log_lengthscales, self.x_log_sigma_n, self.x_log_lin_coeff,
                                      self.sigma_n_num_X, None, flg_noise)
        return self.get_rbf_kernel(X1, X2, self.x_log_lengthscales, self.x_log_sigma_n, self.sigma_n_num_X, flg_noise) \
            + self.get_lin_kernel(X1, X2, self.x_log_lin_coeff)

    def get_rbf_kernel(self, X1, X2, log_lengthscales_par, log_sigma_n_par, sigma_n_num=0, flg_noise=True):
        if X1 is X2:  # training-side symmetric build: CUDA kernel with hand-written backward
            return _KernelBuild.apply(X1, log_lengthscales_par, log_sigma_n_par, None, sigma_n_num, None, flg_noise)
        # rectangular cross-kernels are not on any hot path here (the filter never materialises them);
        # kept as plain device expressions for API parity with gpmdm.py:474-481
        K = torch.exp(-self.get_weighted_distances(X1, X2, log_lengthscales_par))
        if flg_noise:
            N = X1.shape[0]
            eye = torch.eye(N, dtype=F64, device=self.device)
            K = K + torch.exp(log_sigma_n_par) ** 2 * eye + sigma_n_num ** 2 * eye
        return K

    def get_weighted_distances(self, X1, X2, log_lengthscales_par):
        lengthscales = torch.exp(log_lengthscales_par)
        A = X1 / lengthscales
        A2 = torch.sum(A.mul(A), dim=1, keepdim=True)
        B = X2 / lengthscales
        B2 = torch.sum(B.mul(B), dim=1, keepdim=True)
        return A2 + B2.transpose(0, 1) - 2 * torch.matmul(A, B.transpose(0, 1))

    def get_lin_kernel(self, X1, X2, log_lin_coeff_par):
        Sigma = torch.diag(torch.exp(log_lin_coeff_par) ** 2)
        X1 = torch.cat([X1, torch.ones(X1.shape[0], 1, dtype=F64, device=self.device)], 1)
        X2 = torch.cat([X2, torch.ones(X2.shape[0], 1, dtype=F64, device=self.device)], 1)
        return torch.matmul(X1, torch.matmul(Sigma, X2.transpose(0, 1)))

    def get_masked_x_kernel(self, Xin):
        """`get_x_kernel(Xin, Xin) * self.M` (gpmdm.py:616, 1292) without the dense mask."""
        return _KernelBuild.apply(Xin, self.x_log_lengthscales, self.x_log_sigma_n, self.x_log_lin_coeff,
                                  self.sigma_n_num_X, self._pair_offsets_dev(), True)

    def get_x_diag_kernel(self, X, flg_noise=False):
        c2 = torch.exp(self.x_log_lin_coeff) ** 2
        Xa = torch.cat([X, torch.ones(X.shape[0], 1, dtype=F64, device=self.device)], 1)
        out = torch.ones(X.shape[0], dtype=F64, device=self.device) + torch.sum((Xa * c2) * Xa, dim=1)
        if flg_noise:
            out = out + torch.exp(self.x_log_sigma_n) ** 2 + self.sigma_n_num_X ** 2
        return out

    def get_y_diag_kernel(self, X, flg_noise=False):
        out = torch.ones(X.shape[0], dtype=F64, device=self.device)
        if flg_noise:
            out = out + torch.exp(self.y_log_sigma_n) ** 2 + self.sigma_n_num_Y ** 2
        return out

    # ---- NLL (gpmdm.py:550-628, 721-760) --------------------------------------------------------------
    @staticmethod
    def _logdet_and_trace(K, T, offsets=None):
        """log det K and tr(K^-1 T T^T) (closed-form backward, block-wise factorisation: `_LogdetTrace`)."""
        return _LogdetTrace.apply(K, T.to(K.dtype), offsets)

    def get_y_neg_log_likelihood(self, Y, X, N):
        K_y = self.get_y_kernel(X, X)
        logdet, tr = self._logdet_and_trace(K_y, Y * torch.exp(self.y_log_lambdas))
        log_det_W = 2 * torch.sum(self.y_log_lambdas)
        return self.D / 2 * logdet + 1 / 2 * tr - N * log_det_W

    def get_x_neg_log_likelihood(self, Xout, Xin):
        K_x = self.get_masked_x_kernel(Xin)
        logdet, tr = self._logdet_and_trace(K_x, Xout * torch.exp(self.x_log_lambdas), self.class_pair_offsets())
        log_det_W = 2 * torch.sum(self.x_log_lambdas)
        return self.d / 2 * logdet + 1 / 2 * tr - Xin.shape[0] * log_det_W

    def _pair_index(self, target=None, back_step=None):
        """Row indices realising `get_Xin_Xout_matrices` (gpmdm.py:630-718) as gathers."""
        target = self.dyn_target if target is None else target
        b = self.dyn_back_step if back_step is None else back_step
        if target not in ('full', 'delta') or b not in (1, 2):
            raise ValueError('target must be either \'full\' or \'delta\' \n back_step must be either 1 or 2')
        idx_in, idx_prev, idx_out, starts, s = [], [], [], [], 0
        for seq in self.observations_list:
            L = seq.shape[0]
            starts.append(s)
            r = np.arange(s + b - 1, s + L - 1)
            idx_in.append(r)
            idx_prev.append(r - 1)
            idx_out.append(r + 1)
            s += L
        cat = lambda v: torch.as_tensor(np.concatenate(v), device=self.device)
        return cat(idx_in), cat(idx_prev), cat(idx_out), starts, target, b

    def get_Xin_Xout_matrices(self, X=None, target=None, back_step=None):
        if X is None:
            X = self.X
        i_in, i_prev, i_out, starts, target, b = self._pair_index(target, back_step)
        Xin = X[i_in] if b == 1 else torch.cat((X[i_in], X[i_prev]), 1)
        Xout = X[i_out] if target == 'full' else X[i_out] - X[i_in]
        return Xin, Xout, starts

    def gpdm_loss(self, Y, N, M=None, balance=1):
        Xin, Xout, _ = self.get_Xin_Xout_matrices()
        lossY = self.get_y_neg_log_likelihood(Y, self.X, N)
        lossX = self.get_x_neg_log_likelihood(Xout, Xin)
        return lossY + balance * lossX

    # ---- init / train (gpmdm.py:762-885) ---------------------------------------------------------------
    def init_X(self):
        from sklearn.decomposition import PCA

        Y = self.get_Y()
        X0 = PCA(n_components=self.d).fit_transform(Y)
        self._precompute_class_matrices()
        self.X = torch.nn.Parameter(torch.tensor(X0, dtype=self.dtype, device=self.device), requires_grad=True)
        self._precompute_kernel_inverses()

    def train_adam(self, num_opt_steps, num_print_steps=0, lr=0.01, balance=1):
        if num_print_steps != 0:
            print('\n### Model Training (Adam) ###')
        Y = self._Y_device()
        N = Y.shape[0]
        self.set_training_mode('all')
        optimizer = torch.optim.Adam(self.parameters(), lr=lr)
        t_start = time.time()
        losses = []
        for epoch in range(num_opt_steps):
            optimizer.zero_grad()
            # NB the reference passes `balance` in the position of the unused `M` argument
            # (gpmdm.py:866 vs :721-726), so its balance is always 1; reproduced.
            loss = self.gpdm_loss(Y, N, balance)
            loss.backward()
            if torch.isnan(loss):
                print('Loss is nan')
                break
            optimizer.step()
            losses.append(loss.item())
            if (num_print_steps != 0) and epoch % num_print_steps == 0:
                print('\nGPDM Opt. EPOCH:', epoch)
                print('Running loss:', "{:.4e}".format(loss.item()))
                t_stop = time.time()
                print('Update time:', t_stop - t_start)
                t_start = t_stop
        self._precompute_kernel_inverses()
        return losses

    def get_latent_sequences(self):
        X_np = self.X.clone().detach().cpu().numpy()
        out, s = [], 0
        for seq in self.observations_list:
            out.append(X_np[s:s + seq.shape[0], :])
            s += seq.shape[0]
        return out

    def get_X_for_class(self, class_index: int):
        offs = self.class_frame_offsets()
        return self.X[offs[class_index]:offs[class_index + 1], :]

    # ---- precompute (gpmdm.py:1275-1305) -----------------------------------------------------------------
    def _precompute_class_matrices(self):
        # the reference builds C+1 dense Nx x Nx masks here; offsets carry the same information
        self.class_offsets = self.class_pair_offsets()

    @staticmethod
    def _inverse_via_upper_cholesky(K):
        """U = chol_upper(K); K^-1 = U^-1 U^-T (gpmdm.py:1287-1289) as a dense matrix -- only for the class-agnostic
        `Kx_inv` (not on the filter path); the filter's factors never take this route (`_factor_block`)."""
        U, _info = torch.linalg.cholesky_ex(K, upper=True)
        eye = torch.eye(K.shape[0], dtype=K.dtype, device=K.device)
        U_inv = torch.linalg.solve_triangular(U, eye, upper=True)
        return torch.matmul(U_inv, U_inv.t())

    # Factor precisions built by `_precompute_kernel_inverses`: "fp64" = the quadratic-form panels of the exact path,
    # "tf32" / "f16x2" = the whitening-factor tiles of the tensor-core variants.  Whatever is missing is built on first use; a tf32-only
    # deployment at N = 50 k sets ("tf32",) to never hold the 10 GB of fp64 panels.
    default_factor_precisions = ("fp64",)

    def _factor_block(self, make_K, targets, want=("fp64",), tri=True):
        """Factors of one GP block from its kernel matrix K = make_K() (built here so that this frame holds the only
        reference and K is released as soon as it is factored).  K = L L^T (the reference's U is L^T), L^-1 in place, then
            alpha  = K^-1 targets = L^-T (L^-1 targets)                         (gpmdm.py:957, 1064)
            panels = column panels of Q, K^-1 = L^-T L^-1 never stored densely  (gpmdm.py:1289, 1305)
            wtiles = W = U^-T = L^-1 as tf32 hi/lo tensor-core tiles            (tf32 variant)
        Peak memory 2 N^2 doubles (K + L inside cholesky_ex)."""
        lib = _cabi.lib()
        K = make_K()
        n = K.shape[0]
        n_pad = _round_up(n, TILE_N)
        L, _info = torch.linalg.cholesky_ex(K, upper=False)  # gpmdm.py:1287: info is ignored by the reference as well
        del K
        if not L.is_contiguous():  # torch.linalg hands back column-major storage; the packers read row-major
            L = L.contiguous()
        Linv = tril_inverse_inplace(L)
        del L
        blk = dict(n=n, n_pad=n_pad, dense=None, panels={}, wtiles=None, wtiles_f16=None,
                   A=torch.mm(Linv.t(), torch.mm(Linv, targets)).contiguous())
        if "fp64" in want:
            blk["panels"][bool(tri)] = quadform_panels_from_tril_inverse(Linv, n_pad, bool(tri))
        if "tf32" in want:
            wt = torch.empty(int(lib.gpmdm_tf32_wtiles_bytes(n_pad)) // 4, dtype=torch.float32, device=self.device)
            check(lib.gpmdm_pack_whitened_tf32(ptr(Linv), n, n_pad, ptr(wt), stream()), "gpmdm_pack_whitened_tf32")
            blk["wtiles"] = wt
        if "f16x2" in want:
            wmax = float(Linv.abs().max())
            if not wmax < 3.0e4:  # fp16 overflows at 65504: |W| <= 1/sigma_n, so this only happens for noise std < ~3e-5
                raise ValueError("precision 'f16x2' needs |W| = |L^-1| < 3e4 (max is %.3g): use 'tf32' or 'fp64'" % wmax)
            wh = torch.empty(int(lib.gpmdm_f16_wtiles_bytes(n_pad)) // 2, dtype=torch.float16, device=self.device)
            check(lib.gpmdm_pack_whitened_f16x2(ptr(Linv), n, n_pad, ptr(wh), stream()), "gpmdm_pack_whitened_f16x2")
            blk["wtiles_f16"] = wh
        return blk

    def _obs_kernel_matrix(self):
        X = self._d("X")
        return self.get_y_kernel(X, X)  # same object twice: the symmetric CUDA build (csrc/train_kernels.cu)

    def _dyn_kernel_matrix(self, c, jitter=1e-6):
        offs = self.class_pair_offsets()
        Xc = self._Xin[offs[c]:offs[c + 1]].contiguous()
        Kc = _KernelBuild.apply(Xc, self.x_log_lengthscales, self.x_log_sigma_n, self.x_log_lin_coeff,
                                self.sigma_n_num_X, None, True)
        if jitter:
            Kc.diagonal().add_(jitter)  # gpmdm.py:1302
        return Kc

    @torch.no_grad()
    def _precompute_kernel_inverses(self):
        """The factors the filter consumes (gpmdm.py:1284-1305) -- per block the alpha matrix and the packed
        quadratic-form panels, built without dense N x N intermediates (`_factor_block`).  `Ky_inv` and
        `Kx_inv_class[c]` (N_c x N_c diagonal blocks) remain available as attributes, materialised on access.  The
        reference's dense `Kx_inv` (:1291-1295) is only used by the class-agnostic `map_x_dynamics`, which the filter
        never calls; it is built lazily by `Kx_inv` below."""
        X = self._d("X")
        Xin, Xout, _ = self.get_Xin_Xout_matrices(X)
        self._Xin, self._Xout = Xin.contiguous(), Xout.contiguous()
        want = self.default_factor_precisions
        self._obs_blk = self._factor_block(self._obs_kernel_matrix, self._Y_device(), want)
        offs = self.class_pair_offsets()
        self._dyn_blks = [self._factor_block(lambda: self._dyn_kernel_matrix(c), self._Xout[offs[c]:offs[c + 1]].contiguous())
                          for c in range(self.n_classes)]
        self._Kx_inv_full = None
        self._factors_version += 1
        self._packed = None
        self._packed_tf32 = None

    def _block_panels(self, blk, tri):
        """Quadratic-form panels of a block in the requested packing: as factored, or packed from a dense inverse
        (injected by `set_inverses`, or materialised from the other packing)."""
        tri = bool(tri)
        if tri not in blk["panels"]:
            lib = _cabi.lib()
            if blk["dense"] is None and not blk["panels"]:  # factored without fp64 panels (tf32-only precompute)
                return None
            Kinv = self._block_dense(blk)
            L = torch.empty(int(lib.gpmdm_quadform_bytes(blk["n_pad"], int(tri))) // 8, dtype=F64, device=self.device)
            check(lib.gpmdm_pack_quadform_f64(ptr(Kinv), blk["n"], blk["n_pad"], int(tri), ptr(L), stream()),
                  "gpmdm_pack_quadform_f64")
            blk["panels"][tri] = L
        return blk["panels"][tri]

    def _block_dense(self, blk):
        if blk["dense"] is not None:
            return blk["dense"]
        tri = True if True in blk["panels"] else False
        return dense_from_quadform_panels(blk["panels"][tri], blk["n"], blk["n_pad"], tri)

    def _ensure_fp64_obs_panels(self, tri):
        blk = self._obs_blk
        if blk["dense"] is None and not blk["panels"]:
            fresh = self._factor_block(self._obs_kernel_matrix, self._Y_device(), ("fp64",), tri)
            blk["panels"] = fresh["panels"]

    @property
    def Ky_inv(self):
        """Dense K_y^-1 [N, N] (gpmdm.py:1289).  Not stored: materialised from the packed panels on access."""
        self._ensure_fp64_obs_panels(True)
        return self._block_dense(self._obs_blk)

    @property
    def Kx_inv_class(self):
        """Per class, the N_c x N_c diagonal block of the reference's `Kx_inv_class[c]` (gpmdm.py:1305)."""
        return [self._block_dense(b) for b in self._dyn_blks]

    @property
    def Kx_inv(self):
        if getattr(self, "_Kx_inv_full", None) is None:
            with torch.no_grad():
                self._Kx_inv_full = self._inverse_via_upper_cholesky(self.get_masked_x_kernel(self._Xin))
        return self._Kx_inv_full

    @torch.no_grad()
    def set_inverses(self, Ky_inv=None, Kx_inv_blocks=None):
        """Inject precomputed inverses (e.g. the reference's own `Ky_inv` and the diagonal blocks of its
        `Kx_inv_class[c]`) -- used by the parity tests to compare kernels on identical factors."""
        def injected(Kinv, targets):
            Kinv = to_tensor(Kinv, F64, self.device).contiguous()
            n = Kinv.shape[0]
            return dict(n=n, n_pad=_round_up(n, TILE_N), dense=Kinv, panels={}, wtiles=None, wtiles_f16=None,
                        A=torch.matmul(Kinv.t(), targets).contiguous())

        if Ky_inv is not None:
            self._obs_blk = injected(Ky_inv, self._Y_device())
        if Kx_inv_blocks is not None:
            offs = self.class_pair_offsets()
            self._dyn_blks = [injected(b, self._Xout[offs[c]:offs[c + 1]]) for c, b in enumerate(Kx_inv_blocks)]
        self._factors_version += 1
        self._packed = None
        self._packed_tf32 = None

    # ---- packing for the fused predict kernels ---------------------------------------------------------------
    def _pack_block(self, Xtrain, log_ls, blk, alpha_ld, lin_c2, tri, with_L=True):
        lib = _cabi.lib()
        n, d = Xtrain.shape
        n_pad = blk["n_pad"]
        cols = [Xtrain / torch.exp(log_ls)]
        if lin_c2 is not None:
            cols.append(Xtrain * lin_c2[:d])
        rec = torch.cat(cols, 1)
        width = (rec.shape[1] + 1) & ~1  # records are padded to an even number of doubles (16-byte loads)
        coords = torch.zeros(n_pad, width, dtype=F64, device=self.device)
        coords[:n, :rec.shape[1]] = rec
        L = self._block_panels(blk, tri) if with_L else None  # column panels (include/gpmdm_b200.h: gpmdm_gp_block)
        A = blk["A"]
        alpha = torch.empty(int(lib.gpmdm_alpha_bytes(n_pad, alpha_ld)) // 8, dtype=F64, device=self.device)
        check(lib.gpmdm_pack_alpha_f64(ptr(A), n, n_pad, A.shape[1], alpha_ld, ptr(alpha), stream()), "gpmdm_pack_alpha_f64")
        return dict(coords=coords, L=L, alpha=alpha, n=n, n_pad=n_pad)

    @torch.no_grad()
    def packed_models(self, tri: bool = True, with_obs_L: bool = True):
        """Device-resident operands of the fused kernels (include/gpmdm_b200.h: gpmdm_gp_model).
        with_obs_L=False leaves out the fp64 quadratic-form matrix of the observation block (8 N^2 bytes): enough for
        gpmdm_pf_loglik_f64 when the tf32 variant supplies the variances."""
        if getattr(self, "_packed", None) is not None and self._packed["tri"] == tri \
                and (self._packed["obs_has_L"] or not with_obs_L):
            return self._packed
        X = self._d("X")
        dev = self.device
        keep = []  # tensors that must outlive the C structs

        def model(blocks, d, dout, alpha_ld, kind, ls, c2, lam):
            table = torch.tensor([[b["coords"].data_ptr(), b["L"].data_ptr() if b["L"] is not None else 0,
                                   b["alpha"].data_ptr(), b["n"], b["n_pad"]]
                                  for b in blocks], dtype=torch.int64, device=dev)
            keep.extend([table, ls, c2, lam, blocks])
            return GpModel(blocks=table.data_ptr(), n_blocks=len(blocks), d=d, dout=dout, alpha_ld=alpha_ld, kind=kind,
                           tri=int(tri), lengthscales=ls.data_ptr(), lin_c2=c2.data_ptr() if c2 is not None else None,
                           lambdas=lam.data_ptr())

        # observation GP: one block over all frames
        ald_y = _round_up(self.D, TILE_N)
        if with_obs_L:
            self._ensure_fp64_obs_panels(tri)
        oblk = self._pack_block(X, self._d("y_log_lengthscales"), self._obs_blk, ald_y, None, tri, with_L=with_obs_L)
        ls_y = torch.exp(self._d("y_log_lengthscales")).contiguous()
        lam2_y = (torch.exp(self._d("y_log_lambdas")) ** 2).contiguous()
        obs = model([oblk], self.d, self.D, ald_y, 0, ls_y, None, lam2_y)
        # dynamics GP: one block per class ('full', back_step 1 -- the only mode the filter supports)
        Xin = self._Xin
        if self.dyn_back_step == 1:
            c2 = (torch.exp(self._d("x_log_lin_coeff")) ** 2).contiguous()
            ls_x = torch.exp(self._d("x_log_lengthscales")).contiguous()
            lam_x = (torch.exp(self._d("x_log_lambdas")) ** -2).contiguous()
            offs = self.class_pair_offsets()
            dblks = [self._pack_block(Xin[offs[c]:offs[c + 1]], self._d("x_log_lengthscales"), self._dyn_blks[c],
                                      TILE_N, c2, tri) for c in range(self.n_classes)]
            dyn = model(dblks, self.d, self.d, TILE_N, 1, ls_x, c2, lam_x)
        else:
            dyn = None
        self._packed = dict(tri=tri, obs=obs, obs_has_L=with_obs_L, dyn=dyn, dyn_blocks=dblks if dyn is not None else None,
                            keep=keep, obs_n_pad=oblk["n_pad"],
                            dyn_max_n_pad=max(b["n_pad"] for b in dblks) if dyn is not None else 0,
                            ll_const_terms=(2.0 * torch.sum(self._d("y_log_lambdas"))).item())
        return self._packed

    @torch.no_grad()
    def packed_model_tf32(self, kind="tf32"):
        """Operands of the tensor-core observation kernels (include/gpmdm_b200.h: gpmdm_gp_model_tf32): the whitening
        factor W = U^-T = L^-1 of K_y = U^T U = L L^T as tensor-core operand tiles -- tf32 hi/lo (kind "tf32", plus alpha_y
        tiles) or fp16 hi / scaled lo (kind "f16x2")."""
        cache = self.__dict__.setdefault("_packed_tc", {})
        if kind in cache and cache[kind]["version"] == self._factors_version:
            return cache[kind]
        lib = _cabi.lib()
        X = self._d("X")
        n, d = X.shape
        n_pad = _round_up(n, TILE_N)
        key = "wtiles" if kind == "tf32" else "wtiles_f16"
        if self._obs_blk.get(key) is None:  # not part of the precompute: factor K_y (again) for the W tiles only
            self._obs_blk[key] = self._factor_block(self._obs_kernel_matrix, self._Y_device(), (kind,))[key]
        wt = self._obs_blk[key]
        at = None
        if kind == "tf32":
            at = torch.empty(int(lib.gpmdm_tf32_atiles_bytes(n_pad)) // 4, dtype=torch.float32, device=self.device)
            check(lib.gpmdm_pack_alpha_tf32(ptr(self._obs_blk["A"]), n, n_pad, self.D, ptr(at), stream()), "gpmdm_pack_alpha_tf32")
        coords = torch.zeros(n_pad, 8, dtype=torch.float32, device=self.device)
        coords[:n, :d] = (X / torch.exp(self._d("y_log_lengthscales")) * 1.2011224087864498).to(torch.float32)  # sqrt(log2 e)
        coords = coords.view(n_pad // 2, 2, 8).transpose(1, 2).contiguous()  # [pair][coordinate][row in pair]
        ls = torch.exp(self._d("y_log_lengthscales")).contiguous()
        lam2 = (torch.exp(self._d("y_log_lambdas")) ** 2).contiguous()
        model = GpModelTf32(coords=coords.data_ptr(), wtiles=wt.data_ptr(), atiles=at.data_ptr() if at is not None else None,
                            n=n, n_pad=n_pad, d=d, dout=self.D, lengthscales=ls.data_ptr(), lambdas=lam2.data_ptr())
        cache[kind] = dict(version=self._factors_version, model=model, keep=(coords, wt, at, ls, lam2),
                           ll_const_terms=(2.0 * torch.sum(self._d("y_log_lambdas"))).item())
        return cache[kind]

    @torch.no_grad()
    def packed_model_tc_dyn(self, kind="tf32"):
        """Operands of the tensor-core dynamics variance (include/gpmdm_b200.h, "dynamics GP variance on the tensor
        cores"), one block per class with K_c = L_c L_c^T the class block incl. the 1e-6 jitter (gpmdm.py:1301-1303):
          table   gpmdm_tc_block per class: W_c = L_c^-1 as tensor-core tiles + the RBF training records
          model   the dynamics gpmdm_gp_model with alpha tiles [alpha_c | G_c], G_c = K_c^-1 [X_in, 1] diag(c^2)
          H       [C, d+1, d+1]: diag(c^2) [X_in, 1]^T G_c"""
        cache = self.__dict__.setdefault("_packed_tc_dyn", {})
        if kind in cache and cache[kind]["version"] == self._factors_version:
            return cache[kind]
        lib = _cabi.lib()
        pk = self.packed_models()
        if pk["dyn"] is None:
            raise ValueError("fused dynamics prediction supports dyn_back_step == 1 only")
        key = "wtiles" if kind == "tf32" else "wtiles_f16"
        offs = self.class_pair_offsets()
        ls = torch.exp(self._d("x_log_lengthscales")).contiguous()
        c2 = (torch.exp(self._d("x_log_lin_coeff")) ** 2).contiguous()
        lam = (torch.exp(self._d("x_log_lambdas")) ** -2).contiguous()
        d = self.d
        rows, mrows, Hs, keep = [], [], [], [ls, c2, lam]
        for c in range(self.n_classes):
            Xc = self._Xin[offs[c]:offs[c + 1]]
            n = Xc.shape[0]
            XS = torch.cat([Xc, torch.ones(n, 1, dtype=F64, device=self.device)], 1) * c2  # [X_in, 1] diag(c^2)
            fb = self._factor_block(lambda: self._dyn_kernel_matrix(c),
                                    torch.cat([self._Xout[offs[c]:offs[c + 1]], XS], 1).contiguous(), (kind,))
            n_pad = fb["n_pad"]
            Hs.append(XS.t() @ fb["A"][:, d:])
            alpha = torch.empty(int(lib.gpmdm_alpha_bytes(n_pad, TILE_N)) // 8, dtype=F64, device=self.device)
            check(lib.gpmdm_pack_alpha_f64(ptr(fb["A"]), n, n_pad, 2 * d + 1, TILE_N, ptr(alpha), stream()),
                  "gpmdm_pack_alpha_f64")
            coords = torch.zeros(n_pad, 8, dtype=torch.float32, device=self.device)
            coords[:n, :d] = (Xc / ls * 1.2011224087864498).to(torch.float32)  # sqrt(log2 e)
            coords = coords.view(n_pad // 2, 2, 8).transpose(1, 2).contiguous()  # [pair][coordinate][row in pair]
            base = pk["dyn_blocks"][c]  # fp64 training records (and the quadratic-form panels, unused here)
            keep += [coords, fb[key], alpha]
            rows.append([coords.data_ptr(), fb[key].data_ptr(), n, n_pad])
            mrows.append([base["coords"].data_ptr(), base["L"].data_ptr(), alpha.data_ptr(), n, n_pad])
        table = torch.tensor(rows, dtype=torch.int64, device=self.device)
        mtable = torch.tensor(mrows, dtype=torch.int64, device=self.device)
        H = torch.stack(Hs).contiguous()
        model = GpModel(blocks=mtable.data_ptr(), n_blocks=self.n_classes, d=d, dout=d, alpha_ld=TILE_N, kind=1, tri=1,
                        lengthscales=ls.data_ptr(), lin_c2=c2.data_ptr(), lambdas=lam.data_ptr())
        cache[kind] = dict(version=self._factors_version, table=table, model=model, H=H, ls=ls, keep=keep + [mtable, pk],
                           mode=0 if kind == "tf32" else 1)
        return cache[kind]

    LOWLAT_MAX_TILES = 110  # below this many 64-particle tiles (of 148 SMs) the column tiles are split over CTAs

    def _use_lowlat(self, P, low_latency):
        return ((P + TILE_P - 1) // TILE_P <= self.LOWLAT_MAX_TILES) if low_latency is None else bool(low_latency)

    def _stream_scratch(self, name, need, dtype):
        """Scratch of the `map_x_*` calls, one buffer per (purpose, CUDA stream): calls issued on different streams
        never share a K* slice, a work-item counter or a partial-sum workspace.  (`GPMDM_PF` owns its own.)"""
        pool = self.__dict__.setdefault("_scratch_pool", {})
        key = (name, torch.cuda.current_stream().cuda_stream)
        buf = pool.get(key)
        if buf is None or buf.numel() < need:
            buf = pool[key] = torch.zeros(need, dtype=dtype, device=self.device)
        return buf

    def _lowlat_workspace(self, P, max_n_pad, dout):
        need = int(_cabi.lib().gpmdm_predict_lowlat_workspace_bytes(P, max_n_pad, dout, 0, self.n_classes)) // 8 + 1
        return self._stream_scratch("lowlat", need, torch.float64)

    # ---- prediction (gpmdm.py:923-963, 1032-1068) ----------------------------------------------------------
    KSTAR_CACHE_MIN_TILES = 4  # column tiles of L from which caching K* pays (a K* entry is reused nq/2 times)

    def _use_kstar_cache(self, n_pad, kstar_cache):
        env = os.environ.get("GPMDM_KSTAR_CACHE")
        if kstar_cache is None and env is not None:
            kstar_cache = env not in ("0", "")
        return (n_pad // TILE_N >= self.KSTAR_CACHE_MIN_TILES) if kstar_cache is None else bool(kstar_cache)

    def _kstar_workspace(self, n_pad):
        need = int(_cabi.lib().gpmdm_pf_observe_kstar_workspace_bytes(n_pad)) // 8
        return self._stream_scratch("kstar", need, F64)

    def _scratch_counter(self):
        return self._stream_scratch("counter", 4, torch.int32)

    @torch.no_grad()
    def map_x_to_y(self, Xstar, flg_noise=False, precision="fp64", low_latency=None, kstar_cache=None):
        """precision: 'fp64' (exact path, DMMA) or 'tf32' (tcgen05 variant, ~1e-4 relative; an addition).
        low_latency: None = automatic (few particles: split the column tiles over the SMs), True / False to force.
        kstar_cache: None = automatic (fused mode, N_pad >= 1024: K* of a particle tile is evaluated once into a per-SM
        scratch); True / False to force.  Bit-identical results either way."""
        lib = _cabi.lib()
        Xs = to_tensor(Xstar, F64, self.device).contiguous()
        P = Xs.shape[0]
        mu = torch.empty(P, self.D, dtype=F64, device=self.device)
        v = torch.empty(P, dtype=F64, device=self.device)
        if precision == "tf32":  # variances on tcgen05 (tf32 x3, whitened); means stay fp64 (DMMA, alpha tile only)
            pk32, pk = self.packed_model_tf32(), self.packed_models(with_obs_L=False)
            check(lib.gpmdm_pf_observe_tf32(ctypes.byref(pk32["model"]), ptr(Xs), P, None, 0.0, None, None, ptr(v),
                                            ptr(self._scratch_counter()), stream()), "gpmdm_pf_observe_tf32")
            check(lib.gpmdm_pf_loglik_f64(ctypes.byref(pk["obs"]), ptr(Xs), P, None, 0.0, ptr(v), None, ptr(mu),
                                          ptr(self._scratch_counter()), stream()), "gpmdm_pf_loglik_f64")
        elif precision == "f16x2":  # variances on tcgen05 with fp16-split operands; means fp64 as above
            pk16, pk = self.packed_model_tf32("f16x2"), self.packed_models(with_obs_L=False)
            check(lib.gpmdm_pf_observe_f16x2(ctypes.byref(pk16["model"]), ptr(Xs), P, ptr(v), ptr(self._scratch_counter()),
                                             stream()), "gpmdm_pf_observe_f16x2")
            check(lib.gpmdm_pf_loglik_f64(ctypes.byref(pk["obs"]), ptr(Xs), P, None, 0.0, ptr(v), None, ptr(mu),
                                          ptr(self._scratch_counter()), stream()), "gpmdm_pf_loglik_f64")
        elif precision == "tf32-pure":  # mean on the tensor cores as well (error ~1e-4..1e-3 of the row scale)
            pk32 = self.packed_model_tf32()
            check(lib.gpmdm_pf_observe_tf32(ctypes.byref(pk32["model"]), ptr(Xs), P, None, 0.0, None, ptr(mu), ptr(v),
                                            ptr(self._scratch_counter()), stream()), "gpmdm_pf_observe_tf32")
        elif precision == "fp64":
            pk = self.packed_models()
            if P > 0 and self._use_lowlat(P, low_latency):
                ws = self._lowlat_workspace(P, pk["obs_n_pad"], self.D)
                check(lib.gpmdm_pf_observe_lowlat_f64(ctypes.byref(pk["obs"]), ptr(Xs), P, None, 0.0, None, None, ptr(mu),
                                                      ptr(v), pk["obs_n_pad"], 0, ptr(self._scratch_counter()), ptr(ws),
                                                      stream()), "gpmdm_pf_observe_lowlat_f64")
            elif P > 0 and self._use_kstar_cache(pk["obs_n_pad"], kstar_cache):
                ws = self._kstar_workspace(pk["obs_n_pad"])
                check(lib.gpmdm_pf_observe_cached_f64(ctypes.byref(pk["obs"]), ptr(Xs), P, None, 0.0, None, ptr(mu),
                                                      ptr(v), pk["obs_n_pad"], ptr(self._scratch_counter()), ptr(ws),
                                                      ws.numel() * 8, stream()), "gpmdm_pf_observe_cached_f64")
            else:
                check(lib.gpmdm_pf_observe_f64(ctypes.byref(pk["obs"]), ptr(Xs), P, None, 0.0, None, ptr(mu), ptr(v),
                                               ptr(self._scratch_counter()), stream()), "gpmdm_pf_observe_f64")
        else:
            raise ValueError("precision must be 'fp64', 'tf32' or 'f16x2'")
        if flg_noise:
            v = v + torch.exp(self._d("y_log_sigma_n")) ** 2 + self.sigma_n_num_Y ** 2
        var = v.unsqueeze(1) * (torch.exp(self._d("y_log_lambdas")) ** -2).unsqueeze(0)
        return (mu + torch.tensor(self.meanY, dtype=F64, device=self.device)).to(self.dtype), var.to(self.dtype)

    @torch.no_grad()
    def map_x_dynamics_for_class(self, Xstar, class_index: int, flg_noise=False, low_latency=None, kstar_cache=None,
                                 precision="fp64"):
        """gpmdm.py:1032-1068.  `precision` "tf32" / "f16x2": the variance prior - |W_c k*|^2 on the tensor cores
        (gpmdm_pf_dynvar_tc), the mean in fp64 on the alpha tiles only -- what GPMDM_PF(precision=...) runs per step."""
        if precision not in ("fp64", "tf32", "f16x2"):
            raise ValueError(f"unknown precision {precision!r}")
        lib = _cabi.lib()
        pk = self.packed_models()
        if pk["dyn"] is None:
            raise ValueError("fused dynamics prediction supports dyn_back_step == 1 only")
        Xs = to_tensor(Xstar, F64, self.device).contiguous()
        P = Xs.shape[0]
        mean = torch.empty(P, self.d, dtype=F64, device=self.device)
        var = torch.empty(P, self.d, dtype=F64, device=self.device)
        if P == 0:
            return mean, var
        perm = torch.arange(P, dtype=torch.int32, device=self.device)
        nt = (P + TILE_P - 1) // TILE_P
        t = torch.arange(nt, dtype=torch.int32, device=self.device)
        tiles = torch.stack([torch.full_like(t, class_index), t * TILE_P, torch.clamp(P - t * TILE_P, max=TILE_P),
                             torch.zeros_like(t)], 1).contiguous()
        n_tiles = torch.tensor([nt], dtype=torch.int32, device=self.device)
        if precision != "fp64":
            tc = self.packed_model_tc_dyn(precision)
            nt2 = (P + 127) // 128
            t2 = torch.arange(nt2, dtype=torch.int32, device=self.device)
            tiles2 = torch.stack([torch.full_like(t2, class_index), t2 * 128, torch.clamp(P - t2 * 128, max=128),
                                  torch.zeros_like(t2)], 1).contiguous()
            n_tiles2 = torch.tensor([nt2], dtype=torch.int32, device=self.device)
            v = torch.empty(P, dtype=F64, device=self.device)
            check(lib.gpmdm_pf_dynvar_tc(ptr(tc["table"]), self.n_classes, self.d, tc["mode"], ptr(tc["ls"]), ptr(Xs), ptr(perm),
                                         ptr(tiles2), ptr(n_tiles2), P, ptr(v), stream()), "gpmdm_pf_dynvar_tc")
            check(lib.gpmdm_pf_propagate_meanonly_f64(ctypes.byref(tc["model"]), ptr(tc["H"]), ptr(Xs), ptr(perm), ptr(tiles),
                                                      ptr(n_tiles), P, None, ptr(v), None, ptr(mean), ptr(var),
                                                      ptr(self._scratch_counter()), stream()),
                  "gpmdm_pf_propagate_meanonly_f64")
        elif self._use_lowlat(P, low_latency):
            ws = self._lowlat_workspace(P, pk["dyn_max_n_pad"], self.d)
            check(lib.gpmdm_pf_propagate_lowlat_f64(ctypes.byref(pk["dyn"]), ptr(Xs), ptr(perm), ptr(tiles), ptr(n_tiles),
                                                    P, None, None, ptr(mean), ptr(var), pk["dyn_max_n_pad"], 0,
                                                    ptr(self._scratch_counter()), ptr(ws), stream()),
                  "gpmdm_pf_propagate_lowlat_f64")
        elif self._use_kstar_cache(pk["dyn_max_n_pad"], kstar_cache):
            ws = self._kstar_workspace(pk["dyn_max_n_pad"])
            check(lib.gpmdm_pf_propagate_cached_f64(ctypes.byref(pk["dyn"]), ptr(Xs), ptr(perm), ptr(tiles), ptr(n_tiles), P,
                                                    None, None, ptr(mean), ptr(var), pk["dyn_max_n_pad"],
                                                    ptr(self._scratch_counter()), ptr(ws), ws.numel() * 8, stream()),
                  "gpmdm_pf_propagate_cached_f64")
        else:
            check(lib.gpmdm_pf_propagate_f64(ctypes.byref(pk["dyn"]), ptr(Xs), ptr(perm), ptr(tiles), ptr(n_tiles), P,
                                             None, None, ptr(mean), ptr(var), ptr(self._scratch_counter()), stream()),
                  "gpmdm_pf_propagate_f64")
        if flg_noise:
            var = var + (torch.exp(self._d("x_log_sigma_n")) ** 2 + self.sigma_n_num_X ** 2) \
                * (torch.exp(self._d("x_log_lambdas")) ** -2).unsqueeze(0)
        return mean.to(self.dtype), var.to(self.dtype)

    # ---- class-agnostic dynamics map and the notebook helpers (gpmdm.py:993-1030, 1103-1273) -----------------------
    # Not on the filter path; kept so that the reference's notebooks run unchanged.  The masked
    # K_x o M is block diagonal, so k*^T (K_x o M)^-1 (.) is the SUM over classes of the per-class fused prediction
    # with the un-jittered block inverses -- the same kernel as the filter, no dense Nx x Nx product.
    @torch.no_grad()
    def _agnostic_dyn_model(self):
        if getattr(self, "_packed_agn", None) is not None and self._packed_agn["version"] == self._factors_version:
            return self._packed_agn
        offs = self.class_pair_offsets()
        c2 = (torch.exp(self._d("x_log_lin_coeff")) ** 2).contiguous()
        ls_x = torch.exp(self._d("x_log_lengthscales")).contiguous()
        lam_x = (torch.exp(self._d("x_log_lambdas")) ** -2).contiguous()
        blocks = []
        for c in range(self.n_classes):
            Xc = self._Xin[offs[c]:offs[c + 1]].contiguous()
            # block of K_x o M: no 1e-6 jitter (gpmdm.py:1292)
            blk = self._factor_block(lambda: self._dyn_kernel_matrix(c, jitter=0.0), self._Xout[offs[c]:offs[c + 1]].contiguous())
            blocks.append(self._pack_block(Xc, self._d("x_log_lengthscales"), blk, TILE_N, c2, True))
        table = torch.tensor([[b["coords"].data_ptr(), b["L"].data_ptr(), b["alpha"].data_ptr(), b["n"], b["n_pad"]]
                              for b in blocks], dtype=torch.int64, device=self.device)
        model = GpModel(blocks=table.data_ptr(), n_blocks=len(blocks), d=self.d, dout=self.d, alpha_ld=TILE_N, kind=1,
                        tri=1, lengthscales=ls_x.data_ptr(), lin_c2=c2.data_ptr(), lambdas=lam_x.data_ptr())
        self._packed_agn = dict(version=self._factors_version, model=model, keep=(blocks, table, c2, ls_x, lam_x))
        return self._packed_agn

    @torch.no_grad()
    def map_x_dynamics(self, Xstar, flg_noise=False):
        """gpmdm.py:993-1030: GP prediction of the dynamics map with the class-masked (block-diagonal) K_x."""
        if self.dyn_back_step != 1:
            raise ValueError("fused dynamics prediction supports dyn_back_step == 1 only")
        lib = _cabi.lib()
        pk = self._agnostic_dyn_model()
        Xs = to_tensor(Xstar, F64, self.device).contiguous()
        P = Xs.shape[0]
        lam = torch.exp(self._d("x_log_lambdas")) ** -2
        prior = self.get_x_diag_kernel(Xs, flg_noise)
        mean = torch.zeros(P, self.d, dtype=F64, device=self.device)
        q = torch.zeros(P, dtype=F64, device=self.device)
        if P == 0:
            return mean, prior.unsqueeze(1) * lam.unsqueeze(0)
        perm = torch.arange(P, dtype=torch.int32, device=self.device)
        nt = (P + TILE_P - 1) // TILE_P
        t = torch.arange(nt, dtype=torch.int32, device=self.device)
        n_tiles = torch.tensor([nt], dtype=torch.int32, device=self.device)
        m_c = torch.empty(P, self.d, dtype=F64, device=self.device)
        v_c = torch.empty(P, self.d, dtype=F64, device=self.device)
        prior0 = self.get_x_diag_kernel(Xs, False)
        for c in range(self.n_classes):
            tiles = torch.stack([torch.full_like(t, c), t * TILE_P, torch.clamp(P - t * TILE_P, max=TILE_P),
                                 torch.zeros_like(t)], 1).contiguous()
            check(lib.gpmdm_pf_propagate_f64(ctypes.byref(pk["model"]), ptr(Xs), ptr(perm), ptr(tiles), ptr(n_tiles), P,
                                             None, None, ptr(m_c), ptr(v_c), ptr(self._scratch_counter()), stream()),
                  "gpmdm_pf_propagate_f64")
            mean += m_c
            q += prior0 - v_c[:, 0] / lam[0]  # the kernel returns (prior - q_c) * lambda^-2
        return mean, (prior - q).unsqueeze(1) * lam.unsqueeze(0)

    def get_next_x(self, gp_mean_out, gp_out_var, Xold, flg_sample=False):
        """gpmdm.py:1103-1145."""
        from torch.distributions.normal import Normal

        distribution = Normal(gp_mean_out, torch.sqrt(gp_out_var))
        step = distribution.rsample() if flg_sample else gp_mean_out
        if self.dyn_target == 'full':
            return step
        if self.dyn_target == 'delta':
            return Xold + step

    def get_dynamics_map_performance_for_class(self, class_index, flg_noise=False):
        """gpmdm.py:1147-1194 (including its floor-division `//` in the NMSE, reproduced as is)."""
        with torch.no_grad():
            Xin, Xout, _ = self.get_Xin_Xout_matrices()
            mean, var = self.map_x_dynamics_for_class(Xin, class_index, flg_noise=flg_noise)
            mean, var = mean.cpu().numpy(), var.cpu().numpy()
            Xout, Xin = Xout.detach().cpu().numpy(), Xin.detach().cpu().numpy()
            NMSE = np.mean((Xout - mean) ** 2 // var)
        return mean, var, Xout, Xin, NMSE

    def get_latent_map_performance(self, flg_noise=False):
        """gpmdm.py:1196-1237."""
        with torch.no_grad():
            mean, var = self.map_x_to_y(self.X, flg_noise=flg_noise)
            mean, var = mean.cpu().numpy(), var.cpu().numpy()
            Y = self.get_Y() + self.meanY
            return mean, var, Y, np.mean((Y - mean) ** 2 // var)

    def get_latent_map_performance_for_class(self, class_index, flg_noise=False):
        """gpmdm.py:1239-1273."""
        with torch.no_grad():
            mean, var = self.map_x_to_y(self.get_X_for_class(class_index), flg_noise=flg_noise)
            mean, var = mean.cpu().numpy(), var.cpu().numpy()
            Y = self.get_Y_for_class(class_index) + self.meanY
            return mean, var, Y, np.mean((Y - mean) ** 2 // var)

    # ---- save / load (gpmdm.py:1307-1414); same on-disk format ---------------------------------------------------
    def save(self, file_path):
        config_dict = {
            'class_aware_observations_list': self.class_aware_observations_list,
            'dyn_target': self.dyn_target, 'dyn_back_step': self.dyn_back_step,
            'D': self.D, 'd': self.d, 'n_classes': self.n_classes,
            'sigma_n_num_X': self.sigma_n_num_X, 'sigma_n_num_Y': self.sigma_n_num_Y,
            'dtype': str(self.dtype), 'device': str(self.device),
            'y_lengthscales_init': self.y_log_lengthscales.detach().exp().tolist(),
            'y_lambdas_init': self.y_log_lambdas.detach().exp().tolist(),
            'y_sigma_n_init': self.y_log_sigma_n.detach().exp().item(),
            'x_lengthscales_init': self.x_log_lengthscales.detach().exp().tolist(),
            'x_lambdas_init': self.x_log_lambdas.detach().exp().tolist(),
            'x_sigma_n_init': self.x_log_sigma_n.detach().exp().item(),
            'x_lin_coeff_init': self.x_log_lin_coeff.detach().exp().tolist(),
        }
        state = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        torch.save({'state_dict': state, 'config_dict': config_dict}, file_path)
        print(f"Model and hyperparameters saved to {file_path}")

    @classmethod
    def load(cls, file_path, flg_print: bool = False, device=None) -> 'GPMDM':
        # weights_only=False: the config holds numpy arrays (the reference's plain torch.load fails on torch>=2.6)
        save_dict = torch.load(file_path, weights_only=False, map_location="cpu")
        cfg, state_dict = save_dict['config_dict'], save_dict['state_dict']
        dtype = cfg['dtype'][6:] if cfg['dtype'].startswith('torch.') else cfg['dtype']
        model = cls(
            D=cfg['D'], d=cfg['d'], n_classes=cfg['n_classes'], dyn_target=cfg['dyn_target'],
            dyn_back_step=cfg['dyn_back_step'],
            y_lambdas_init=torch.tensor(cfg['y_lambdas_init']), y_lengthscales_init=torch.tensor(cfg['y_lengthscales_init']),
            y_sigma_n_init=cfg['y_sigma_n_init'], x_lambdas_init=torch.tensor(cfg['x_lambdas_init']),
            x_lengthscales_init=torch.tensor(cfg['x_lengthscales_init']), x_sigma_n_init=cfg['x_sigma_n_init'],
            x_lin_coeff_init=torch.tensor(cfg['x_lin_coeff_init']),
            sigma_n_num_X=cfg['sigma_n_num_X'], sigma_n_num_Y=cfg['sigma_n_num_Y'],
            dtype=getattr(torch, dtype), device=device)  # files written on 'cpu' load onto the CUDA device
        model.class_aware_observations_list = cfg['class_aware_observations_list']
        model.init_X()
        model.load_state_dict(state_dict)
        model._precompute_kernel_inverses()
        print("\nModel and hyperparameters correctly loaded")
        if flg_print:
            print("Loaded params:")
            for name, val in model.state_dict().items():
                print(name, "\t", val)
        return model
