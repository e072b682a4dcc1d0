"""Shared test plumbing: golden-fixture loading (tests/golden/*.npz, written by oracle/make_golden.py
from the unmodified reference) and tolerance helpers."""
import glob
import os

import numpy as np
import torch

from oracle import gpmdm_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "c*.npz")))

F64 = torch.float64


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=F64)


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        z = self.z
        self.C = z["seq_lengths"].shape[0]
        self.P = int(z["P"])
        self.steps = int(z["steps"])
        if "Y" in z.files:
            Y = z["Y"]
        else:  # compact fixture: Y comes from the seeded generator, pinned by the sha256 of what the reference was fed
            import hashlib

            from gpmdm_b200 import synthetic

            C_, D_, spc_, frames_, seed_ = (int(v) for v in z["gen_cfg"])
            wl = synthetic.make_sequences(C_, D_, spc_, frames_, seed=seed_, n_test_trials=1, test_frames=8)
            Y = np.concatenate([s for cls in wl.sequences for s in cls], 0)
            assert hashlib.sha256(np.ascontiguousarray(Y).tobytes()).digest() == bytes(z["Y_sha256"].tobytes()), \
                "the synthetic generator no longer reproduces the observations this fixture was recorded on"
        self.spec = orc.ModelSpec(
            X=t64(z["X"]), Y=t64(Y), seq_lengths=[[int(v) for v in row] for row in z["seq_lengths"]],
            y_log_lengthscales=t64(z["y_log_lengthscales"]), y_log_lambdas=t64(z["y_log_lambdas"]),
            y_log_sigma_n=t64(z["y_log_sigma_n"]), x_log_lengthscales=t64(z["x_log_lengthscales"]),
            x_log_lambdas=t64(z["x_log_lambdas"]), x_log_sigma_n=t64(z["x_log_sigma_n"]),
            x_log_lin_coeff=t64(z["x_log_lin_coeff"]))
        self.T = torch.as_tensor(z["T_f32"])  # float32, cast by the filter like gpmdm_pf.py:71
        self.init_idx = [torch.as_tensor(z[f"init_idx_{c}"]) for c in range(self.C)]

    def reference_factors(self):
        """Factors built from the REFERENCE's own inverses (dense Ky_inv, diagonal blocks of Kx_inv_class)."""
        z = self.z
        return orc.precompute_factors(self.spec, Ky_inv=t64(z["Ky_inv"]),
                                      Kx_inv_blocks=[t64(z[f"Kx_inv_block_{c}"]) for c in range(self.C)])

    def step(self, t):
        pre = f"s{t}_"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}


def rel_err(a, b):
    a, b = t64(a), t64(b)
    return float(torch.max(torch.abs(a - b) / torch.clamp(torch.abs(b), min=1e-300)))


def scaled_err(a, b, scale):
    a, b = t64(a), t64(b)
    return float(torch.max(torch.abs(a - b) / t64(scale)))


# ---- product-side helpers (GPU tests) -----------------------------------------------------------------
def product_model_from_spec(spec, Ky_inv=None, Kx_inv_blocks=None, own_factors=True):
    """Build the product `GPMDM` (CUDA) holding exactly the latents / hyper-parameters of an oracle
    ModelSpec, optionally with injected inverses (the reference's own, for stage-wise parity)."""
    from gpmdm_b200 import GPMDM

    d, D, C = spec.d, spec.D, spec.n_classes
    ones = lambda n: np.ones(n)
    m = GPMDM(D=D, d=d, n_classes=C, dyn_target="full", dyn_back_step=1,
              y_lambdas_init=ones(D), y_lengthscales_init=ones(d), y_sigma_n_init=1.0,
              x_lambdas_init=ones(d), x_lengthscales_init=ones(d), x_sigma_n_init=1.0,
              x_lin_coeff_init=ones(d + 1))
    Y = spec.Y.numpy().astype(np.float32)
    s = 0
    for c, lens in enumerate(spec.seq_lengths):
        for L in lens:
            m.add_data(Y[s:s + L], c)
            s += L
    with torch.no_grad():
        for k in ("y_log_lengthscales", "y_log_lambdas", "y_log_sigma_n", "x_log_lengthscales",
                  "x_log_lambdas", "x_log_sigma_n", "x_log_lin_coeff"):
            getattr(m, k).data.copy_(getattr(spec, k).to(m.device))
    m._precompute_class_matrices()
    m.X = torch.nn.Parameter(spec.X.to(m.device).clone(), requires_grad=False)
    if own_factors:
        m._precompute_kernel_inverses()
    else:  # large N: skip the product's own factorisation, both inverses are injected
        assert Ky_inv is not None and Kx_inv_blocks is not None
        m._Xin, m._Xout, _ = m.get_Xin_Xout_matrices(m.X.detach())
        m._Xin, m._Xout = m._Xin.contiguous(), m._Xout.contiguous()
    if Ky_inv is not None or Kx_inv_blocks is not None:
        m.set_inverses(Ky_inv, Kx_inv_blocks)
    return m


def synthetic_spec(C, d, D, seqs_per_class, frames, sigma_n=1e-1, seed=0):
    """Oracle ModelSpec on seeded synthetic sequences with PCA latents (no reference needed)."""
    from sklearn.decomposition import PCA

    from gpmdm_b200 import synthetic

    wl = synthetic.make_sequences(C, D, seqs_per_class, frames, seed=seed, n_test_trials=2, test_frames=8)
    Y = np.concatenate([s for cls in wl.sequences for s in cls], 0)
    X0 = PCA(n_components=d).fit_transform(Y)
    hp = synthetic.notebook_hyperparameters(D, d, sigma_n)
    lg = lambda v: torch.log(t64(v))
    spec = orc.ModelSpec(
        X=t64(X0), Y=t64(Y), seq_lengths=[[frames] * seqs_per_class for _ in range(C)],
        y_log_lengthscales=lg(hp["y_lengthscales_init"]), y_log_lambdas=lg(hp["y_lambdas_init"]),
        y_log_sigma_n=lg(hp["y_sigma_n_init"]), x_log_lengthscales=lg(hp["x_lengthscales_init"]),
        x_log_lambdas=lg(hp["x_lambdas_init"]), x_log_sigma_n=lg(hp["x_sigma_n_init"]),
        x_log_lin_coeff=lg(hp["x_lin_coeff_init"]))
    return spec, wl
