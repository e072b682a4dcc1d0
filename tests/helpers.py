"""Shared test plumbing: golden-fixture loading (tests/golden/*.npz, written by oracle/make_golden.py
from the unmodified reference) and tolerance helpers."""
import glob
import os

import numpy as np
import torch

from oracle import gpmdm_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))

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
        self.spec = orc.ModelSpec(
            X=t64(z["X"]), Y=t64(z["Y"]), seq_lengths=[[int(v) for v in row] for row in z["seq_lengths"]],
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
