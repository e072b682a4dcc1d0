"""GPU: error of the tensor-core dynamics variance (gpmdm_pf_dynvar_tc + the fp64 low-rank finish) against the fp64 kernel
on the benchmark model (BASELINE configs[2] shape by default), one JSON line.   python tools/dynvar_check.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--seqs-per-class", type=int, default=10)
    ap.add_argument("--frames", type=int, default=250)
    ap.add_argument("--latent", type=int, default=3)
    ap.add_argument("--sample", type=int, default=16384)
    o = ap.parse_args()
    a = argparse.Namespace(classes=o.classes, seqs_per_class=o.seqs_per_class, frames=o.frames, latent=o.latent, obs_dim=62)
    wl, X0, hp = bench.synthetic_inputs(a)
    model = bench.build_product_model(a, wl, X0, hp)
    N = X0.shape[0]
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, N, (o.sample,), generator=g)
    xs = (torch.tensor(X0[idx.numpy()]) + 0.3 * torch.randn(o.sample, a.latent, dtype=torch.float64, generator=g)).cuda()
    lam = (torch.exp(model.x_log_lambdas.detach()) ** -2)[0]
    c2 = torch.exp(model.x_log_lin_coeff.detach()) ** 2
    prior = 1.0 + (xs * xs * c2[:-1]).sum(1) + c2[-1]
    out = {"workload": f"C={a.classes} N={N} d={a.latent} sample={o.sample}, class 0 block", "prior_median": float(prior.median()),
           "prior_max": float(prior.max())}
    _, v64 = model.map_x_dynamics_for_class(xs, 0, low_latency=False)
    v64 = v64[:, 0] / lam
    out["v64_median"], out["v64_min"] = float(v64.median()), float(v64.min())
    for prec in ("tf32", "f16x2"):
        _, v = model.map_x_dynamics_for_class(xs, 0, precision=prec)
        e = (v[:, 0] / lam - v64).abs()
        out[prec] = {"abs_max": float(e.max()), "abs_median": float(e.median()), "rel_max": float((e / v64).max()),
                     "rel_median": float((e / v64).median())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
