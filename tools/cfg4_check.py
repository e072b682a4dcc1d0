"""BASELINE config 4 at full factor size: 64-class GPMDM, N_train = 50 176, latent d = 8, D = 62.
Tolerance check of the tensor-core variants (tf32x3, f16x2) against the fp64 exact path on a particle sample (the fp64 path at N = 50 k costs
2.5 GFLOP per particle), printed as one JSON line.   python tools/cfg4_check.py [--sample 512]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sample", type=int, default=512)
    o = ap.parse_args()
    a = argparse.Namespace(classes=64, seqs_per_class=8, frames=98, latent=8, obs_dim=62)
    t0 = time.time()
    wl, X0, hp = bench.synthetic_inputs(a)
    model = bench.build_product_model(a, wl, X0, hp)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    N = X0.shape[0]
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, N, (o.sample,), generator=g)
    xs = (torch.tensor(X0[idx.numpy()]) + 0.3 * torch.randn(o.sample, a.latent, dtype=torch.float64, generator=g)).cuda()
    t0 = time.time()
    mu64, var64 = model.map_x_to_y(xs)
    torch.cuda.synchronize()
    t64 = time.time() - t0
    scale = torch.clamp(mu64.abs().max(dim=1, keepdim=True).values, min=1e-2)
    v64 = var64[:, 0]  # lambda = 1
    ok = v64 > 0.05
    out = {
        "config": f"BASELINE configs[3]: 64-class GPMDM, N_train={N}, d={a.latent}, D={a.obs_dim}, {o.sample}-particle sample",
        "frac_v_gt_5pct": float(ok.double().mean()), "v_min": float(v64.min()), "v_median": float(v64.median()),
        "build_s": t_build, "fp64_pack_and_run_s": t64,
    }
    for prec in ("tf32", "f16x2"):
        t0 = time.time()
        mu32, var32 = model.map_x_to_y(xs, precision=prec)
        torch.cuda.synchronize()
        v32 = var32[:, 0]
        out[prec] = {
            "mean_err_rel_max": float(torch.max(torch.abs(mu32 - mu64) / scale)),
            "var_err_of_prior_max": float(torch.max(torch.abs(v32 - v64))),
            "var_err_rel_max_where_v_gt_5pct": float(torch.max(torch.abs(v32[ok] - v64[ok]) / v64[ok])) if bool(ok.any()) else None,
            "pack_and_run_s": time.time() - t0,
        }
    out["gpu_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
