"""Time and peak memory of the factor precompute (GPMDM._precompute_kernel_inverses; reference gpmdm.py:1284-1305) at
BASELINE configs[2] (C=8, N=20 000, d=3) and configs[3] (C=64, N=50 176, d=8) sizes, one JSON line per config.
    python tools/factor_bench.py [--cfg 3 4] [--old]      (--old: also the round-1 recipe with dense intermediates)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

CFG = {3: dict(classes=8, seqs_per_class=25, frames=100, latent=3, obs_dim=62),
       4: dict(classes=64, seqs_per_class=8, frames=98, latent=8, obs_dim=62)}


def old_recipe(model):
    """Round 1: upper Cholesky, solve against eye(N), U^-1 U^-T, then gpmdm_pack_quadform_f64 from the dense inverse."""
    from gpmdm_b200 import _cabi
    from gpmdm_b200._cabi import check, ptr, stream
    lib = _cabi.lib()
    X = model.X.detach()
    K = model.get_y_kernel(X, X)
    U, _ = torch.linalg.cholesky_ex(K, upper=True)
    eye = torch.eye(K.shape[0], dtype=K.dtype, device=K.device)
    Ui = torch.linalg.solve_triangular(U, eye, upper=True)
    Kinv = Ui @ Ui.t()
    n = K.shape[0]
    n_pad = (n + 255) // 256 * 256
    L = torch.empty(int(lib.gpmdm_quadform_bytes(n_pad, 1)) // 8, dtype=K.dtype, device=K.device)
    check(lib.gpmdm_pack_quadform_f64(ptr(Kinv), n, n_pad, 1, ptr(L), stream()), "pack")
    A = Kinv.t() @ model._Y_device()
    return L, A


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, nargs="+", default=[3, 4])
    ap.add_argument("--old", action="store_true")
    ap.add_argument("--precisions", nargs="+", default=["fp64"])
    o = ap.parse_args()
    from gpmdm_b200 import GPMDM
    GPMDM.default_factor_precisions = tuple(o.precisions)
    for c in o.cfg:
        a = argparse.Namespace(**CFG[c])
        wl, X0, hp = bench.synthetic_inputs(a)
        N = X0.shape[0]
        n2 = N * N * 8
        model = bench.build_product_model(a, wl, X0, hp)  # first build: warms cuSOLVER / cuBLAS handles
        torch.cuda.synchronize()
        del model._obs_blk, model._dyn_blks
        model._packed = None
        torch.cuda.empty_cache()
        base = torch.cuda.memory_allocated()
        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        model._precompute_kernel_inverses()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out = {"config": f"C={a.classes} N={N} d={a.latent} D={a.obs_dim}", "precisions": o.precisions,
               "precompute_s": dt, "peak_over_baseline_N2_doubles": (torch.cuda.max_memory_allocated() - base) / n2,
               "resident_after_N2_doubles": (torch.cuda.memory_allocated() - base) / n2,
               "peak_gb": (torch.cuda.max_memory_allocated() - base) / 1e9}
        if o.old:
            del model._obs_blk
            torch.cuda.empty_cache()
            base = torch.cuda.memory_allocated()
            torch.cuda.reset_peak_memory_stats()
            t0 = time.perf_counter()
            r = old_recipe(model)
            torch.cuda.synchronize()
            out["round1_recipe_obs_block_s"] = time.perf_counter() - t0
            out["round1_recipe_peak_N2_doubles"] = (torch.cuda.max_memory_allocated() - base) / n2
            del r
        print(json.dumps(out), flush=True)
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
