"""GPU: sweep of the leaf sizes of the block recursions behind the training backward (K^-1 from the Cholesky factor) and the
factor precompute, at N = 20 000.   python tools/leaf_sweep.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gpmdm_b200.gpmdm as G


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.time() - t0)
    return best * 1e3


def main():
    n = 20000
    g = torch.Generator(device="cuda").manual_seed(0)
    M = torch.randn(n, 64, dtype=torch.float64, device="cuda", generator=g)
    K = M @ M.t() / 64 + torch.eye(n, dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(K)
    del K, M
    ref = None
    out = {"n": n, "spd_inverse_from_cholesky_ms": {}, "tril_inverse_inplace_ms": {}}
    for inv_leaf in (2560, 1280, 640, 320):
        for trmm_leaf in (2048, 1024, 512):
            G._INV_LEAF, G._TRMM_LEAF = inv_leaf, trmm_leaf
            t = timed(lambda: G.spd_inverse_from_cholesky(L))
            out["spd_inverse_from_cholesky_ms"][f"inv_leaf={inv_leaf},trmm_leaf={trmm_leaf}"] = round(t, 1)
    G._INV_LEAF = 2560
    for tri_leaf in (1024, 512, 256):
        for trmm_leaf in (2048, 1024, 512):
            G._TRINV_LEAF, G._TRMM_LEAF = tri_leaf, trmm_leaf
            W = L.clone()
            torch.cuda.synchronize()
            t0 = time.time()
            G.tril_inverse_inplace(W)
            torch.cuda.synchronize()
            out["tril_inverse_inplace_ms"][f"trinv_leaf={tri_leaf},trmm_leaf={trmm_leaf}"] = round((time.time() - t0) * 1e3, 1)
            if ref is None:
                ref = W.clone()
            else:
                out.setdefault("max_rel_diff_between_settings", 0.0)
                out["max_rel_diff_between_settings"] = max(out["max_rel_diff_between_settings"],
                                                           float((W - ref).abs().max() / ref.abs().max()))
            del W
    print(json.dumps(out))


if __name__ == "__main__":
    main()
