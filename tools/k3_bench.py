"""HBM roofline of the reduction / scan / resampling kernels (north_star item 3) at a particle count large enough to be
bandwidth-bound (default 2^24): CUDA-event time per C-ABI call, algorithmic bytes (56 + 16(d+1) per particle in
total; per call as listed below) and the bytes each call actually moves, against the measured copy peak 6453 GB/s.

    python tools/k3_bench.py [--particles 16777216] [--d 3] [--classes 8]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

HBM_PEAK = 6453.1


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1 << 24)
    ap.add_argument("--d", type=int, default=3)
    ap.add_argument("--classes", type=int, default=8)
    a = ap.parse_args()
    from gpmdm_b200 import _cabi
    from gpmdm_b200._cabi import check, ptr, stream
    lib = _cabi.lib()
    P, d, C = a.particles, a.d, a.classes
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    ll = -1e3 * torch.rand(P, dtype=torch.float64, device=dev, generator=g) ** 4   # a few dominant particles
    lw, w, cdf = (torch.empty(P, dtype=torch.float64, device=dev) for _ in range(3))
    u = torch.rand(P, dtype=torch.float64, device=dev, generator=g)
    us = (torch.arange(P, dtype=torch.float64, device=dev) + 0.37) / P
    x = torch.randn(P, d, dtype=torch.float64, device=dev, generator=g)
    c = torch.randint(0, C, (P,), device=dev, generator=g)
    xo, co, anc = torch.empty_like(x), torch.empty_like(c), torch.empty(P, dtype=torch.int64, device=dev)
    E = torch.rand(P, C, dtype=torch.float64, device=dev, generator=g) + 0.01
    T = torch.full((C, C), 0.1 / (C - 1), dtype=torch.float64, device=dev); T.fill_diagonal_(0.9)
    cn = torch.empty_like(c)
    perm = torch.empty(P, dtype=torch.int32, device=dev)
    tiles = torch.empty(P // 64 + C + 1, 4, dtype=torch.int32, device=dev)
    nt = torch.empty(1, dtype=torch.int32, device=dev)
    stats = torch.empty(2, dtype=torch.float64, device=dev)
    out = torch.empty(C + d + 1, dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.gpmdm_workspace_bytes(P, C)) // 8 + 1, dtype=torch.float64, device=dev)
    st = stream()
    calls = {
        # name: (callable, algorithmic bytes per particle, bytes per particle this implementation moves)
        "transition": (lambda: check(lib.gpmdm_pf_transition_f64(ptr(c), ptr(T), ptr(E), P, C, ptr(cn), st), "t"),
                       16 + 8 * C, 16 + 8 * C),
        "bucket_by_class": (lambda: check(lib.gpmdm_pf_bucket_by_class(ptr(c), P, C, ptr(perm), ptr(tiles), ptr(nt), ptr(ws), st), "b"),
                            8 + 4, 8 + 8 + 4),
        "normalize": (lambda: check(lib.gpmdm_pf_normalize_f64(ptr(ll), P, ptr(lw), ptr(w), ptr(stats), ptr(ws), st), "n"),
                      8 + 16, 8 + 24 + 16),
        "cdf_blocked": (lambda: check(lib.gpmdm_pf_cdf_f64(ptr(w), P, 1, ptr(cdf), ptr(ws), st), "c"), 16, 16 + 16),
        "resample_multinomial": (lambda: check(lib.gpmdm_pf_resample_f64(ptr(cdf), P, ptr(u), P, ptr(x), ptr(c), d, ptr(anc), ptr(xo), ptr(co), st), "r"),
                                 8 + 8 + 2 * (8 * d + 8), 8 + 8 + 2 * (8 * d + 8)),
        "resample_systematic": (lambda: check(lib.gpmdm_pf_resample_sorted_f64(ptr(cdf), P, ptr(us), P, ptr(x), ptr(c), d, ptr(anc), ptr(xo), ptr(co), st), "r"),
                                8 + 8 + 2 * (8 * d + 8), 8 + 8 + 2 * (8 * d + 8)),
        "summaries": (lambda: check(lib.gpmdm_pf_summaries_f64(ptr(ll), ptr(lw), ptr(w), ptr(co), ptr(xo), P, C, d, ptr(out), ptr(ws), st), "s"),
                      24 + 8 + 8 * d, 16 + 24 + 8 + 8 * d),
    }
    res = {"particles": P, "d": d, "classes": C, "hbm_peak_gbs": HBM_PEAK, "kernels": {}}
    calls["normalize"][0](); calls["cdf_blocked"][0](); torch.cuda.synchronize()
    for name, (fn, alg, moved) in calls.items():
        ms = timed(fn)
        res["kernels"][name] = {"ms": ms, "algorithmic_gbs": alg * P / (ms * 1e-3) / 1e9, "moved_gbs": moved * P / (ms * 1e-3) / 1e9,
                                "frac_of_peak_algorithmic": alg * P / (ms * 1e-3) / 1e9 / HBM_PEAK,
                                "frac_of_peak_moved": moved * P / (ms * 1e-3) / 1e9 / HBM_PEAK}
    total_ms = sum(v["ms"] for k, v in res["kernels"].items() if k not in ("resample_systematic", "transition", "bucket_by_class"))
    res["k3_total_ms"] = total_ms
    res["k3_algorithmic_gbs"] = (56 + 16 * (d + 1)) * P / (total_ms * 1e-3) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
