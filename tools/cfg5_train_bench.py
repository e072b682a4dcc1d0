"""BASELINE config 5: training kernel build (block-masked dynamics + observation kernel matrices) and the NLL gradient
terms at N_train = 20 000, feeding torch.linalg Cholesky.  Times each CUDA kernel with CUDA events and reports achieved
HBM GB/s against the measured copy peak (MEASURED_PEAKS.json: 6453 GB/s), plus one full `gpdm_loss` forward+backward.

    python tools/cfg5_train_bench.py [--n-classes 8 --seqs-per-class 25 --frames 100]
Algorithmic bytes: build = 8 N^2 written; gradient = 8 N^2 read (every mirror-image tile pair of G once) + 8 N d written."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench

HBM_PEAK = 6453.1


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--seqs-per-class", type=int, default=25)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--profile", action="store_true", help="also print the top CUDA kernels of one forward + backward (torch.profiler) to stderr")
    o = ap.parse_args()
    a = argparse.Namespace(classes=o.classes, seqs_per_class=o.seqs_per_class, frames=o.frames, latent=3, obs_dim=62)
    from gpmdm_b200 import GPMDM
    from gpmdm_b200.gpmdm import _KernelBuild

    wl, X0, hp = bench.synthetic_inputs(a)
    m = GPMDM(D=62, d=3, n_classes=o.classes, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(o.classes):
        for s in wl.sequences[c]:
            m.add_data(s, c)
    m._precompute_class_matrices()
    m.X = torch.nn.Parameter(torch.tensor(X0, dtype=torch.float64, device=m.device))
    N = X0.shape[0]
    Xin, Xout, _ = m.get_Xin_Xout_matrices()
    Xin = Xin.detach().contiguous()
    Nx = Xin.shape[0]
    offs = m._pair_offsets_dev()
    out = {"N": N, "Nx": Nx, "hbm_peak_gbs": HBM_PEAK}
    # ---- the CUDA kernels themselves, through the C ABI (no autograd / allocator overhead in the timed region)
    from gpmdm_b200 import _cabi
    from gpmdm_b200._cabi import check, ptr, stream
    lib = _cabi.lib()
    dev = m.device
    X = m.X.detach().contiguous()
    ls_y = torch.exp(m.y_log_lengthscales.detach()).contiguous()
    ls_x = torch.exp(m.x_log_lengthscales.detach()).contiguous()
    c2 = (torch.exp(m.x_log_lin_coeff.detach()) ** 2).contiguous()
    s2y = float(torch.exp(m.y_log_sigma_n.detach()) ** 2)
    s2x = float(torch.exp(m.x_log_sigma_n.detach()) ** 2)
    K = torch.empty(N, N, dtype=torch.float64, device=dev)
    t = timed(lambda: check(lib.gpmdm_kernel_build_f64(ptr(X), N, 3, 0, ptr(ls_y), None, s2y, None, 0, ptr(K), stream()), "build"))
    out["build_Ky_ms"], out["build_Ky_gbs"] = t, 8.0 * N * N / (t * 1e-3) / 1e9
    Kx = K[:Nx * Nx // N].reshape(-1)[:Nx * Nx].view(Nx, Nx) if Nx * Nx <= N * N else torch.empty(Nx, Nx, dtype=torch.float64, device=dev)
    t = timed(lambda: check(lib.gpmdm_kernel_build_f64(ptr(Xin), Nx, 3, 1, ptr(ls_x), ptr(c2), s2x, ptr(offs), o.classes, ptr(Kx), stream()), "build"))
    out["build_Kx_masked_ms"], out["build_Kx_masked_gbs"] = t, 8.0 * Nx * Nx / (t * 1e-3) / 1e9
    G = torch.randn(N, N, dtype=torch.float64, device=dev)
    gX = torch.empty(N, 3, dtype=torch.float64, device=dev)
    gl, gs, gc = (torch.empty(k, dtype=torch.float64, device=dev) for k in (3, 1, 4))
    ws = torch.empty(int(lib.gpmdm_kernel_grad_workspace_bytes(N, 3)) // 8 + 1, dtype=torch.float64, device=dev)
    t = timed(lambda: check(lib.gpmdm_kernel_grad_f64(ptr(X), ptr(G), N, 3, 0, ptr(ls_y), None, s2y, None, 0, ptr(gX), ptr(gl), ptr(gs), None, ptr(ws), stream()), "grad"))
    out["grad_Ky_ms"], out["grad_Ky_gbs"] = t, 8.0 * N * N / (t * 1e-3) / 1e9
    Gx = G.reshape(-1)[:Nx * Nx].view(Nx, Nx)
    t = timed(lambda: check(lib.gpmdm_kernel_grad_f64(ptr(Xin), ptr(Gx), Nx, 3, 1, ptr(ls_x), ptr(c2), s2x, ptr(offs), o.classes, ptr(gX), ptr(gl), ptr(gs), ptr(gc), ptr(ws), stream()), "grad"))
    blocks = float(sum((int(offs[i + 1]) - int(offs[i])) ** 2 for i in range(o.classes)))  # only class blocks of G are read
    out["grad_Kx_masked_ms"], out["grad_Kx_masked_gbs"] = t, 8.0 * blocks / (t * 1e-3) / 1e9
    del K, G, Kx, Gx
    for p in m.parameters():
        p.requires_grad_(True)
    # one full loss forward + backward (kernel builds + torch.linalg Cholesky / triangular solves + our gradient kernels)
    Y = m._Y_device()

    def step():
        for p in m.parameters():
            p.grad = None
        loss = m.gpdm_loss(Y, N)
        loss.backward()
        return loss

    t = timed(step, reps=2)
    out["gpdm_loss_fwd_bwd_ms"] = t
    out["loss"] = float(step().detach())
    if o.profile:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70), file=sys.stderr)
    for k in list(out):
        if k.endswith("_gbs"):
            out[k.replace("_gbs", "_frac_of_hbm_peak")] = out[k] / HBM_PEAK
    print(json.dumps(out))


if __name__ == "__main__":
    main()
