#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the GPMDM filter step (BASELINE.json metric) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
8-class GPMDM, N_train = 20 000 frames (200 sequences x 100), D = 62, latent d = 3, P = 1 048 576
particles in fp64, sharded by contiguous particle range over the ranks (strong scaling: P is fixed).
A "step" is one full `GPMDM_PF.update(z)` -- class transition, dynamics-GP draw, observation-GP
log-likelihood, weight normalisation, multinomial resampling -- plus the class-posterior query, on
synthetic data (gpmdm_b200/synthetic.py), device-side Philox draws.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the
public API with the observation in pinned host memory and the class posterior read back to the host.
`roofline` describes the dominant kernel (the observation-GP contraction, gp_predict_kernel<0,3>),
`cpu_baseline` / `--impl reference` time the CPU oracle (a restatement of the reference's torch-CPU
algorithm; the reference itself is Python and is not present on the GPU box) on a bounded particle sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "particle_updates_per_sec"
UNIT = "particle-updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--seqs-per-class", type=int, default=25)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--latent", type=int, default=3)
    ap.add_argument("--obs-dim", type=int, default=62)
    ap.add_argument("--cpu-sample", type=int, default=4096, help="particles of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense", action="store_true", help="dense K^-1 instead of the triangular packing")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "tf32", "f16x2"],
                    help="fp64 = exact path (the headline); tf32 / f16x2 = tcgen05 variants of the observation GP (config 4)")
    return ap.parse_args()


def workload_name(a):
    n = a.classes * a.seqs_per_class * a.frames
    pr = getattr(a, "precision", "fp64")
    which = "configs[3] (fp32-accuracy variant)" if pr != "fp64" and a.classes == 64 else "configs[2]"
    prec = {"fp64": "fp64", "tf32": "tf32x3 variances + fp64 means/dynamics",
            "f16x2": "fp16-split (2 x fp16, 3 MMAs) variances + fp64 means/dynamics"}[pr]
    return (f"BASELINE {which}: {a.classes}-class GPMDM, N_train={n}, D={a.obs_dim}, d={a.latent}, "
            f"P={a.particles} particles, {prec}")


# ---- synthetic model ------------------------------------------------------------------------------------
def synthetic_inputs(a):
    from sklearn.decomposition import PCA

    from gpmdm_b200 import synthetic

    wl = synthetic.make_sequences(a.classes, a.obs_dim, a.seqs_per_class, a.frames, seed=0, n_test_trials=1,
                                  test_frames=64)
    Y = np.concatenate([s for cls in wl.sequences for s in cls], 0)
    X0 = PCA(n_components=a.latent).fit_transform(Y)
    hp = synthetic.notebook_hyperparameters(a.obs_dim, a.latent, sigma_n=1e-1)
    return wl, X0, hp


def build_product_model(a, wl, X0, hp):
    from gpmdm_b200 import GPMDM

    if getattr(a, "precision", "fp64") != "fp64":  # tensor-core variants: never build (or hold) the fp64 panels
        GPMDM.default_factor_precisions = (a.precision,)
    m = GPMDM(D=a.obs_dim, d=a.latent, n_classes=a.classes, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(a.classes):
        for s in wl.sequences[c]:
            m.add_data(s, c)
    m._precompute_class_matrices()
    m.X = torch.nn.Parameter(torch.tensor(X0, dtype=torch.float64, device=m.device), requires_grad=False)
    m._precompute_kernel_inverses()
    return m


# ---- clocks ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].startswith("Active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ---- CPU oracle timing ------------------------------------------------------------------------------------------
def oracle_spec(a, wl, X0, hp):
    from oracle import gpmdm_oracle as orc

    t64 = lambda v: torch.as_tensor(np.asarray(v), dtype=torch.float64)
    Y = np.concatenate([s for cls in wl.sequences for s in cls], 0)
    lg = lambda v: torch.log(t64(v))
    return orc.ModelSpec(
        X=t64(X0), Y=t64(Y), seq_lengths=[[a.frames] * a.seqs_per_class for _ in range(a.classes)],
        y_log_lengthscales=lg(hp["y_lengthscales_init"]), y_log_lambdas=lg(hp["y_lambdas_init"]),
        y_log_sigma_n=lg(hp["y_sigma_n_init"]), x_log_lengthscales=lg(hp["x_lengthscales_init"]),
        x_log_lambdas=lg(hp["x_lambdas_init"]), x_log_sigma_n=lg(hp["x_sigma_n_init"]),
        x_log_lin_coeff=lg(hp["x_lin_coeff_init"]))


def oracle_factors(spec, device):
    """The oracle's factor recipe with the O(N^3) part evaluated by plain torch on `device` (setup only -- it is not
    part of any timed region), returned on the CPU."""
    from oracle import gpmdm_oracle as orc

    return orc.precompute_factors_on(spec, device)


def time_oracle(a, wl, spec, factors, sample, steps, warmup, loop_ll=True):
    """Times `FilterOracle.update` + class query (the region the reference's notebook times,
    test_gpmdm_pf.ipynb cell 4) on `sample` particles with all host threads.  Also returns the oracle with the trace
    of its last step and that step's inputs (the parity block compares the CUDA path on exactly those)."""
    from gpmdm_b200 import synthetic
    from oracle import gpmdm_oracle as orc

    torch.set_num_threads(os.cpu_count() or 1)
    C, d = spec.n_classes, spec.d
    T = synthetic.markov_matrix(C)
    parts = orc.divide_into_n_parts(sample, C)
    g = torch.Generator().manual_seed(1)
    init_idx = [torch.randint(0, hi - lo, (parts[c],), generator=g) for c, (lo, hi) in enumerate(spec.class_row_ranges())]
    o = orc.FilterOracle(spec, T, sample, init_idx, factors)
    trial = wl.test_trials[0][1]
    times, last = [], None
    with torch.no_grad():
        for t in range(warmup + steps):
            E, eps, u = synthetic.raw_draws(sample, C, d, 100 + t)
            last = dict(x_prev=o.states.clone(), c_prev=o.classes.clone(), z=trial[t % trial.shape[0]], E=E, eps=eps, u=u)
            t0 = time.perf_counter()
            o.update(last["z"], E, eps, u, loop_ll=loop_ll)  # loop_ll: the reference's per-particle loop
            o.get_most_likely_class()
            dt = time.perf_counter() - t0
            if t >= warmup:
                times.append(dt)
    return sample * len(times) / sum(times), 1e3 * sum(times) / len(times), torch.get_num_threads(), o, last


def parity_block(pf, model, o, last, T):
    """The CUDA path (the kernel instances of the timed region: fused observation kernel with the K* cache, fused
    dynamics kernel, filter stages) on the inputs of the oracle's last step, against the oracle's outputs for that step.
    Runs outside every timed region.  Reference lines: gpmdm.py:923-963, :1032-1068, gpmdm_pf.py:137-213."""
    import ctypes

    from gpmdm_b200 import _cabi
    from gpmdm_b200._cabi import check, ptr, stream

    lib = _cabi.lib()
    tr, spec = o.trace, o.m
    n, C, d, D = o.P, spec.n_classes, spec.d, spec.D
    dev = model.device
    f64 = torch.float64
    cu = lambda t: t.to(dev).contiguous()
    # observation GP + fused log-likelihood on the oracle's pre-resample states
    x = cu(tr["x_new"])
    z = torch.as_tensor(np.asarray(last["z"]), dtype=f64, device=dev)
    ll, mu, v = (torch.empty(n, dtype=f64, device=dev), torch.empty(n, D, dtype=f64, device=dev),
                 torch.empty(n, dtype=f64, device=dev))
    pk = pf._packed
    counter = torch.zeros(4, dtype=torch.int32, device=dev)
    if pf._kstar_cache:
        check(lib.gpmdm_pf_observe_cached_f64(ctypes.byref(pk["obs"]), ptr(x), n, ptr(z), pf._ll_const, ptr(ll), ptr(mu),
                                              ptr(v), pk["obs_n_pad"], ptr(counter), ptr(pf._ws_kstar),
                                              pf._ws_kstar.numel() * 8, stream()), "gpmdm_pf_observe_cached_f64")
        kern = "gpmdm_pf_observe_cached_f64"
    else:
        check(lib.gpmdm_pf_observe_f64(ctypes.byref(pk["obs"]), ptr(x), n, ptr(z), pf._ll_const, ptr(ll), ptr(mu), ptr(v),
                                       ptr(counter), stream()), "gpmdm_pf_observe_f64")
        kern = "gpmdm_pf_observe_f64"
    mu, v, ll = mu.cpu(), v.cpu(), ll.cpu()
    mscale = torch.clamp(tr["mu"].abs().max(dim=1, keepdim=True).values, min=1e-3)
    ok = tr["v"] > 1e-3
    ll_rel = (ll - tr["ll"]).abs() / tr["ll"].abs()
    # ll = -S/(2v) - D log v + const: an error dv in the variance moves it by (S/(2 v^2) + D/v) dv.  `ll_err_vs_bound` <= 1
    # means every ll difference is explained by variances matched to 1e-9 of the prior plus 1e-9 of the terms' magnitude.
    lam2 = torch.exp(spec.y_log_lambdas) ** 2
    S = (lam2.unsqueeze(0) * (torch.as_tensor(np.asarray(last["z"]), dtype=f64).unsqueeze(0) - tr["mu"]) ** 2).sum(1)
    vo = tr["v"]
    terms = S / (2 * vo) + D * vo.log().abs() + abs(pf._ll_const)
    ll_bound = (S / (2 * vo * vo) + D / vo) * 1e-9 + 1e-9 * terms
    ll_vs_bound = (ll - tr["ll"]).abs() / ll_bound
    # dynamics GP (fused kernel) per class on the oracle's previous states
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    dm = dv = 0.0
    for c in range(C):
        rows = torch.nonzero(tr["c_new"] == c).squeeze(-1)
        if rows.numel() == 0:
            continue
        xs = last["x_prev"][rows]
        mean, var = model.map_x_dynamics_for_class(cu(xs), c, low_latency=False)
        prior = (1.0 + (xs * xs * torch.exp(spec.x_log_lin_coeff[:-1]) ** 2).sum(1)
                 + torch.exp(spec.x_log_lin_coeff[-1]) ** 2).unsqueeze(1) * lam_x.unsqueeze(0)
        sc = torch.clamp(tr["dyn_mean"][rows].abs().max(dim=1, keepdim=True).values, min=1e-3)
        dm = max(dm, float(((mean.cpu() - tr["dyn_mean"][rows]).abs() / sc).max()))
        dv = max(dv, float(((var.cpu() - tr["dyn_var"][rows]).abs() / prior).max()))
    # integer stages on the oracle's own inputs: class transition, normalise + sequential cdf + search
    c_new = torch.empty(n, dtype=torch.int64, device=dev)
    # (device copies are bound to names: a temporary would be returned to the allocator -- and possibly handed to the next
    # argument's copy -- before the asynchronous kernel has read it)
    c_prev_d, T_d, E_d, u_d = cu(last["c_prev"]), cu(T.to(f64)), cu(last["E"]), cu(last["u"])
    check(lib.gpmdm_pf_transition_f64(ptr(c_prev_d), ptr(T_d), ptr(E_d), n, C, ptr(c_new), stream()), "gpmdm_pf_transition_f64")
    ws = torch.empty(int(lib.gpmdm_workspace_bytes(n, C)) // 8 + 1, dtype=f64, device=dev)
    lw, w, cdf, st2 = (torch.empty(n, dtype=f64, device=dev) for _ in range(4))
    anc = torch.empty(n, dtype=torch.int64, device=dev)
    ll_o = cu(tr["ll"])
    check(lib.gpmdm_pf_normalize_f64(ptr(ll_o), n, ptr(lw), ptr(w), ptr(st2), ptr(ws), stream()), "gpmdm_pf_normalize_f64")
    check(lib.gpmdm_pf_cdf_f64(ptr(w), n, 0, ptr(cdf), ptr(ws), stream()), "gpmdm_pf_cdf_f64")
    check(lib.gpmdm_pf_resample_f64(ptr(cdf), n, ptr(u_d), n, None, None, d, ptr(anc), None, None, stream()),
          "gpmdm_pf_resample_f64")
    return {
        "against": "CPU oracle (oracle/gpmdm_oracle.py, pinned to the unmodified reference by tests/golden), its last step",
        "sample_particles": n, "observe_entry_point": kern,
        "obs_mean_err_of_row_scale_max": float(((mu - tr["mu"]).abs() / mscale).max()),
        "obs_var_err_of_prior_max": float((v - tr["v"]).abs().max()),
        "ll_rel_err_max_where_v_gt_1e-3": float(ll_rel[ok].max()) if bool(ok.any()) else None,
        "ll_rel_err_median": float(ll_rel.median()),
        "ll_compared": int(ok.sum()), "v_min": float(tr["v"].min()), "v_median": float(tr["v"].median()),
        "ll_err_vs_bound_max": float(ll_vs_bound.max()),
        "ll_err_vs_bound": "|ll - ll_oracle| / ((S/(2v^2) + D/v) 1e-9 + 1e-9 (S/(2v) + D|log v| + |const|)) over ALL sample "
                           "particles: <= 1 means the difference is what a variance matched to 1e-9 of the prior allows",
        "dyn_mean_err_of_row_scale_max": dm, "dyn_var_err_of_prior_max": dv,
        "classes_equal": bool(torch.equal(c_new.cpu(), tr["c_new"])),
        "log_weights_equal": bool(torch.equal(lw.cpu(), tr["lw"])),
        "ancestors_equal": bool(torch.equal(anc.cpu(), tr["anc"])),
        "tolerances": "north_star: integers bit-exact; means 1e-9 of the row scale; variances 1e-9 of the prior (dynamics: "
                      "4e-9 -- the oracle's own fp64 evaluation of prior - k^T K^-1 k carries ~1e-9 of cancellation noise, "
                      "tests/test_gpu_vs_oracle.py); ll: ll_err_vs_bound <= 1 (DESIGN.md section 2)",
    }


def state_digest(pf):
    """sha256 over the replicated filter state after the last timed step: post-resample classes || ancestors || the
    bits of the log-likelihoods.  Identical for every GPU count iff a G-GPU run equals the 1-GPU run bit for bit."""
    import hashlib

    h = hashlib.sha256()
    for t in (pf._particle_classes, pf.last_ancestors, pf._log_likelihoods):
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


# ---- main ---------------------------------------------------------------------------------------------------
def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl, X0, hp = synthetic_inputs(a)
    spec = oracle_spec(a, wl, X0, hp)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    f = oracle_factors(spec, dev)
    warm = min(a.warmup, 1)
    val, ms, cores, _, _ = time_oracle(a, wl, spec, f, a.cpu_sample, a.steps, warm)
    sample = f"{a.cpu_sample} of {a.particles} particles per step, {a.steps} steps (+{warm} warm-up), N_train={spec.N}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU oracle port of the reference's torch-CPU filter "
                   "(the Python reference does not travel to the GPU box); per-particle cost is independent of P"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(a):
    import torch.distributed as dist

    from gpmdm_b200 import GPMDM_PF, _cabi, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its banner ("NCCL version ...") to fd 1 when the communicator
        # is created, so fd 1 points at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    lib = _cabi.lib()

    wl, X0, hp = synthetic_inputs(a)
    model = build_product_model(a, wl, X0, hp)
    N = X0.shape[0]
    C, d, D, P = a.classes, a.latent, a.obs_dim, a.particles
    T = synthetic.markov_matrix(C)
    pf = GPMDM_PF(model, T, P, seed=1234, tri=not a.dense, cdf_order="blocked", precision=a.precision)
    trial = wl.test_trials[0][1]
    z_dev = [torch.tensor(trial[t], dtype=torch.float64, device="cuda") for t in range(trial.shape[0])]
    z_pinned = [torch.tensor(trial[t]).pin_memory() for t in range(trial.shape[0])]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(t):
        pf._update(z_dev[t % len(z_dev)])
        pf._summaries()

    def step_e2e(t):
        pf.update(z_pinned[t % len(z_pinned)])          # H2D of the observation inside update()
        return pf.class_probabilities().cpu()           # D2H of the class posterior

    for t in range(a.warmup):
        step_resident(t)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    pf._profile_events = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(a.steps):
        step_resident(a.warmup + t)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    obs_ms = [ev[0].elapsed_time(ev[1]) for ev in pf._profile_events]
    tc_ms = [ev[0].elapsed_time(ev[2]) for ev in pf._profile_events if len(ev) > 2]  # tensor-core kernel alone (variants)
    pf._profile_events = None
    clk = clocks.stop() if rank == 0 else None
    digest = state_digest(pf) if rank == 0 else None  # after the last timed step, outside the timed region

    # end-to-end leg (host buffers in, host result out)
    ke = max(1, min(a.steps, 2))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(ke):
        probs = step_e2e(t)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    if world > 1:
        tt = torch.tensor([elapsed_ms, e2e_ms, sum(obs_ms) / max(len(obs_ms), 1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, obs_avg_ms = tt.tolist()
    else:
        obs_avg_ms = sum(obs_ms) / max(len(obs_ms), 1)

    if rank == 0:
        value = P * a.steps / (elapsed_ms * 1e-3)
        e2e_val = P * ke / (e2e_ms * 1e-3)
        # roofline of the dominant kernel: observation-GP contraction, one launch per step per rank
        Pl = P // world
        TN = 256                                                # column-tile width of the predict kernel
        n_pad = (N + TN - 1) // TN * TN
        nq = n_pad // TN
        secs = obs_avg_ms * 1e-3
        tiles = (Pl + 63) // 64
        if a.precision == "fp64":
            tf = ctypes_probe(lib)
            # flops the algorithm in use needs per particle (un-padded): k^T Q k on the triangular packing of the
            # symmetric K^-1 (N^2: N(N+1)/2 multiply-adds) + the mean K* alpha (2ND)
            flops_need = Pl * (1.0 * N * N + 2.0 * N * D) if not a.dense else Pl * (2.0 * N * N + 2.0 * N * D)
            # flops the kernel issues per 64-particle tile: [16 k x 256 col] DMMA chunks of the lower triangle of column
            # panels, + for the alpha tile only the groups of 8 column blocks (64 columns) that hold real outputs
            alpha_cols = min((D + 63) // 64 * 64, TN) + (max(D - TN, 0) + 63) // 64 * 64
            nkc = (N + 15) // 16  # 16-row k-chunks holding real training rows (chunks of pure padding are skipped)
            chunk_cols = sum((nkc - 16 * J if not a.dense else nkc) * TN for J in range(nq)) + nkc * alpha_cols
            flops_exec = tiles * 64 * 2.0 * 16 * chunk_cols
            flops_dense = Pl * (2.0 * N * N + 2.0 * N * D)       # the dense formulation's count
            achieved = flops_need / secs / 1e12
            cached = bool(getattr(pf, "_kstar_cache", False))
            traffic, traffic_note = measured_traffic(N, d, cached, not a.dense, tiles)
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf,
                "traffic": traffic, "traffic_note": traffic_note,
                "kernel": f"gp_predict_kernel<0,{d},{'true' if cached else 'false'}> "
                          f"({'gpmdm_pf_observe_cached_f64' if cached else 'gpmdm_pf_observe_f64'})",
                "launch_ms": obs_avg_ms, "launch_share_of_step": obs_avg_ms / (elapsed_ms / a.steps),
                "peak_source": "fp64 mma.sync m8n8k4 issue-rate probe measured in this run (MEASURED_PEAKS.json holds "
                               "no fp64 figure); = 148 SMs x 64 FMA/clk x 2 x SM clock",
                "executed_tflops": flops_exec / secs / 1e12, "executed_frac": flops_exec / secs / 1e12 / tf,
                "dense_equivalent_tflops": flops_dense / secs / 1e12,
                "note": ("achieved = (N^2 + 2ND) flops per particle, the un-padded need of the algorithm in use (quadratic form "
                         "on the triangular packing of the symmetric K^-1); executed = DMMA flops issued (16-row chunks x 256-column "
                         "tiles: padding of N to 16 rows / 256 columns, of D to 64); dense_equivalent = SURVEY 8(d)'s 2N^2 + 2ND per particle over the same time "
                         "(not a roofline fraction)") if not a.dense else "dense K^-1",
            }
        else:
            # tensor-core variants: whitened form |W k|^2 on the lower-triangular W (N^2 + 2ND algorithmic flops per
            # particle), executed as 3 MMAs per product (tf32 triple, or fp16 pair); peak = tcgen05 issue-rate probe of
            # the same MMA kind measured in this run
            f16 = a.precision == "f16x2"
            tf = ctypes_probe(lib, "gpmdm_probe_f16_tflops" if f16 else "gpmdm_probe_tf32_tflops", 20000)
            tiles128 = (Pl + 127) // 128
            flops_alg32 = Pl * (1.0 * N * N + 2.0 * N * D)
            flops_mma = 3.0 * tiles128 * 128 * (2.0 * TN * TN) * (nq * (nq + 1) / 2 + (0 if f16 else nq))
            tc_avg_ms = sum(tc_ms) / max(len(tc_ms), 1) if tc_ms else obs_avg_ms
            secs = tc_avg_ms * 1e-3  # the tensor-core kernel alone; the fp64 mean-tile kernel is reported beside it
            achieved = flops_mma / secs / 1e12
            kname = "observe_tf32_kernel<%d,%s>" % (d, "MODE_F16X2" if f16 else "MODE_TF32X3")
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": tf, "unit": "TFLOP/s", "frac": achieved / tf,
                "traffic": None, "kernel": f"{kname} ({'gpmdm_pf_observe_f16x2' if f16 else 'gpmdm_pf_observe_tf32'}) + fp64 mean tile (gpmdm_pf_loglik_f64)",
                "launch_ms": tc_avg_ms, "launch_share_of_step": tc_avg_ms / (elapsed_ms / a.steps),
                "fp64_mean_tile_ms": obs_avg_ms - tc_avg_ms,
                "peak_source": ("tcgen05.mma kind::%s M128 N256 K%d issue-rate probe (resident pseudo-random operands) measured in this "
                                "run; achieved counts the MMA flops issued (3 per product) over the tensor-core kernel's own launch "
                                "time (rank 0)" % (("f16", 16) if f16 else ("tf32", 8))),
                # kind::f16 MMAs run on the pipe the driver's cuBLAS bf16 GEMM measures (MEASURED_PEAKS.json): burst, and
                # sustained for seconds under the power cap -- the regime of this kernel inside a long step
                "measured_peaks_bf16_tflops": measured_peak("bf16_tflops"),
                "measured_peaks_bf16_tflops_sustained": measured_peak("bf16_tflops_sustained"),
                "frac_of_measured_bf16_sustained": (achieved / measured_peak("bf16_tflops_sustained"))
                if (f16 and measured_peak("bf16_tflops_sustained")) else None,
                "algorithmic_tflops": flops_alg32 / secs / 1e12, "algorithmic_frac": flops_alg32 / secs / 1e12 / tf,
            }
        cpu, parity = None, None
        if not a.no_cpu_baseline:
            spec = oracle_spec(a, wl, X0, hp)
            f = oracle_factors(spec, "cuda")
            if world == 1:
                v, ms, cores, orc_f, last = time_oracle(a, wl, spec, f, a.cpu_sample, 2, 1)
                cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": f"{a.cpu_sample} of {P} particles, 2 steps (+1 warm-up), N_train={N}; per-particle cost "
                                 f"is independent of P"}
            else:  # N > 1: no CPU timing (contract), a smaller untimed oracle step for the parity block only
                _, _, _, orc_f, last = time_oracle(a, wl, spec, f, min(a.cpu_sample, 512), 1, 0, loop_ll=False)
            parity = parity_block(pf, model, orc_f, last, T)
            del f, orc_f
        if parity is not None:
            parity["digest"] = digest
            parity["digest_of"] = (f"sha256(classes_post || ancestors || ll bits) after {a.warmup} warm-up + {a.steps} timed "
                                   f"steps, seed 1234: equal across --gpus N iff the sharded run is bit-identical")
        else:
            parity = {"digest": digest}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp64": "f64", "tf32": "tf32x3 (observation GP) + f64", "f16x2": "f16x2 (observation GP) + f64"}[a.precision],
            "data": "synthetic",
            "config": {"workload": workload_name(a),
                       "l2": "inputs larger than L2 (the packed K^-1 panels streamed by every particle tile are %.2f GB)"
                             % (1e-9 * lib.gpmdm_quadform_bytes((N + 255) // 256 * 256, 0 if a.dense else 1)),
                       "resampling": "multinomial", "draws": "device Philox4x32-10", "tri": not a.dense,
                       "parallelism": f"particles sharded over {world} rank(s), factors replicated"},
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(z_pinned[0].numel() * z_pinned[0].element_size()),
                    "d2h_bytes_per_step": int(probs.numel() * probs.element_size()), "steps": ke},
            "gpu_launches": int(pf.launches_per_step * a.steps), "clocks": clk,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ctypes_probe(lib, name="gpmdm_probe_dmma_tflops", iters=20000):
    import ctypes

    tf = ctypes.c_double(0.0)
    rc = getattr(lib, name)(iters, ctypes.byref(tf))
    if rc != 0:
        raise RuntimeError(name + " failed")
    return tf.value


def measured_peak(key):
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get(key)
    except (OSError, ValueError):
        return None


def measured_traffic(N, d, cached, tri, tiles):
    """DRAM bytes per launch of the observation kernel = bytes per 64-particle tile measured by ncu on the committed code
    (profiles/traffic_obs_kernel.json, written by tools/ncu_traffic.py from an `ncu --set full` capture) x the tiles of
    this launch.  None when no capture matches this configuration."""
    path = os.path.join(ROOT, "profiles", "traffic_obs_kernel.json")
    try:
        rec = json.load(open(path))
    except (OSError, ValueError):
        return None, "no ncu capture committed (profiles/traffic_obs_kernel.json missing)"
    for r in rec.get("captures", []):
        if r["N"] == N and r["d"] == d and bool(r["kstar_cache"]) == cached and bool(r["tri"]) == tri:
            return r["dram_bytes_per_tile"] * tiles, (
                f"offline ncu measurement ({r['source']}): {r['dram_bytes_per_tile'] / 1e6:.1f} MB of DRAM traffic per "
                f"64-particle tile x {tiles} tiles of this launch; {r.get('note', '')}")
    return None, "no ncu capture for this configuration in profiles/traffic_obs_kernel.json"


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    run_ours(a)


if __name__ == "__main__":
    main()
