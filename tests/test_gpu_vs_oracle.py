"""GPU: the CUDA path against the CPU oracle on seeded synthetic workloads of BASELINE config sizes the
oracle finishes in seconds (N = 2000; P = 100 and P = 3000), plus size-independent properties."""
import ctypes

import numpy as np
import pytest
import torch

from gpmdm_b200 import synthetic
from oracle import gpmdm_oracle as orc
from tests.helpers import product_model_from_spec, rel_err, scaled_err, synthetic_spec, t64

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def cfg1():
    """BASELINE config 1 shapes: C=2, d=3, D=62, N=2000 (20 x 100 frames)."""
    spec, wl = synthetic_spec(2, 3, 62, 10, 100, sigma_n=1e-1, seed=3)
    f = orc.precompute_factors(spec)
    model = product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=f.Kx_inv_blocks)
    return spec, wl, f, model


def particles_near_data(spec, P, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, spec.N, (P,), generator=g)
    return spec.X[idx] + 0.05 * torch.randn(P, spec.d, dtype=torch.float64, generator=g)


@pytest.mark.parametrize("low_latency", [False, True])
@pytest.mark.parametrize("P", [1, 64, 65, 100, 3000])
def test_observation_gp_vs_oracle(cfg1, P, low_latency):
    spec, wl, f, model = cfg1
    xs = particles_near_data(spec, P, 5)
    mu, var = model.map_x_to_y(xs.cuda(), low_latency=low_latency)
    mu_o, var_o, v_o = orc.map_x_to_y(spec, f, xs)
    scale = torch.clamp(torch.abs(mu_o).max(dim=1, keepdim=True).values, min=1e-3)
    assert scaled_err(mu.cpu(), mu_o, scale) < TOL
    lam = (torch.exp(spec.y_log_lambdas) ** -2).unsqueeze(0).expand(var_o.shape)
    assert scaled_err(var.cpu(), var_o, lam) < TOL


@pytest.mark.parametrize("P", [1, 65, 3000, 25000])
def test_kstar_cache_is_bit_identical_to_on_the_fly_evaluation(cfg1, P):
    """The cached instantiation of the observation kernel (K* of a particle tile evaluated once into the per-SM scratch)
    against the one that re-evaluates K* per column tile: same values, bit for bit -- also over several rounds of
    particle tiles per CTA (P = 25 000: 391 tiles on 148 SMs), where every scratch slice is overwritten and re-read."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, f, model = cfg1
    xs = particles_near_data(spec, P, 9).cuda()
    mu_c, var_c = model.map_x_to_y(xs, low_latency=False, kstar_cache=True)
    mu_f, var_f = model.map_x_to_y(xs, low_latency=False, kstar_cache=False)
    assert torch.equal(mu_c, mu_f) and torch.equal(var_c, var_f)
    if P >= 3000:
        T = synthetic.markov_matrix(spec.n_classes)
        lls = []
        for cache in (True, False):
            pf = GPMDM_PF(model, T, P, seed=4, low_latency=False, kstar_cache=cache)
            assert pf._kstar_cache == cache
            for z in wl.test_trials[0][1][:2]:
                pf.update(z)
            lls.append((pf._log_likelihoods.clone(), pf._particle_states.clone(), pf.last_ancestors.clone()))
        assert all(torch.equal(a, b) for a, b in zip(*lls))


@pytest.mark.parametrize("low_latency", [False, True])
def test_row_results_do_not_depend_on_the_batch(cfg1, low_latency):
    """A particle's prediction is bit-identical whatever else is in the batch (tile hand-out, work-item decomposition
    and k segmentation depend on the model size only) -- the property that makes a sharded run equal the 1-GPU run."""
    spec, wl, f, model = cfg1
    xs = particles_near_data(spec, 700, 17).cuda()
    mu_a, var_a = model.map_x_to_y(xs, low_latency=low_latency)
    mu_b, var_b = model.map_x_to_y(xs[130:235].contiguous(), low_latency=low_latency)
    assert torch.equal(mu_a[130:235], mu_b) and torch.equal(var_a[130:235], var_b)
    m_a, v_a = model.map_x_dynamics_for_class(xs, 1, low_latency=low_latency)
    m_b, v_b = model.map_x_dynamics_for_class(xs[3:40].contiguous(), 1, low_latency=low_latency)
    assert torch.equal(m_a[3:40], m_b) and torch.equal(v_a[3:40], v_b)


@pytest.mark.parametrize("low_latency", [False, True])
@pytest.mark.parametrize("P", [1, 100, 1000])
def test_dynamics_gp_vs_oracle(cfg1, P, low_latency):
    spec, wl, f, model = cfg1
    xs = particles_near_data(spec, P, 6)
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    for c in range(spec.n_classes):
        mean, var = model.map_x_dynamics_for_class(xs.cuda(), c, low_latency=low_latency)
        mean_o, var_o, q, prior = orc.map_x_dynamics_for_class(spec, f, xs, c)
        scale = torch.clamp(torch.abs(mean_o).max(dim=1, keepdim=True).values, min=1e-3)
        assert scaled_err(mean.cpu(), mean_o, scale) < TOL
        assert scaled_err(var.cpu(), var_o, prior.unsqueeze(1) * lam_x.unsqueeze(0)) < TOL


def test_variance_error_vs_extended_precision_arbiter(cfg1):
    """The 1 - k^T K^-1 k cancellation limits how well ANY fp64 evaluation reproduces another.
    Against an 80-bit evaluation on the same fp64 factors: the CUDA path's variance error stays below 1e-9 of
    the prior variance and is not worse than a small multiple of the torch-CPU oracle's own error (the multiple
    depends on the summation order over k: 48 particles run in low-latency mode, whose k segments are summed
    separately; measured 4.1x)."""
    spec, wl, f, model = cfg1
    xs = particles_near_data(spec, 48, 11)
    lam_x = (torch.exp(spec.x_log_lambdas) ** -2).numpy()
    for c in range(spec.n_classes):
        mean_t, var_t, prior_t = orc.dynamics_truth_longdouble(spec, f, xs, c)
        mean_g, var_g = model.map_x_dynamics_for_class(xs.cuda(), c)
        mean_o, var_o, _, _ = orc.map_x_dynamics_for_class(spec, f, xs, c)
        scale = (prior_t[:, None] * lam_x[None, :]).astype(np.float64)
        err_g = np.max(np.abs(var_g.cpu().numpy() - var_t.astype(np.float64)) / scale)
        err_o = np.max(np.abs(var_o.numpy() - var_t.astype(np.float64)) / scale)
        assert err_g < TOL, (err_g, err_o)
        assert err_g < 6 * err_o + 1e-12, (err_g, err_o)
        mscale = np.maximum(np.abs(mean_t.astype(np.float64)).max(1, keepdims=True), 1e-3)
        assert np.max(np.abs(mean_g.cpu().numpy() - mean_t.astype(np.float64)) / mscale) < TOL
    mu_t, v_t = orc.observation_truth_longdouble(spec, f, xs)
    mu_g, var_g = model.map_x_to_y(xs.cuda())
    _, _, v_o = orc.map_x_to_y(spec, f, xs)
    lam_y = float(torch.exp(spec.y_log_lambdas[0]) ** -2)
    err_g = np.max(np.abs(var_g.cpu().numpy()[:, 0] / lam_y - v_t.astype(np.float64)))
    err_o = np.max(np.abs(v_o.numpy() - v_t.astype(np.float64)))
    assert err_g < TOL and err_g < 6 * err_o + 1e-13, (err_g, err_o)


def test_tri_and_dense_packings_agree(cfg1):
    spec, wl, f, model = cfg1
    xs = particles_near_data(spec, 300, 7).cuda()
    model._packed = None
    model.packed_models(True)
    mu_t, var_t = model.map_x_to_y(xs)
    model._packed = None
    model.packed_models(False)
    mu_d, var_d = model.map_x_to_y(xs)
    model._packed = None
    assert float(torch.max(torch.abs(mu_t - mu_d))) == 0.0  # the mean tile is identical in both
    assert float(torch.max(torch.abs(var_t - var_d))) < 1e-10


@pytest.mark.parametrize("low_latency", [False, True])
@pytest.mark.parametrize("P", [100, 2500])
def test_filter_trial_vs_oracle(cfg1, P, low_latency):
    """8-frame trial with injected draws (BASELINE config 1 uses P = 100 over 150 frames; the oracle's cost
    bounds the frame count).  Every stage of every step is checked against the oracle evaluated on the CUDA
    path's own inputs to that stage, so rounding-level differences (the dynamics variance tolerance, amplified
    by the steep likelihood) cannot cascade: classes and ancestors bit-exact, argmax class identical."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, f, model = cfg1
    C, d = spec.n_classes, spec.d
    T = synthetic.markov_matrix(C)
    parts = orc.divide_into_n_parts(P, C)
    g = torch.Generator().manual_seed(9)
    init_idx = [torch.randint(0, b - a, (parts[c],), generator=g) for c, (a, b) in enumerate(spec.class_row_ranges())]
    pf = GPMDM_PF(model, T, P, init_indices=init_idx, cdf_order="sequential", low_latency=low_latency)
    cls_true, trial = wl.test_trials[0]
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    T64 = T.to(torch.float64)
    for t in range(min(8, trial.shape[0])):
        E, eps, u = synthetic.raw_draws(P, C, d, 700 + t)
        x_prev, c_prev = pf._particle_states.cpu().clone(), pf._particle_classes.cpu().clone()
        pf.update(trial[t], draws=(E, eps, u))
        z = t64(trial[t])
        # transition
        c_new = orc.transition(c_prev, T64, E)
        assert torch.equal(pf.last_pre_resample_classes.cpu(), c_new)
        # dynamics draw: |dx'| <= |eps| * (1e-9 prior) / (2 std) + 1e-9 (1 + |x'|)
        x_o, mean_o, var_o = orc.dynamics_draw(spec, f, x_prev, c_new, eps)
        prior = orc.x_diag_kernel(spec, x_prev).unsqueeze(1) * lam_x.unsqueeze(0)
        # variance tolerance here is 4e-9 of the prior: two fp64 evaluations of prior - k^T K^-1 k (the oracle's and
        # ours) each carry ~1e-9-of-prior cancellation noise at this conditioning; the error against an 80-bit
        # arbiter is bounded at 1e-9 in test_variance_error_vs_extended_precision_arbiter
        bound = torch.abs(eps) * (4 * TOL * prior) / (2 * torch.sqrt(var_o)) + TOL * (1 + torch.abs(x_o))
        x_gpu = pf.last_pre_resample_states.cpu()
        ratio = torch.abs(x_gpu - x_o) / bound
        worst = int(torch.argmax(ratio))
        wp, wk = divmod(worst, d)
        assert float(ratio.max()) <= 1.0, (
            f"step {t} particle {wp} dim {wk}: |dx|={float(torch.abs(x_gpu - x_o)[wp, wk]):.3e} bound={float(bound[wp, wk]):.3e} "
            f"var_o={float(var_o[wp, wk]):.3e} prior={float(prior[wp, wk]):.3e} eps={float(eps[wp, wk]):.3f} "
            f"mean_o={float(mean_o[wp, wk]):.6f} class={int(c_new[wp])}")
        # observation likelihood on the CUDA path's own x'
        mu_o, _, v_o = orc.map_x_to_y(spec, f, x_gpu)
        ll_o = orc.log_likelihoods_fused(mu_o, v_o, z, spec.y_log_lambdas)
        ll_gpu = pf._log_likelihoods.cpu()
        assert rel_err(ll_gpu, ll_o) < 1e-6  # v >= sigma_n^2-ish here; v itself is matched to 1e-9 absolute
        # weights, ancestors, gathers from the CUDA path's own ll
        lw_o, w_o = orc.normalize(ll_gpu)
        assert torch.equal(pf._log_weights.cpu(), lw_o)
        assert rel_err(pf._weights.cpu(), w_o) < 1e-13
        anc_o = orc.resample(pf._weights.cpu(), u)
        assert torch.equal(pf.last_ancestors.cpu(), anc_o)
        assert torch.equal(pf._particle_states.cpu(), x_gpu[anc_o])
        assert torch.equal(pf._particle_classes.cpu(), c_new[anc_o])
        # queries
        cp_o = orc.class_probabilities(ll_gpu, lw_o, c_new[anc_o], C)
        assert rel_err(pf.class_probabilities().cpu(), cp_o) < 1e-12
        assert pf.get_most_likely_class() == int(torch.argmax(cp_o))
        sm_o = orc.current_state_mean(x_gpu[anc_o], pf._weights.cpu())
        assert float(torch.max(torch.abs(pf.current_state_mean().cpu() - sm_o))) < 1e-12
        assert abs(pf.log_likelihood() - float(orc.weighted_log_sum(ll_gpu, lw_o))) < 1e-12 * abs(pf.log_likelihood())


@pytest.mark.parametrize("tri", [0, 1])
def test_factor_panels_hold_the_quadratic_form_matrix(tri):
    """gpmdm_pack_quadform_f64 / gpmdm_pack_alpha_f64 against a numpy construction of the column-panel layout
    (include/gpmdm_b200.h: gpmdm_gp_block): panel t = columns [256 t, 256 t + 256) of rows >= 256 t (tri) / all rows,
    row pitch 260, zero padded."""
    from gpmdm_b200 import _cabi
    from gpmdm_b200._cabi import check, ptr, stream

    lib = _cabi.lib()
    n, n_pad, dout, ald = 600, 768, 300, 512
    g = torch.Generator().manual_seed(1)
    Kinv = torch.randn(n, n, dtype=torch.float64, generator=g)
    Q = Kinv.clone() if not tri else torch.tril(Kinv, -1) + torch.tril(Kinv.t(), -1) + torch.diag(torch.diagonal(Kinv))
    Qp = torch.zeros(n_pad, n_pad, dtype=torch.float64)
    Qp[:n, :n] = Q
    want = []
    for t in range(n_pad // 256):
        rows = Qp[(256 * t if tri else 0):, 256 * t:256 * t + 256]
        want.append(torch.cat([rows, torch.zeros(rows.shape[0], 4, dtype=torch.float64)], 1))
    want = torch.cat(want, 0).reshape(-1)
    assert int(lib.gpmdm_quadform_bytes(n_pad, tri)) == want.numel() * 8
    L = torch.full((want.numel(),), float("nan"), dtype=torch.float64, device="cuda")
    check(lib.gpmdm_pack_quadform_f64(ptr(Kinv.cuda()), n, n_pad, tri, ptr(L), stream()), "pack")
    assert torch.equal(L.cpu(), want)
    A = torch.randn(n, dout, dtype=torch.float64, generator=g)
    Ap = torch.zeros(ald // 256, n_pad, 260, dtype=torch.float64)
    for t in range(ald // 256):
        w = min(256, dout - 256 * t)
        Ap[t, :n, :w] = A[:, 256 * t:256 * t + w]
    assert int(lib.gpmdm_alpha_bytes(n_pad, ald)) == Ap.numel() * 8
    out = torch.full((Ap.numel(),), float("nan"), dtype=torch.float64, device="cuda")
    check(lib.gpmdm_pack_alpha_f64(ptr(A.cuda()), n, n_pad, dout, ald, ptr(out), stream()), "pack alpha")
    assert torch.equal(out.cpu(), Ap.reshape(-1))


@pytest.mark.parametrize("P,kw", [(100, {}), (3000, {}), (20000, {}), (300, dict(resampling="systematic", cdf_order="blocked"))])
def test_native_step_equals_the_stage_by_stage_sequence(cfg1, P, kw):
    """gpmdm_pf_step_local_f64 / _global_f64 (csrc/pf_step.cu) against the same launches issued stage by stage from
    Python, in low-latency, fused and cached modes, with device draws and with injected draws: identical state."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, f, model = cfg1
    T = synthetic.markov_matrix(spec.n_classes)
    g = torch.Generator().manual_seed(21)
    inj = (-torch.log1p(-torch.rand(P, spec.n_classes, dtype=torch.float64, generator=g)),
           torch.randn(P, spec.d, dtype=torch.float64, generator=g), torch.rand(P, dtype=torch.float64, generator=g))
    outs = []
    for native in (True, False):
        pf = GPMDM_PF(model, T, P, seed=8, native_step=native, **kw)
        assert pf._native_step == native
        for t, z in enumerate(wl.test_trials[0][1][:3]):
            pf.update(z, draws=inj if t == 1 else None)
        outs.append((pf._particle_states.clone(), pf._particle_classes.clone(), pf._log_likelihoods.clone(),
                     pf._log_weights.clone(), pf._weights.clone(), pf.last_ancestors.clone(), pf.class_probabilities()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))


@pytest.mark.parametrize("kind", ["uniform", "peaked", "sparse", "one_hot"])
@pytest.mark.parametrize("P,n_out", [(5000, 5000), (70001, 70001), (3000, 9000)])
def test_windowed_resample_for_ascending_u_equals_the_generic_search(kind, P, n_out):
    """gpmdm_pf_resample_sorted_f64 (systematic comb: one staged cdf window per 1024 outputs) against
    gpmdm_pf_resample_f64 and the oracle's definition, on weight profiles that make the windows tiny (a few heavy
    particles), wide beyond the staging buffer (long runs of zero weight) or degenerate (one particle has it all)."""
    from gpmdm_b200 import _cabi
    from gpmdm_b200._cabi import check, ptr, stream

    lib = _cabi.lib()
    g = torch.Generator().manual_seed(P + len(kind))
    w = torch.rand(P, dtype=torch.float64, generator=g)
    if kind == "peaked":
        w = w ** 40
    elif kind == "sparse":
        w = w * (torch.rand(P, generator=g) < 2e-4)  # runs of ~5000 zero weights
        w[P // 2] += 1.0
    elif kind == "one_hot":
        w = torch.zeros(P, dtype=torch.float64)
        w[P // 3] = 1.0
    w = (w / w.sum()).cuda()
    ws = torch.empty(int(lib.gpmdm_workspace_bytes(P, 4)) // 8 + 1, dtype=torch.float64, device="cuda")
    cdf = torch.empty(P, dtype=torch.float64, device="cuda")
    check(lib.gpmdm_pf_cdf_f64(ptr(w), P, 1, ptr(cdf), ptr(ws), stream()), "cdf")
    u = ((0.37 + torch.arange(n_out, dtype=torch.float64)) / n_out).cuda()
    x = torch.randn(P, 3, dtype=torch.float64, generator=g).cuda()
    c = torch.randint(0, 4, (P,), generator=g).cuda()
    out = []
    for fn in (lib.gpmdm_pf_resample_sorted_f64, lib.gpmdm_pf_resample_f64):
        anc = torch.full((n_out,), -1, dtype=torch.int64, device="cuda")
        xo = torch.empty(n_out, 3, dtype=torch.float64, device="cuda")
        co = torch.empty(n_out, dtype=torch.int64, device="cuda")
        check(fn(ptr(cdf), P, ptr(u), n_out, ptr(x), ptr(c), 3, ptr(anc), ptr(xo), ptr(co), stream()), "resample")
        out.append((anc, xo, co))
    assert all(torch.equal(a, b) for a, b in zip(*out))
    anc = out[0][0]
    assert torch.equal(anc.cpu(), torch.clamp(torch.searchsorted(cdf.cpu(), u.cpu(), right=False), max=P - 1))
    assert torch.equal(out[0][1], x[anc]) and torch.equal(out[0][2], c[anc])


def test_bucket_by_class_is_a_stable_partition():
    from gpmdm_b200 import _cabi

    lib = _cabi.lib()
    for P, C in ((1, 1), (100, 2), (5000, 8), (70000, 64)):
        g = torch.Generator().manual_seed(P)
        cls = torch.randint(0, C, (P,), generator=g)
        if C > 2:
            cls[cls == 1] = 0  # an empty class
        cls_d = cls.cuda()
        perm = torch.empty(P, dtype=torch.int32, device="cuda")
        tiles = torch.zeros(P // 64 + C + 1, 4, dtype=torch.int32, device="cuda")
        nt = torch.zeros(1, dtype=torch.int32, device="cuda")
        ws = torch.empty(int(lib.gpmdm_workspace_bytes(P, C)) // 8 + 1, dtype=torch.float64, device="cuda")
        _cabi.check(lib.gpmdm_pf_bucket_by_class(cls_d.data_ptr(), P, C, perm.data_ptr(), tiles.data_ptr(),
                                                 nt.data_ptr(), ws.data_ptr(), _cabi.stream()), "bucket")
        expect = torch.sort(cls, stable=True).indices.to(torch.int32)
        assert torch.equal(perm.cpu(), expect)
        n = int(nt.item())
        tl = tiles.cpu()[:n]
        covered = 0
        for blk, first, count, _ in tl.tolist():
            assert 1 <= count <= 64 and first == covered
            assert bool((cls[expect[first:first + count].long()] == blk).all())
            covered += count
        assert covered == P


def test_philox_draws_are_shard_independent_and_sane():
    from gpmdm_b200 import _cabi

    lib = _cabi.lib()
    P, C, d = 200000, 3, 3

    def draw(first, n, systematic=0, step=4):
        E = torch.empty(n, C, dtype=torch.float64, device="cuda")
        eps = torch.empty(n, d, dtype=torch.float64, device="cuda")
        u = torch.empty(n, dtype=torch.float64, device="cuda")
        _cabi.check(lib.gpmdm_pf_draws_philox(1234, step, first, n, P, C, d, systematic, E.data_ptr(), eps.data_ptr(),
                                              u.data_ptr(), _cabi.stream()), "philox")
        return E.cpu(), eps.cpu(), u.cpu()

    E, eps, u = draw(0, P)
    E2, eps2, u2 = draw(P // 2, P // 2)
    assert torch.equal(E[P // 2:], E2) and torch.equal(eps[P // 2:], eps2) and torch.equal(u[P // 2:], u2)
    assert abs(float(E.mean()) - 1.0) < 0.01 and abs(float(E.var()) - 1.0) < 0.03
    assert abs(float(eps.mean())) < 0.01 and abs(float(eps.var()) - 1.0) < 0.01
    assert abs(float(u.mean()) - 0.5) < 0.005 and float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float((eps[:, 0] * eps[:, 1]).mean())) < 0.01
    E3, _, _ = draw(0, P, step=5)
    assert not torch.equal(E, E3)
    _, _, us = draw(0, P, systematic=1)
    diffs = us[1:] - us[:-1]
    assert float(torch.max(torch.abs(diffs - 1.0 / P))) < 1e-12 and 0 <= float(us[0]) < 1.0 / P


def test_filter_runs_with_device_draws_and_systematic_resampling(cfg1):
    from gpmdm_b200 import GPMDM_PF

    spec, wl, f, model = cfg1
    T = synthetic.markov_matrix(2)
    for kw in (dict(), dict(resampling="systematic", cdf_order="blocked")):
        pf = GPMDM_PF(model, T, 1000, seed=5, **kw)
        hits = 0
        cls_true, trial = wl.test_trials[1]
        for t in range(trial.shape[0]):
            pf.update(trial[t])
            hits += int(pf.get_most_likely_class() == cls_true)
        p = pf.class_probabilities()
        assert abs(float(p.sum()) - 1.0) < 1e-12 and bool(torch.isfinite(pf._particle_states).all())
        assert hits >= trial.shape[0] - 2  # the synthetic classes are well separated


def test_constructor_errors_match_reference():
    from gpmdm_b200 import GPMDM, GPMDM_PF

    spec, wl = synthetic_spec(2, 3, 10, 2, 20, seed=1)
    model = product_model_from_spec(spec)
    with pytest.raises(ValueError):
        GPMDM_PF(model, synthetic.markov_matrix(3), 10)  # gpmdm_pf.py:74-75
    with pytest.raises(ValueError):
        model.add_data(np.zeros((5, 11), dtype=np.float32), 0)  # gpmdm.py:295-296


def test_cfg2_100k_particles_sample_vs_oracle(cfg1):
    """BASELINE config 2: 2-class model, N = 2000, P = 100 000 on one B200, fp64 exact path.  The oracle is
    evaluated on a 1500-particle sample; batch-size independence makes that a statement about all rows."""
    spec, wl, f, model = cfg1
    P = 100_000
    xs = particles_near_data(spec, P, 12)
    xs_d = xs.cuda()
    mu, var = model.map_x_to_y(xs_d)
    sample = torch.randperm(P, generator=torch.Generator().manual_seed(1))[:1500]
    mu_s, var_s = model.map_x_to_y(xs_d[sample.cuda()].contiguous(), low_latency=False)
    assert torch.equal(mu[sample.cuda()], mu_s) and torch.equal(var[sample.cuda()], var_s)  # row results are batch independent
    mu_l, var_l = model.map_x_to_y(xs_d[sample.cuda()].contiguous(), low_latency=True)  # other summation order over k
    assert float(torch.max(torch.abs(mu_l - mu_s))) < 1e-11 and float(torch.max(torch.abs(var_l - var_s))) < 1e-11
    mu_o, var_o, v_o = orc.map_x_to_y(spec, f, xs[sample])
    scale = torch.clamp(torch.abs(mu_o).max(dim=1, keepdim=True).values, min=1e-3)
    assert scaled_err(mu_s.cpu(), mu_o, scale) < TOL
    lam = (torch.exp(spec.y_log_lambdas) ** -2).unsqueeze(0).expand(var_o.shape)
    assert scaled_err(var_s.cpu(), var_o, lam) < TOL
    v = var[:, 0] / lam[0, 0].item()
    assert bool(torch.all(v > 0)) and bool(torch.all(v <= 1 + 1e-12))
    # a full filter step at P = 100k with injected draws: transition / ancestors exact w.r.t. the oracle stages
    from gpmdm_b200 import GPMDM_PF
    C, d = spec.n_classes, spec.d
    T = synthetic.markov_matrix(C)
    pf = GPMDM_PF(model, T, P, seed=3, cdf_order="sequential")
    E, eps, u = synthetic.raw_draws(P, C, d, 77)
    c_prev = pf._particle_classes.cpu().clone()
    pf.update(wl.test_trials[0][1][0], draws=(E, eps, u))
    assert torch.equal(pf.last_pre_resample_classes.cpu(), orc.transition(c_prev, T.to(torch.float64), E))
    lw_o, w_o = orc.normalize(pf._log_likelihoods.cpu())
    assert torch.equal(pf._log_weights.cpu(), lw_o)
    assert torch.equal(pf.last_ancestors.cpu(), orc.resample(pf._weights.cpu(), u))
    assert abs(float(pf._weights.sum()) - 1.0) < 1e-12


def test_cfg3_scale_properties():
    """BASELINE config 3 factor sizes (C = 8, N = 20 000, D = 62), where the oracle's dense objects are out of
    reach: size-independent properties -- the triangular and dense packings are two independent evaluations of
    k^T K^-1 k and must agree; variances lie in (0, prior]; far from the data the prediction is the prior;
    weights sum to one; systematic-resampling ancestors are sorted; the step is reproducible."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl = synthetic_spec(8, 3, 62, 25, 100, sigma_n=1e-1, seed=0)
    model = product_model_from_spec(spec)
    g = torch.Generator().manual_seed(2)
    idx = torch.randint(0, spec.N, (1024,), generator=g)
    xs = (spec.X[idx] + 0.02 * torch.randn(1024, 3, dtype=torch.float64, generator=g)).cuda()
    model._packed = None
    model.packed_models(True)
    mu_t, var_t = model.map_x_to_y(xs)
    m_t, v_t = model.map_x_dynamics_for_class(xs, 3)
    model._packed = None
    model.packed_models(False)
    mu_d, var_d = model.map_x_to_y(xs)
    m_d, v_d = model.map_x_dynamics_for_class(xs, 3)
    model._packed = None
    # the kernel instance bench.py measures (fused, K* cache, 79 column panels) against the on-the-fly instance and the
    # low-latency decomposition at this size
    model.packed_models(True)
    mu_c, var_c = model.map_x_to_y(xs[:300], low_latency=False, kstar_cache=True)
    mu_f, var_f = model.map_x_to_y(xs[:300], low_latency=False, kstar_cache=False)
    assert torch.equal(mu_c, mu_f) and torch.equal(var_c, var_f)
    assert float(torch.max(torch.abs(mu_c - mu_t[:300]))) < 1e-10 and float(torch.max(torch.abs(var_c - var_t[:300]))) < 1e-10
    model._packed = None
    assert float(torch.max(torch.abs(mu_t - mu_d))) == 0.0 and float(torch.max(torch.abs(m_t - m_d))) == 0.0
    assert float(torch.max(torch.abs(var_t - var_d))) < 1e-9
    prior = 1 + (xs ** 2).sum(1) + 1  # all-ones linear coefficients
    assert float(torch.max(torch.abs(v_t - v_d) / prior.unsqueeze(1))) < 1e-9
    assert bool(torch.all(var_t > 0)) and bool(torch.all(var_t <= 1 + 1e-9))
    # far from all training data the prediction reverts to the prior: mean 0, variance 1 (RBF) / prior (dynamics)
    far = torch.full((64, 3), 1.0e3, dtype=torch.float64, device="cuda")
    mu_f, var_f = model.map_x_to_y(far)
    assert float(torch.max(torch.abs(mu_f))) == 0.0 and float(torch.max(torch.abs(var_f - 1.0))) == 0.0
    T = synthetic.markov_matrix(8)
    outs = []
    for rep in range(2):
        pf = GPMDM_PF(model, T, 4096, seed=11, resampling="systematic", cdf_order="blocked")
        pf.update(wl.test_trials[0][1][0])
        outs.append((pf._particle_states.clone(), pf.last_ancestors.clone(), pf.class_probabilities()))
        assert abs(float(pf._weights.sum()) - 1.0) < 1e-12
        anc = pf.last_ancestors
        assert bool(torch.all(anc[1:] >= anc[:-1]))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_class_agnostic_dynamics_map_and_helpers():
    """`map_x_dynamics` (gpmdm.py:993-1030, not on the filter path): the sum of per-class fused predictions equals the
    reference's dense expression with K_x o M; plus the notebook helper methods."""
    spec, wl = synthetic_spec(3, 3, 10, 2, 30, sigma_n=1e-1, seed=7)
    model = product_model_from_spec(spec)
    xs = particles_near_data(spec, 70, 4)
    mean, var = model.map_x_dynamics(xs.cuda())
    Xin, Xout = orc.xin_xout(spec)
    Kinv = orc.inverse_via_upper_cholesky(orc.x_kernel(spec, Xin, Xin) * orc.class_mask(spec))  # gpmdm.py:1292-1295
    Ks = orc.x_kernel(spec, Xin, xs, False)
    mean_o = torch.linalg.multi_dot([Xout.t(), Kinv, Ks]).t()
    prior = orc.x_diag_kernel(spec, xs)
    common = prior - torch.sum(torch.matmul(Ks.t(), Kinv) * Ks.t(), dim=1)
    lam = torch.exp(spec.x_log_lambdas) ** -2
    var_o = common.unsqueeze(1) * lam.unsqueeze(0)
    scale = torch.clamp(mean_o.abs().max(dim=1, keepdim=True).values, min=1e-3)
    assert scaled_err(mean.cpu(), mean_o, scale) < 1e-7
    assert scaled_err(var.cpu(), var_o, (prior.unsqueeze(1) * lam.unsqueeze(0))) < 1e-7
    # (the masked-kernel formula can go negative away from a class' data, in the reference as well: clamp for Normal)
    nxt = model.get_next_x(mean, torch.clamp(var, min=1e-12), xs.cuda())
    assert torch.equal(nxt, mean)
    m1, v1, Xo, Xi, nmse = model.get_dynamics_map_performance_for_class(1)
    assert m1.shape == Xo.shape and np.isfinite(nmse)
    mu, vy, Y, nmse_y = model.get_latent_map_performance()
    assert mu.shape == Y.shape and np.isfinite(nmse_y)


@pytest.mark.parametrize("C,d,D", [(2, 8, 35), (2, 1, 5), (3, 2, 300), (2, 5, 257)])
def test_other_latent_and_observation_dimensions_vs_oracle(C, d, D):
    """Template instances beyond d = 3/4 (d = 1, 2, 5, 8) and observation widths that need more than one 256-column
    alpha tile (D = 257, 300): GP predictions and one injected-draw filter step against the oracle."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl = synthetic_spec(C, d, D, 3, 40, sigma_n=1e-1, seed=13)
    f = orc.precompute_factors(spec)
    model = product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=f.Kx_inv_blocks)
    xs = particles_near_data(spec, 200, 3)
    mu, var = model.map_x_to_y(xs.cuda())
    mu_o, var_o, v_o = orc.map_x_to_y(spec, f, xs)
    scale = torch.clamp(torch.abs(mu_o).max(dim=1, keepdim=True).values, min=1e-3)
    assert scaled_err(mu.cpu(), mu_o, scale) < TOL
    lam = (torch.exp(spec.y_log_lambdas) ** -2).unsqueeze(0).expand(var_o.shape)
    assert scaled_err(var.cpu(), var_o, lam) < TOL
    # fused mode, with and without the K* cache (these template instances are otherwise only reached in low-latency mode)
    mu_c, var_c = model.map_x_to_y(xs.cuda(), low_latency=False, kstar_cache=True)
    mu_f, var_f = model.map_x_to_y(xs.cuda(), low_latency=False, kstar_cache=False)
    assert torch.equal(mu_c, mu_f) and torch.equal(var_c, var_f)
    assert scaled_err(mu_c.cpu(), mu_o, scale) < TOL and scaled_err(var_c.cpu(), var_o, lam) < TOL
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    for c in range(C):
        mean, dvar = model.map_x_dynamics_for_class(xs.cuda(), c)
        mean_o, dvar_o, q, prior = orc.map_x_dynamics_for_class(spec, f, xs, c)
        sc = torch.clamp(torch.abs(mean_o).max(dim=1, keepdim=True).values, min=1e-3)
        assert scaled_err(mean.cpu(), mean_o, sc) < TOL
        assert scaled_err(dvar.cpu(), dvar_o, prior.unsqueeze(1) * lam_x.unsqueeze(0)) < 4 * TOL
    P = 130
    T = synthetic.markov_matrix(C)
    pf = GPMDM_PF(model, T, P, seed=2, cdf_order="sequential")
    E, eps, u = synthetic.raw_draws(P, C, d, 5)
    x_prev, c_prev = pf._particle_states.cpu().clone(), pf._particle_classes.cpu().clone()
    z = wl.test_trials[0][1][0]
    pf.update(z, draws=(E, eps, u))
    c_new = orc.transition(c_prev, T.to(torch.float64), E)
    assert torch.equal(pf.last_pre_resample_classes.cpu(), c_new)
    x_gpu = pf.last_pre_resample_states.cpu()
    mu_o, _, v_o = orc.map_x_to_y(spec, f, x_gpu)
    ll_o = orc.log_likelihoods_fused(mu_o, v_o, t64(z), spec.y_log_lambdas)
    ok = v_o > 1e-3
    assert rel_err(pf._log_likelihoods.cpu()[ok], ll_o[ok]) < 1e-6
    assert torch.equal(pf.last_ancestors.cpu(), orc.resample(pf._weights.cpu(), u))


def test_cfg1_readme_trial_150_frames(cfg1):
    """BASELINE config 1 exactly: 2-class model, d = 3, D = 62, N = 2000, 100 particles, a 150-frame trial with
    injected draws; every frame checked stage-wise against the oracle (classes / ancestors bit-exact, argmax class
    identical), the filter itself running free on the CUDA path."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, f, model = cfg1
    C, d, P = spec.n_classes, spec.d, 100
    trial = synthetic.make_sequences(C, spec.D, 1, 150, seed=77).sequences[1][0]
    T = synthetic.markov_matrix(C)
    T64 = T.to(torch.float64)
    parts = orc.divide_into_n_parts(P, C)
    g = torch.Generator().manual_seed(10)
    init_idx = [torch.randint(0, b - a, (parts[c],), generator=g) for c, (a, b) in enumerate(spec.class_row_ranges())]
    pf = GPMDM_PF(model, T, P, init_indices=init_idx, cdf_order="sequential")
    for t in range(150):
        E, eps, u = synthetic.raw_draws(P, C, d, 9000 + t)
        c_prev = pf._particle_classes.cpu().clone()
        pf.update(trial[t], draws=(E, eps, u))
        c_new = orc.transition(c_prev, T64, E)
        assert torch.equal(pf.last_pre_resample_classes.cpu(), c_new), t
        x_gpu = pf.last_pre_resample_states.cpu()
        mu_o, _, v_o = orc.map_x_to_y(spec, f, x_gpu)
        ll_o = orc.log_likelihoods_fused(mu_o, v_o, t64(trial[t]), spec.y_log_lambdas)
        assert rel_err(pf._log_likelihoods.cpu(), ll_o) < 1e-6, t
        lw_o, w_o = orc.normalize(pf._log_likelihoods.cpu())
        anc_o = orc.resample(pf._weights.cpu(), u)
        assert torch.equal(pf.last_ancestors.cpu(), anc_o), t
        cp_o = orc.class_probabilities(pf._log_likelihoods.cpu(), lw_o, c_new[anc_o], C)
        assert pf.get_most_likely_class() == int(torch.argmax(cp_o)), t


@pytest.mark.parametrize("C,P", [(1, 50), (3, 2), (3, 1), (2, 64), (2, 65)])
def test_degenerate_particle_and_class_counts(C, P):
    """One-class models, fewer particles than classes (some classes start empty), a single particle, and exactly one
    tile / one tile plus one: two filter steps with injected draws against the oracle stages."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl = synthetic_spec(C, 2, 7, 2, 30, sigma_n=1e-1, seed=5)
    f = orc.precompute_factors(spec)
    model = product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=f.Kx_inv_blocks)
    T = torch.eye(1, dtype=torch.float32) if C == 1 else synthetic.markov_matrix(C)
    pf = GPMDM_PF(model, T, P, seed=2, cdf_order="sequential")
    assert pf._particle_states.shape == (P, 2) and pf._particle_classes.shape == (P,)
    for t in range(2):
        E, eps, u = synthetic.raw_draws(P, C, 2, 90 + t)
        c_prev, x_prev = pf._particle_classes.cpu().clone(), pf._particle_states.cpu().clone()
        z = wl.test_trials[0][1][t]
        pf.update(z, draws=(E, eps, u))
        c_new = orc.transition(c_prev, T.to(torch.float64), E)
        assert torch.equal(pf.last_pre_resample_classes.cpu(), c_new)
        x_new = pf.last_pre_resample_states.cpu()
        for c in range(C):
            m = c_new == c
            if bool(m.any()):
                mean_o, var_o, _, prior = orc.map_x_dynamics_for_class(spec, f, x_prev[m], c)
                x_o = eps[m] * torch.sqrt(var_o) + mean_o
                assert float(torch.max(torch.abs(x_new[m] - x_o))) < 1e-8
        mu_o, _, v_o = orc.map_x_to_y(spec, f, x_new)
        ll_o = orc.log_likelihoods_fused(mu_o, v_o, torch.as_tensor(z, dtype=torch.float64), spec.y_log_lambdas)
        assert float(torch.max(torch.abs(pf._log_likelihoods.cpu() - ll_o) / torch.abs(ll_o))) < 1e-6
        lw_o, w_o = orc.normalize(pf._log_likelihoods.cpu())
        assert torch.equal(pf._log_weights.cpu(), lw_o)
        assert torch.equal(pf.last_ancestors.cpu(), orc.resample(pf._weights.cpu(), u))
        probs = pf.class_probabilities()
        assert probs.shape == (C,) and abs(float(probs.sum()) - 1.0) < 1e-12
        assert 0 <= pf.get_most_likely_class() < C


def test_unequal_class_blocks_low_latency_vs_fused():
    """Dynamics blocks of very different sizes (N_c = 39, 299, 1295: 1, 2 and 6 column panels) in one model: the
    low-latency decomposition (per-block column tiles and k segments inside a uniform item grid) against the fused
    kernel, per class and through a filter step."""
    from gpmdm_b200 import GPMDM, GPMDM_PF

    D, d = 9, 2
    wl = synthetic.make_sequences(3, D, 5, 300, seed=4, n_test_trials=1, test_frames=4)
    hp = synthetic.notebook_hyperparameters(D, d, 1e-1)
    m = GPMDM(D=D, d=d, n_classes=3, dyn_target="full", dyn_back_step=1, **hp)
    m.add_data(wl.sequences[0][0][:40], 0)
    m.add_data(wl.sequences[1][0], 1)
    for s in wl.sequences[2]:
        m.add_data(s[:260], 2)
    m.init_X()
    m.set_evaluation_mode()
    g = torch.Generator().manual_seed(3)
    X = m.X.detach().cpu()
    xs = (X[torch.randint(0, X.shape[0], (333,), generator=g)] + 0.05 * torch.randn(333, d, dtype=torch.float64, generator=g)).cuda()
    for c in range(3):
        mf, vf = m.map_x_dynamics_for_class(xs, c, low_latency=False)
        ml, vl = m.map_x_dynamics_for_class(xs, c, low_latency=True)
        prior = (2.0 + (xs ** 2).sum(1)).unsqueeze(1)  # all-ones linear coefficients; the variance contract is 1e-9 of it
        assert float(torch.max(torch.abs(mf - ml))) < 1e-10 and float(torch.max(torch.abs(vf - vl) / prior)) < 1e-9
    mu_f, var_f = m.map_x_to_y(xs, low_latency=False)
    mu_l, var_l = m.map_x_to_y(xs, low_latency=True)
    assert float(torch.max(torch.abs(mu_f - mu_l))) < 1e-10 and float(torch.max(torch.abs(var_f - var_l))) < 1e-10
    T = synthetic.markov_matrix(3)
    outs = []
    for low in (False, True):
        pf = GPMDM_PF(m, T, 500, seed=6, low_latency=low)
        pf.update(wl.test_trials[0][1][0])
        outs.append((pf.last_pre_resample_classes.clone(), pf.last_pre_resample_states.clone(), pf._log_likelihoods.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    # x' = eps * sqrt(var) + mean: where the predictive variance is ~1e-6 of the prior, the 1e-9-of-the-prior agreement of
    # two summation orders becomes ~1e-7 after the square root (the reference itself is no more reproducible there,
    # SURVEY fact 8)
    assert float(torch.max(torch.abs(outs[0][1] - outs[1][1]))) < 1e-5
    assert float(torch.max(torch.abs(outs[0][2] - outs[1][2]) / torch.abs(outs[0][2]))) < 1e-5
