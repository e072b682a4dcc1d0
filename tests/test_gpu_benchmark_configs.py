"""GPU: the CUDA path against the CPU oracle AT THE CONFIGURATIONS THE BENCHMARK NUMBERS ARE QUOTED ON.

  * BASELINE configs[2] factor sizes -- C = 8, N = 20 000, D = 62, d = 3 (the model bench.py builds): 256 particles
    near the data and one injected-draw filter step, fused kernel with the K* cache (the instance bench.py times),
    fused without it, and the low-latency decomposition, against `orc.map_x_to_y` / `log_likelihoods_fused` /
    `transition` / `dynamics_draw` / `resample` (reference gpmdm.py:923-963, 1032-1068, gpmdm_pf.py:126-213).
  * C = 64, d = 8 (the class / latent limits of configs[3]) at N = 6 400: the `<0,8,*>` / `<1,8,*>` kernel instances,
    the C = 64 transition / bucketing / summaries, fp64 against the oracle and the tf32 variant against fp64.

The oracle's O(N^3) factor recipe is evaluated with plain torch on the GPU (test setup, `orc.precompute_factors_on`);
every oracle prediction runs on the CPU.  Tolerances are those of tests/test_gpu_parity_golden.py."""
import numpy as np
import pytest
import torch

from gpmdm_b200 import synthetic
from oracle import gpmdm_oracle as orc
from tests.helpers import product_model_from_spec, rel_err, scaled_err, synthetic_spec, t64

pytestmark = pytest.mark.gpu
TOL = 1e-9


def near_data(spec, P, seed, spread=0.05):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, spec.N, (P,), generator=g)
    return spec.X[idx] + spread * torch.randn(P, spec.d, dtype=torch.float64, generator=g)


def check_step_against_oracle(spec, f, model, wl, P, seed, var_tol=4 * TOL, **pf_kw):
    """One injected-draw filter step; every stage checked against the oracle evaluated on the CUDA path's own inputs
    to that stage (same scheme and tolerances as test_gpu_vs_oracle.test_filter_trial_vs_oracle)."""
    from gpmdm_b200 import GPMDM_PF

    C, d = spec.n_classes, spec.d
    T = synthetic.markov_matrix(C)
    T64 = T.to(torch.float64)
    parts = orc.divide_into_n_parts(P, C)
    g = torch.Generator().manual_seed(seed)
    init_idx = [torch.randint(0, b - a, (parts[c],), generator=g) for c, (a, b) in enumerate(spec.class_row_ranges())]
    pf = GPMDM_PF(model, T, P, init_indices=init_idx, cdf_order="sequential", **pf_kw)
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    stats = {}
    for t in range(2):
        z_np = wl.test_trials[0][1][t]
        z = t64(z_np)
        E, eps, u = synthetic.raw_draws(P, C, d, seed * 100 + t)
        x_prev, c_prev = pf._particle_states.cpu().clone(), pf._particle_classes.cpu().clone()
        pf.update(z_np, draws=(E, eps, u))
        c_new = orc.transition(c_prev, T64, E)
        assert torch.equal(pf.last_pre_resample_classes.cpu(), c_new)
        x_o, mean_o, var_o = orc.dynamics_draw(spec, f, x_prev, c_new, eps)
        prior = orc.x_diag_kernel(spec, x_prev).unsqueeze(1) * lam_x.unsqueeze(0)
        bound = torch.abs(eps) * (var_tol * prior) / (2 * torch.sqrt(var_o)) + TOL * (1 + torch.abs(x_o))
        x_gpu = pf.last_pre_resample_states.cpu()
        assert float((torch.abs(x_gpu - x_o) / bound).max()) <= 1.0
        mu_o, _, v_o = orc.map_x_to_y(spec, f, x_gpu)
        ll_o = orc.log_likelihoods_fused(mu_o, v_o, z, spec.y_log_lambdas)
        ll_gpu = pf._log_likelihoods.cpu()
        ok = v_o > 1e-3
        assert rel_err(ll_gpu[ok], ll_o[ok]) < 1e-6
        stats["ll_rel_max"] = max(stats.get("ll_rel_max", 0.0), rel_err(ll_gpu[ok], ll_o[ok]))
        lw_o, w_o = orc.normalize(ll_gpu)
        assert torch.equal(pf._log_weights.cpu(), lw_o)
        assert rel_err(pf._weights.cpu(), w_o) < 1e-13
        anc_o = orc.resample(pf._weights.cpu(), u)
        assert torch.equal(pf.last_ancestors.cpu(), anc_o)
        assert torch.equal(pf._particle_states.cpu(), x_gpu[anc_o])
        assert torch.equal(pf._particle_classes.cpu(), c_new[anc_o])
        cp_o = orc.class_probabilities(ll_gpu, lw_o, c_new[anc_o], C)
        assert rel_err(pf.class_probabilities().cpu(), cp_o) < 1e-12
        assert pf.get_most_likely_class() == int(torch.argmax(cp_o))
        sm_o = orc.current_state_mean(x_gpu[anc_o], pf._weights.cpu())
        assert float(torch.max(torch.abs(pf.current_state_mean().cpu() - sm_o))) < 1e-12
    return stats


# ---- BASELINE configs[2]: C = 8, N = 20 000, D = 62, d = 3 ---------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg3():
    spec, wl = synthetic_spec(8, 3, 62, 25, 100, sigma_n=1e-1, seed=0)
    assert spec.N == 20000
    f = orc.precompute_factors_on(spec, "cuda")
    model = product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=f.Kx_inv_blocks, own_factors=False)
    torch.cuda.empty_cache()
    return spec, wl, f, model


def test_cfg3_observation_gp_vs_oracle(cfg3):
    spec, wl, f, model = cfg3
    xs = near_data(spec, 256, 41)
    mu_o, var_o, v_o = orc.map_x_to_y(spec, f, xs)
    scale = torch.clamp(torch.abs(mu_o).max(dim=1, keepdim=True).values, min=1e-3)
    lam = (torch.exp(spec.y_log_lambdas) ** -2).unsqueeze(0).expand(var_o.shape)
    outs = {}
    for name, kw in (("fused+K* cache", dict(low_latency=False, kstar_cache=True)),
                     ("fused", dict(low_latency=False, kstar_cache=False)),
                     ("low latency", dict(low_latency=True))):
        mu, var = model.map_x_to_y(xs.cuda(), **kw)
        outs[name] = (mu, var)
        assert scaled_err(mu.cpu(), mu_o, scale) < TOL, name
        assert scaled_err(var.cpu(), var_o, lam) < TOL, name
    assert torch.equal(outs["fused"][0], outs["fused+K* cache"][0]) and torch.equal(outs["fused"][1], outs["fused+K* cache"][1])


def test_cfg3_dynamics_gp_vs_oracle(cfg3):
    spec, wl, f, model = cfg3
    xs = near_data(spec, 256, 42)
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    for c in (0, 3, 7):
        mean_o, var_o, q, prior = orc.map_x_dynamics_for_class(spec, f, xs, c)
        sc = torch.clamp(torch.abs(mean_o).max(dim=1, keepdim=True).values, min=1e-3)
        for low in (False, True):
            mean, var = model.map_x_dynamics_for_class(xs.cuda(), c, low_latency=low)
            assert scaled_err(mean.cpu(), mean_o, sc) < TOL
            assert scaled_err(var.cpu(), var_o, prior.unsqueeze(1) * lam_x.unsqueeze(0)) < 4 * TOL
        # the K* cache instance of the dynamics kernel (10 column panels per class block here) against the on-the-fly one
        m_c, v_c = model.map_x_dynamics_for_class(xs.cuda(), c, low_latency=False, kstar_cache=True)
        m_f, v_f = model.map_x_dynamics_for_class(xs.cuda(), c, low_latency=False, kstar_cache=False)
        assert torch.equal(m_c, m_f) and torch.equal(v_c, v_f)


@pytest.mark.parametrize("mode", ["cached", "lowlat"])
def test_cfg3_filter_step_vs_oracle(cfg3, mode):
    spec, wl, f, model = cfg3
    kw = dict(low_latency=False, kstar_cache=True) if mode == "cached" else dict(low_latency=True)
    check_step_against_oracle(spec, f, model, wl, 256, 7, **kw)


# ---- C = 64, d = 8 (configs[3] limits) at N = 6 400 -----------------------------------------------------------------------
@pytest.fixture(scope="module")
def c64():
    spec, wl = synthetic_spec(64, 8, 62, 2, 50, sigma_n=1e-1, seed=4)
    assert spec.N == 6400 and spec.n_classes == 64 and spec.d == 8
    f = orc.precompute_factors_on(spec, "cuda")
    model = product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=f.Kx_inv_blocks)
    return spec, wl, f, model


def test_c64_d8_gp_predictions_vs_oracle(c64):
    spec, wl, f, model = c64
    xs = near_data(spec, 300, 43)
    mu_o, var_o, v_o = orc.map_x_to_y(spec, f, xs)
    scale = torch.clamp(torch.abs(mu_o).max(dim=1, keepdim=True).values, min=1e-3)
    lam = (torch.exp(spec.y_log_lambdas) ** -2).unsqueeze(0).expand(var_o.shape)
    for kw in (dict(low_latency=False, kstar_cache=True), dict(low_latency=False, kstar_cache=False), dict(low_latency=True)):
        mu, var = model.map_x_to_y(xs.cuda(), **kw)
        assert scaled_err(mu.cpu(), mu_o, scale) < TOL and scaled_err(var.cpu(), var_o, lam) < TOL
    lam_x = torch.exp(spec.x_log_lambdas) ** -2
    for c in (0, 31, 63):
        mean_o, dvar_o, q, prior = orc.map_x_dynamics_for_class(spec, f, xs, c)
        sc = torch.clamp(torch.abs(mean_o).max(dim=1, keepdim=True).values, min=1e-3)
        for low in (False, True):
            mean, dvar = model.map_x_dynamics_for_class(xs.cuda(), c, low_latency=low)
            assert scaled_err(mean.cpu(), mean_o, sc) < TOL
            assert scaled_err(dvar.cpu(), dvar_o, prior.unsqueeze(1) * lam_x.unsqueeze(0)) < 4 * TOL


@pytest.mark.parametrize("P,kw", [(640, dict(low_latency=True)), (8192, dict(low_latency=False, kstar_cache=True))])
def test_c64_d8_filter_step_vs_oracle(c64, P, kw):
    """C = 64 through transition_kernel / bucket_by_class (64 class-homogeneous tile groups, some classes empty after the
    transition at P = 640) / summaries (64 class sums), d = 8 through both predict instances."""
    spec, wl, f, model = c64
    check_step_against_oracle(spec, f, model, wl, P, 11, **kw)


def test_c64_d8_tf32_variant_vs_fp64(c64):
    """configs[3]'s tolerance check at reduced N (the full-size run is tools/cfg4_check.py): the tf32 variant of the
    observation GP (tcgen05, 3 x tf32, whitened variance) against the fp64 exact path, north_star's fp32 tolerance 1e-4."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, f, model = c64
    g = torch.Generator().manual_seed(3)
    xs = near_data(spec, 1024, 44, spread=0.3).cuda()
    mu64, var64 = model.map_x_to_y(xs, low_latency=False)
    mu32, var32 = model.map_x_to_y(xs, precision="tf32")
    assert torch.equal(mu32, mu64)  # hybrid variant: the mean contraction is the fp64 kernel on the alpha tile
    v64, v32 = var64[:, 0], var32[:, 0]  # lambda = 1
    assert float(torch.max(torch.abs(v32 - v64))) < 1e-4  # of the prior variance (= 1)
    ok = v64 > 0.05
    assert bool(ok.any()) and float(torch.max(torch.abs(v32[ok] - v64[ok]) / v64[ok])) < 1e-3
    T = synthetic.markov_matrix(64)
    pf64 = GPMDM_PF(model, T, 4096, seed=5)
    pf32 = GPMDM_PF(model, T, 4096, seed=5, precision="tf32")
    z = wl.test_trials[0][1][0]
    pf64.update(z)
    pf32.update(z)
    assert torch.equal(pf64.last_pre_resample_classes, pf32.last_pre_resample_classes)
    assert torch.equal(pf64.last_pre_resample_states, pf32.last_pre_resample_states)
    _, var = model.map_x_to_y(pf64.last_pre_resample_states)
    ok = var[:, 0] > 0.05
    rel = torch.abs(pf32._log_likelihoods[ok] - pf64._log_likelihoods[ok]) / torch.abs(pf64._log_likelihoods[ok])
    assert float(rel.max()) < 1e-3 and float(rel.median()) < 1e-4
    assert pf64.get_most_likely_class() == pf32.get_most_likely_class()
