"""GPU: the tf32 variant of the observation GP (tcgen05 + TMEM, error-compensated tf32, whitened variance) against
the fp64 exact path, at the tolerance north_star states for the fp32 variant: 1e-4 relative on means, variances and
log-weights (variances and log-likelihoods where the variance is above 5 % of the prior: below that the fp32
accumulation error of 1 - |W k|^2 is no longer 1e-4 of v; documented in DESIGN.md)."""
import numpy as np
import pytest
import torch

from gpmdm_b200 import synthetic
from tests.helpers import product_model_from_spec, synthetic_spec

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


@pytest.fixture(scope="module", params=[(2, 3, 62, 4, 60), (3, 8, 35, 5, 100), (2, 4, 300, 3, 90)])
def setup(request):
    C, d, D, spc, frames = request.param
    spec, wl = synthetic_spec(C, d, D if D <= 256 else 256, spc, frames, sigma_n=1e-1, seed=9)
    return spec, wl, product_model_from_spec(spec)


def particles(spec, P, seed, spread):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, spec.N, (P,), generator=g)
    return spec.X[idx] + spread * torch.randn(P, spec.d, dtype=torch.float64, generator=g)


@pytest.mark.parametrize("prec", ["tf32", "f16x2"])
@pytest.mark.parametrize("P", [1, 128, 129, 1000, 40000])
def test_tf32_observation_gp_matches_fp64(setup, P, prec):
    """Both tensor-core variants: 3 x tf32 and the fp16 split (a = hi + 2^-11 lo, two accumulators)."""
    spec, wl, model = setup
    xs = particles(spec, P, 3, 0.3).cuda()
    mu64, var64 = model.map_x_to_y(xs)
    mu32, var32 = model.map_x_to_y(xs, precision=prec)
    # the hybrid variant keeps the mean contraction in fp64 (same kernel, alpha tile only); the fp64 call above runs in
    # low-latency mode at these sizes (k range split over the SMs), so only the summation order over k differs
    assert float(torch.max(torch.abs(mu32 - mu64) / torch.clamp(mu64.abs().max(dim=1, keepdim=True).values, min=1e-2))) < 1e-12
    mu64f, _ = model.map_x_to_y(xs, low_latency=False)
    assert torch.equal(mu32, mu64f)
    lam = (torch.exp(model.y_log_lambdas.detach()) ** -2).unsqueeze(0)
    v64, v32 = var64 / lam, var32 / lam
    assert float(torch.max(torch.abs(v32 - v64))) < TOL32          # 1e-4 of the prior variance (= 1)
    ok = v64[:, 0] > 0.05
    if bool(ok.any()):
        assert float(torch.max(torch.abs(v32[ok] - v64[ok]) / v64[ok])) < 1e-3
    if prec != "tf32":
        return
    # the all-tensor-core mode: means from tf32 x3 products too (alpha = K^-1 Y cancels, so only ~1e-3 of the row scale)
    mup, varp = model.map_x_to_y(xs, precision="tf32-pure")
    scale = torch.clamp(mu64.abs().max(dim=1, keepdim=True).values, min=1e-2)
    assert float(torch.max(torch.abs(mup - mu64) / scale)) < 2e-3
    assert torch.equal(varp, var32)


@pytest.mark.parametrize("prec", ["tf32", "f16x2"])
@pytest.mark.parametrize("P", [1, 127, 128, 300, 20000])
def test_tensor_core_dynamics_variance_matches_fp64(setup, P, prec):
    """The class-block dynamics variance with its O(N_c^2) part on tcgen05 (gpmdm_pf_dynvar_tc) against the fp64 kernel,
    every class; the means come from the same fp64 alpha contraction in both."""
    spec, wl, model = setup
    lam = (torch.exp(model.x_log_lambdas.detach()) ** -2).unsqueeze(0)
    for c in range(spec.n_classes):
        xs = particles(spec, P, 11 + c, 0.3).cuda()
        mu64, var64 = model.map_x_dynamics_for_class(xs, c, low_latency=False, kstar_cache=False)
        mu32, var32 = model.map_x_dynamics_for_class(xs, c, precision=prec)
        assert torch.equal(mu32, mu64)
        v64, v32 = (var64 / lam)[:, 0], (var32 / lam)[:, 0]
        # the tensor cores see the RBF part only (entries <= 1): the error is 1e-4 of ONE, not of the prior
        # 1 + [x,1] diag(c^2) [x,1]^T -- the linear-kernel part of the variance is low rank and finished in fp64
        assert float(torch.max(torch.abs(v32 - v64))) < TOL32, (c, float(torch.max(torch.abs(v32 - v64))))
        assert torch.equal(var32 / lam, (var32 / lam)[:, :1].expand(-1, spec.d))


@pytest.mark.parametrize("prec", ["tf32", "f16x2"])
def test_tensor_core_dynamics_in_the_filter(setup, prec):
    """Above the low-latency size the variant's filter takes the dynamics variance from the tensor cores too: classes
    identical to the fp64 filter (the transition does not depend on it), states within the variance tolerance, no faults."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, model = setup
    T = synthetic.markov_matrix(spec.n_classes)
    P = 16384
    pf64 = GPMDM_PF(model, T, P, seed=5)
    pf32 = GPMDM_PF(model, T, P, seed=5, precision=prec)
    assert pf32._tc_dyn is not None
    z = wl.test_trials[0][1][0]
    pf64.update(z)
    pf32.update(z)
    assert torch.equal(pf64.last_pre_resample_classes, pf32.last_pre_resample_classes)
    x64, x32 = pf64.last_pre_resample_states, pf32.last_pre_resample_states
    assert bool(torch.isfinite(x32).all())
    # x = mu + sqrt(v / lambda^2) eps with |dv| <= 1e-4: |dx| <= |eps| dv / (2 sqrt(v)); v >= ~sigma_n^2 here
    assert float(torch.max(torch.abs(x32 - x64))) < 5e-3, float(torch.max(torch.abs(x32 - x64)))
    assert float(torch.median(torch.abs(x32 - x64))) < 1e-4, float(torch.median(torch.abs(x32 - x64)))
    assert pf32.variance_faults() == (0, 0)
    for _ in range(3):
        pf64.update(z)
        pf32.update(z)
    assert abs(float(pf64.class_probabilities().max()) - float(pf32.class_probabilities().max())) < 0.1


@pytest.mark.parametrize("prec", ["tf32", "f16x2"])
def test_tf32_filter_step_log_weights(setup, prec):
    from gpmdm_b200 import GPMDM_PF

    spec, wl, model = setup
    C = spec.n_classes
    T = synthetic.markov_matrix(C)
    P = 2048
    pf64 = GPMDM_PF(model, T, P, seed=5)
    pf32 = GPMDM_PF(model, T, P, seed=5, precision=prec)
    z = wl.test_trials[0][1][0]
    pf64.update(z)
    pf32.update(z)
    # same seed => identical classes and (fp64) dynamics draws; only the observation stage differs
    assert torch.equal(pf64.last_pre_resample_classes, pf32.last_pre_resample_classes)
    assert torch.equal(pf64.last_pre_resample_states, pf32.last_pre_resample_states)
    mu, var = model.map_x_to_y(pf64.last_pre_resample_states)
    v = var[:, 0] * torch.exp(model.y_log_lambdas[0]) ** 2
    ok = v > 0.05
    ll64, ll32 = pf64._log_likelihoods, pf32._log_likelihoods
    assert bool(torch.isfinite(ll32).all())
    rel = torch.abs(ll32[ok] - ll64[ok]) / torch.abs(ll64[ok])
    assert float(rel.max()) < 1e-3, float(rel.max())
    assert float(rel.median()) < TOL32, float(rel.median())
    assert pf64.get_most_likely_class() == pf32.get_most_likely_class()


def test_float32_model_dtype_argument(tmp_path):
    """The reference ctor's `dtype` argument (gpmdm.py:108-109): a float32 model keeps float32 parameters / latents and
    returns float32 predictions; the filter defaults to the tf32 variant.  Against a float64 model holding the same
    (up-cast) parameters: north_star's fp32 tolerance 1e-4 on means, variances (of the prior) and log-weights."""
    from gpmdm_b200 import GPMDM, GPMDM_PF

    C, d, D = 2, 3, 20
    wl = synthetic.make_sequences(C, D, 4, 80, seed=3, n_test_trials=1, test_frames=6)
    hp = synthetic.notebook_hyperparameters(D, d, 1e-1)
    m32 = GPMDM(D=D, d=d, n_classes=C, dyn_target="full", dyn_back_step=1, dtype=torch.float32, **hp)
    for c in range(C):
        for s in wl.sequences[c]:
            m32.add_data(s, c)
    m32.init_X()
    losses = m32.train_adam(3, 0, lr=0.01)
    assert len(losses) == 3 and all(np.isfinite(losses))
    assert m32.X.dtype == torch.float32 and m32.y_log_lambdas.dtype == torch.float32 and m32.X.grad.dtype == torch.float32
    # the same model in float64
    m64 = GPMDM(D=D, d=d, n_classes=C, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(C):
        for s in wl.sequences[c]:
            m64.add_data(s, c)
    m64.init_X()
    m64.load_state_dict({k: v.double() for k, v in m32.state_dict().items()})
    m64._precompute_kernel_inverses()
    m32.set_evaluation_mode()
    xs = (m64.X.detach()[::3] + 0.2).contiguous()
    mu32, var32 = m32.map_x_to_y(xs.float(), precision="tf32")
    mu64, var64 = m64.map_x_to_y(xs)
    assert mu32.dtype == torch.float32 and var32.dtype == torch.float32
    scale = torch.clamp(mu64.abs().max(dim=1, keepdim=True).values, min=1e-2)
    assert float(((mu32.double() - mu64).abs() / scale).max()) < TOL32
    assert float((var32.double() - var64).abs().max()) < TOL32  # lambda = exp(trained) ~ 1: of the prior variance
    dm32, dv32 = m32.map_x_dynamics_for_class(xs.float(), 1)
    dm64, dv64 = m64.map_x_dynamics_for_class(xs, 1)
    assert dm32.dtype == torch.float32 and float((dm32.double() - dm64).abs().max()) < TOL32 * (1 + float(dm64.abs().max()))
    T = synthetic.markov_matrix(C)
    pf32, pf64 = GPMDM_PF(m32, T, 512, seed=2), GPMDM_PF(m64, T, 512, seed=2, precision="tf32")
    assert pf32._precision == "tf32" and pf32.dtype == torch.float32
    z = wl.test_trials[0][1][0]
    pf32.update(z)
    pf64.update(z)
    assert torch.equal(pf32.last_pre_resample_classes, pf64.last_pre_resample_classes)
    ok = torch.isfinite(pf64._log_likelihoods)
    rel = (pf32._log_likelihoods[ok] - pf64._log_likelihoods[ok]).abs() / pf64._log_likelihoods[ok].abs()
    assert float(rel.median()) < TOL32
    assert pf32.class_probabilities().dtype == torch.float64 and pf32.current_state_mean().dtype == torch.float32
    path = str(tmp_path / "m32.pth")
    m32.save(path)
    back = GPMDM.load(path)
    assert back.dtype == torch.float32 and back.X.dtype == torch.float32
    mu_b, _ = back.map_x_to_y(xs.float())
    assert float((mu_b - m32.map_x_to_y(xs.float())[0]).abs().max()) == 0.0


@pytest.mark.parametrize("prec", ["tf32", "f16x2"])
@pytest.mark.parametrize("P", [512, 513, 640, 40000])
def test_cta_pair_variant_equals_single_cta_variant(setup, P, prec, monkeypatch):
    """cta_group::2 (clusters of two CTAs sharing every W tile, odd tile counts leave the last pair half empty) against the
    single-CTA kernel (GPMDM_TC_CLUSTER=0): the same products accumulated in the same order -> identical variances."""
    spec, wl, model = setup
    xs = particles(spec, P, 7, 0.3).cuda()
    monkeypatch.setenv("GPMDM_TC_CLUSTER", "1")
    _, var2 = model.map_x_to_y(xs, precision=prec)
    monkeypatch.setenv("GPMDM_TC_CLUSTER", "0")
    _, var1 = model.map_x_to_y(xs, precision=prec)
    assert torch.equal(var1, var2)
