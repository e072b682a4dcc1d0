"""GPU: the small-cloud step (csrc/pf_small.cu: two fused single-CTA kernels around the low-latency predict calls, replayed
from a CUDA graph) against the staged launch sequence it replaces -- identical state, bit for bit, at every step: the
fused kernels are built from the same device functions and walk the same reduction blocks in the same order."""
import numpy as np
import pytest
import torch

from gpmdm_b200 import GPMDM_PF, synthetic
from oracle import gpmdm_oracle as orc
from tests.helpers import product_model_from_spec, synthetic_spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    spec, wl = synthetic_spec(3, 3, 30, 4, 90, sigma_n=1e-1, seed=12)  # N = 1080
    f = orc.precompute_factors(spec)
    return spec, wl, f, product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=f.Kx_inv_blocks)


def state(pf):
    return (pf._particle_states.clone(), pf._particle_classes.clone(), pf._log_likelihoods.clone(), pf._log_weights.clone(),
            pf._weights.clone(), pf.last_ancestors.clone(), pf.last_pre_resample_states.clone(),
            pf.last_pre_resample_classes.clone(), pf.class_probabilities(), pf.current_state_mean(),
            torch.tensor(pf.log_likelihood()), torch.tensor(pf.get_most_likely_class()))


@pytest.mark.parametrize("kw", [dict(), dict(resampling="systematic", cdf_order="blocked"), dict(cdf_order="blocked")])
@pytest.mark.parametrize("P", [1, 100, 1000, 1024, 1025, 4096])
def test_graph_replay_direct_launch_and_staged_sequence_agree(setup, P, kw):
    spec, wl, f, model = setup
    T = synthetic.markov_matrix(spec.n_classes)
    trial = wl.test_trials[0][1]
    g = torch.Generator().manual_seed(5)
    inj = (-torch.log1p(-torch.rand(P, spec.n_classes, dtype=torch.float64, generator=g)),
           torch.randn(P, spec.d, dtype=torch.float64, generator=g), torch.rand(P, dtype=torch.float64, generator=g))
    variants = [dict(cuda_graph=True), dict(cuda_graph=False), dict(native_step=False)]
    pfs = [GPMDM_PF(model, T, P, seed=9, **kw, **v) for v in variants]
    assert pfs[0]._small and pfs[0]._use_graph and pfs[1]._small and not pfs[1]._use_graph and not pfs[2]._small
    for t in range(9):
        z = trial[t % trial.shape[0]]
        for pf in pfs:
            pf.update(z, draws=inj if t == 4 else None)  # an injected-draw step in the middle of the replayed ones
        ref = state(pfs[2])
        for pf in pfs[:2]:
            assert all(torch.equal(a, b) for a, b in zip(state(pf), ref)), (t, P)
        if t == 5:  # a cloud rebound from outside (tests do that) and a reset are adopted by the graph's own buffers
            for pf in pfs:
                pf._particle_states = pf._particle_states.flip(0).contiguous()
                pf._particle_classes = pf._particle_classes.flip(0).contiguous()
        if t == 7:
            for pf in pfs:
                pf.reset()
    assert pfs[0]._graphs[0] is not None and pfs[0]._graphs[1] is not None
    assert pfs[0].launches_per_step == 6


def test_small_cloud_step_against_the_oracle(setup):
    """The fused path end to end against the oracle stages (injected draws), like every other filter test."""
    spec, wl, f, model = setup
    C, d, P = spec.n_classes, spec.d, 100
    T = synthetic.markov_matrix(C)
    parts = orc.divide_into_n_parts(P, C)
    g = torch.Generator().manual_seed(3)
    init_idx = [torch.randint(0, b - a, (parts[c],), generator=g) for c, (a, b) in enumerate(spec.class_row_ranges())]
    pf = GPMDM_PF(model, T, P, init_indices=init_idx, cdf_order="sequential")
    assert pf._small
    for t in range(4):
        E, eps, u = synthetic.raw_draws(P, C, d, 300 + t)
        c_prev = pf._particle_classes.cpu().clone()
        z = wl.test_trials[1][1][t]
        pf.update(z, draws=(E, eps, u))
        c_new = orc.transition(c_prev, T.to(torch.float64), E)
        assert torch.equal(pf.last_pre_resample_classes.cpu(), c_new)
        x_gpu = pf.last_pre_resample_states.cpu()
        mu_o, _, v_o = orc.map_x_to_y(spec, f, x_gpu)
        ll_o = orc.log_likelihoods_fused(mu_o, v_o, torch.as_tensor(z, dtype=torch.float64), spec.y_log_lambdas)
        assert float(torch.max(torch.abs(pf._log_likelihoods.cpu() - ll_o) / torch.abs(ll_o))) < 1e-6
        lw_o, w_o = orc.normalize(pf._log_likelihoods.cpu())
        assert torch.equal(pf._log_weights.cpu(), lw_o)
        anc_o = orc.resample(pf._weights.cpu(), u)
        assert torch.equal(pf.last_ancestors.cpu(), anc_o)
        assert torch.equal(pf._particle_classes.cpu(), c_new[anc_o]) and torch.equal(pf._particle_states.cpu(), x_gpu[anc_o])
        cp_o = orc.class_probabilities(pf._log_likelihoods.cpu(), lw_o, c_new[anc_o], C)
        assert float(torch.max(torch.abs(pf.class_probabilities().cpu() - cp_o) / cp_o.clamp(min=1e-300))) < 1e-12
        assert pf.get_most_likely_class() == int(torch.argmax(cp_o))


def test_clouds_above_the_limit_take_the_general_path(setup):
    spec, wl, f, model = setup
    pf = GPMDM_PF(model, synthetic.markov_matrix(spec.n_classes), 4097, seed=1)
    assert not pf._small and pf._lowlat
    pf.update(wl.test_trials[0][1][0])
    assert abs(float(pf._weights.sum()) - 1.0) < 1e-12


@pytest.mark.parametrize("P", [100, 6000])
def test_update_many_equals_frame_by_frame_updates(setup, P):
    """The batched multi-frame call (small-cloud kernels at P = 100, the general native step at P = 6000) against the
    reference-style loop of update() + queries: same posteriors, classes, state means and final cloud, bit for bit."""
    spec, wl, f, model = setup
    T = synthetic.markov_matrix(spec.n_classes)
    trial = np.concatenate([wl.test_trials[0][1], wl.test_trials[1][1]], 0)
    a, b = GPMDM_PF(model, T, P, seed=4), GPMDM_PF(model, T, P, seed=4)
    k = 7  # two calls: the second one reuses the captured graphs and the frame buffers of the first
    parts = [a.update_many(trial[:k]), a.update_many(trial[k:])]
    probs, cls, means = (torch.cat([p[i] for p in parts]) for i in range(3))
    for t, z in enumerate(trial):
        b.update(z)
        assert torch.equal(b.class_probabilities(), probs[t]) and b.get_most_likely_class() == int(cls[t])
        assert torch.equal(b.current_state_mean(), means[t])
    assert torch.equal(a._particle_states, b._particle_states) and torch.equal(a._particle_classes, b._particle_classes)
    assert torch.equal(a._log_likelihoods, b._log_likelihoods) and a.get_most_likely_class() == b.get_most_likely_class()
