"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): the oracle against the
UNMODIFIED reference run live at BASELINE config 1 shapes (C=2, d=3, D=62, N=2000, P=100), beyond the committed
golden vectors."""
import numpy as np
import pytest
import torch

from gpmdm_b200 import synthetic
from oracle import gpmdm_oracle as orc
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def live():
    ref = ref_shim.load_reference()
    import gpmdm.gpmdm_pf as ref_pf_module

    wl = synthetic.make_sequences(2, 62, 10, 100, seed=21, n_test_trials=1, test_frames=12)
    hp = synthetic.notebook_hyperparameters(62, 3, 1e-1)
    model = ref.GPMDM(D=62, d=3, n_classes=2, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(2):
        for s in wl.sequences[c]:
            model.add_data(s, c)
    with torch.no_grad():
        model.init_X()
    return ref, ref_pf_module, model, wl


def test_block_factors_equal_reference_dense_matrices(live):
    ref, mod, model, wl = live
    spec = orc.ModelSpec.from_reference(model)
    f = orc.precompute_factors(spec)
    assert torch.equal(f.Ky_inv, model.Ky_inv.detach()) or \
        float(torch.max(torch.abs(f.Ky_inv - model.Ky_inv.detach())) / torch.max(torch.abs(f.Ky_inv))) < 1e-9
    for c, (a, b) in enumerate(spec.class_pair_ranges()):
        full = model.Kx_inv_class[c].detach()
        blk = full[a:b, a:b]
        assert float(torch.max(torch.abs(blk - f.Kx_inv_blocks[c])) / torch.max(torch.abs(blk))) < 1e-6
        off = full.clone()
        off[a:b, a:b] = 0
        expect = 1e6 * torch.eye(full.shape[0], dtype=full.dtype)
        expect[a:b, a:b] = 0
        assert torch.allclose(off, expect, rtol=1e-12, atol=0)


def test_trial_against_live_reference(live):
    ref, mod, model, wl = live
    P, C, d = 100, 2, 3
    spec = orc.ModelSpec.from_reference(model)
    a_b = spec.class_pair_ranges()
    f = orc.precompute_factors(spec, Ky_inv=model.Ky_inv.detach().clone(),
                               Kx_inv_blocks=[model.Kx_inv_class[c].detach()[a:b, a:b].clone()
                                              for c, (a, b) in enumerate(a_b)])
    T = synthetic.markov_matrix(C)
    parts = orc.divide_into_n_parts(P, C)
    g = torch.Generator().manual_seed(3)
    init_idx = [torch.randint(0, hi - lo, (parts[c],), generator=g) for c, (lo, hi) in enumerate(spec.class_row_ranges())]
    with torch.no_grad():
        with ref_shim.InjectedDraws(mod, init_idx=[i.clone() for i in init_idx]):
            pf = ref.GPMDM_PF(model, T, P)
        o = orc.FilterOracle(spec, T, P, init_idx, f)
        assert torch.equal(o.states, pf._particle_states) and torch.equal(o.classes, pf._particle_classes)
        trial = wl.test_trials[0][1]
        for t in range(6):
            E, eps, u = synthetic.raw_draws(P, C, d, 300 + t)
            with ref_shim.InjectedDraws(mod, E=E, eps=eps, u=u) as inj:
                pf.update(trial[t])
            x_prev, c_prev = o.states.clone(), o.classes.clone()
            o.update(trial[t], E, eps, u, loop_ll=(t % 2 == 0))
            assert torch.equal(o.trace["c_new"], inj.record["new_classes"])
            # stage-wise from the reference's own post-dynamics states (ll is steep in x')
            x_ref = torch.zeros(P, d, dtype=torch.float64)
            for c, (rows, mean) in inj.record["dyn_mean"].items():
                x_ref[rows] = eps[rows] * inj.record["dyn_std"][c][1] + mean
            assert float(torch.max(torch.abs(o.trace["x_new"] - x_ref))) < 1e-6
            mu, var, v = orc.map_x_to_y(spec, f, x_ref)
            ll_o = orc.log_likelihoods_fused(mu, v, torch.as_tensor(trial[t], dtype=torch.float64), spec.y_log_lambdas)
            ll_ref = pf._log_likelihoods
            assert float(torch.max(torch.abs(ll_o - ll_ref) / torch.abs(ll_ref))) < 1e-9
            assert torch.equal(orc.resample(pf._weights, u), inj.record["ancestors"])
            cp = orc.class_probabilities(ll_ref, pf._log_weights, pf._particle_classes, C)
            assert float(torch.max(torch.abs(cp - pf.class_probabilities()))) < 1e-12
            # re-synchronise (variance noise floor)
            o.states, o.classes = pf._particle_states.clone(), pf._particle_classes.clone()
