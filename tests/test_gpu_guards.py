"""GPU: memory-safety canaries and the variance-fault status words.

compute-sanitizer is closed on the GPU pool, so out-of-bounds writes are hunted the old way: every scratch and output
buffer a kernel may touch (the per-CTA K* slices, the low-latency partial-sum workspace, perm / tiles / n_tiles, the stage
workspace, the outputs) is carved out of a larger allocation whose guard bands hold a sentinel; after fused, cached and
low-latency runs at ragged sizes the bands must be untouched.  All calls go through the C ABI directly."""
import ctypes

import numpy as np
import pytest
import torch

from gpmdm_b200 import _cabi, synthetic
from gpmdm_b200._cabi import check, ptr, stream
from oracle import gpmdm_oracle as orc
from tests.helpers import product_model_from_spec, synthetic_spec

pytestmark = pytest.mark.gpu
GUARD = 2048  # elements on each side


class Guarded:
    """`n` elements of `dtype` between two sentinel bands (the interior is 256-byte aligned)."""

    def __init__(self, n, dtype=torch.float64, fill=None):
        self.n = int(n)
        self.sent = -7777 if not dtype.is_floating_point else -7777.25
        self.full = torch.full((self.n + 2 * GUARD,), self.sent, dtype=dtype, device="cuda")
        self.t = self.full[GUARD:GUARD + self.n]
        if fill is not None:
            self.t.fill_(fill)

    def intact(self):
        return bool((self.full[:GUARD] == self.sent).all()) and bool((self.full[GUARD + self.n:] == self.sent).all())


@pytest.fixture(scope="module")
def small():
    spec, wl = synthetic_spec(3, 3, 20, 3, 70, sigma_n=1e-1, seed=8)  # N = 630 -> n_pad 768; N_c = 207 -> 256
    model = product_model_from_spec(spec)
    return spec, wl, model


@pytest.mark.parametrize("P", [1, 63, 65, 130, 1000, 9500])
def test_guard_bands_survive_every_predict_mode(small, P, monkeypatch):
    spec, wl, model = small
    lib = _cabi.lib()
    pk = model.packed_models()
    C, d, D = spec.n_classes, spec.d, spec.D
    g = torch.Generator().manual_seed(P)
    xs = (spec.X[torch.randint(0, spec.N, (P,), generator=g)] + 0.05 * torch.randn(P, d, dtype=torch.float64, generator=g)).cuda()
    z = torch.tensor(wl.test_trials[0][1][0], dtype=torch.float64, device="cuda")
    cls = torch.randint(0, C, (P,), generator=g).cuda()
    eps = torch.randn(P, d, dtype=torch.float64, generator=g).cuda()
    bufs = {}

    def G(name, n, dtype=torch.float64, fill=None):
        bufs[name] = Guarded(n, dtype, fill)
        return bufs[name].t

    counter = G("counter", 4, torch.int32, 0)
    # bucketing
    perm, tiles, n_tiles = G("perm", P, torch.int32), G("tiles", (P // 64 + C + 1) * 4, torch.int32), G("n_tiles", 1, torch.int32)
    ws = G("stage_ws", int(lib.gpmdm_workspace_bytes(P, C)) // 8 + 1)
    check(lib.gpmdm_pf_bucket_by_class(ptr(cls), P, C, ptr(perm), ptr(tiles), ptr(n_tiles), ptr(ws), stream()), "bucket")
    # dynamics: fused and low latency
    x_new, mean, var = G("x_new", P * d), G("dyn_mean", P * d), G("dyn_var", P * d)
    check(lib.gpmdm_pf_propagate_f64(ctypes.byref(pk["dyn"]), ptr(xs), ptr(perm), ptr(tiles), ptr(n_tiles), P, ptr(eps),
                                     ptr(x_new), ptr(mean), ptr(var), ptr(counter), stream()), "propagate")
    seg = 0 if P % 2 else 5  # the default segmentation rule, and an explicit (short) segment length
    ll_ws = G("lowlat_ws", max(int(lib.gpmdm_predict_lowlat_workspace_bytes(P, pk["dyn_max_n_pad"], d, seg, C)),
                               int(lib.gpmdm_predict_lowlat_workspace_bytes(P, pk["obs_n_pad"], D, seg, 1))) // 8)
    # low latency, with the K* evaluated inside the work items and with the shared per-tile slices: bit-identical
    x_new2, x_new3 = G("x_new_lowlat", P * d), G("x_new_lowlat_shared", P * d)
    for mode, out in (("inline", x_new2), ("shared", x_new3)):
        monkeypatch.setenv("GPMDM_LOWLAT_KSTAR", mode)
        ll_ws.fill_(float("nan"))  # the workspace is never memset: every slot a finalize kernel reads must have been written
        check(lib.gpmdm_pf_propagate_lowlat_f64(ctypes.byref(pk["dyn"]), ptr(xs), ptr(perm), ptr(tiles), ptr(n_tiles), P,
                                                ptr(eps), ptr(out), None, None, pk["dyn_max_n_pad"], seg, ptr(counter), ptr(ll_ws),
                                                stream()), "propagate lowlat " + mode)
    assert torch.equal(x_new2, x_new3) and bool(torch.isfinite(x_new2).all())
    # observation: fused, cached, low latency
    outs = []
    for mode in ("fused", "cached", "lowlat"):
        ll, mu, v = G("ll_" + mode, P), G("mu_" + mode, P * D), G("v_" + mode, P)
        if mode == "fused":
            check(lib.gpmdm_pf_observe_f64(ctypes.byref(pk["obs"]), ptr(xs), P, ptr(z), 0.5, ptr(ll), ptr(mu), ptr(v),
                                           ptr(counter), stream()), mode)
        elif mode == "cached":
            # exactly as many K* slices as the launch has CTAs: one slice too few would write into the band
            ctas = min((P + 63) // 64, torch.cuda.get_device_properties(0).multi_processor_count)
            kws = G("kstar_ws", ctas * pk["obs_n_pad"] * 64)
            check(lib.gpmdm_pf_observe_cached_f64(ctypes.byref(pk["obs"]), ptr(xs), P, ptr(z), 0.5, ptr(ll), ptr(mu), ptr(v),
                                                  pk["obs_n_pad"], ptr(counter), ptr(kws), kws.numel() * 8, stream()), mode)
        else:
            res = []
            for kmode in ("inline", "shared"):
                monkeypatch.setenv("GPMDM_LOWLAT_KSTAR", kmode)
                ll_ws.fill_(float("nan"))
                check(lib.gpmdm_pf_observe_lowlat_f64(ctypes.byref(pk["obs"]), ptr(xs), P, ptr(z), 0.5, None, ptr(ll), ptr(mu),
                                                      ptr(v), pk["obs_n_pad"], seg, ptr(counter), ptr(ll_ws), stream()), mode)
                res.append((ll.clone(), mu.clone(), v.clone()))
            monkeypatch.delenv("GPMDM_LOWLAT_KSTAR")
            assert all(torch.equal(a, b) for a, b in zip(*res)) and bool(torch.isfinite(res[0][0]).all())
        outs.append((ll.clone(), mu.clone(), v.clone()))
    # stages
    lw, w, cdf, stats = G("lw", P), G("w", P), G("cdf", P), G("stats", 2)
    anc, xo, co = G("anc", P, torch.int64), G("x_out", P * d), G("c_out", P, torch.int64)
    u = torch.rand(P, dtype=torch.float64, generator=g).cuda()
    summ = G("summary", C + d + 1)
    check(lib.gpmdm_pf_normalize_f64(ptr(outs[0][0]), P, ptr(lw), ptr(w), ptr(stats), ptr(ws), stream()), "normalize")
    for cdf_mode in (0, 1):
        check(lib.gpmdm_pf_cdf_f64(ptr(w), P, cdf_mode, ptr(cdf), ptr(ws), stream()), "cdf")
        check(lib.gpmdm_pf_resample_f64(ptr(cdf), P, ptr(u), P, ptr(x_new), ptr(cls), d, ptr(anc), ptr(xo), ptr(co), stream()),
              "resample")
        us = torch.sort(u).values
        check(lib.gpmdm_pf_resample_sorted_f64(ptr(cdf), P, ptr(us), P, ptr(x_new), ptr(cls), d, ptr(anc), ptr(xo), ptr(co),
                                               stream()), "resample sorted")
    check(lib.gpmdm_pf_summaries_f64(ptr(outs[0][0]), ptr(lw), ptr(w), ptr(co), ptr(xo), P, C, d, ptr(summ), ptr(ws), stream()),
          "summaries")
    E, ee, uu = G("E", P * C), G("eps_gen", P * d), G("u_gen", P)
    check(lib.gpmdm_pf_draws_philox(3, 1, 0, P, P, C, d, 0, ptr(E), ptr(ee), ptr(uu), stream()), "philox")
    c_new = G("c_new", P, torch.int64)
    T = synthetic.markov_matrix(C).to(torch.float64).cuda()
    check(lib.gpmdm_pf_transition_f64(ptr(cls), ptr(T), ptr(E), P, C, ptr(c_new), stream()), "transition")
    torch.cuda.synchronize()
    broken = [k for k, b in bufs.items() if not b.intact()]
    assert not broken, f"guard bands overwritten around: {broken}"
    # and the three observation modes agree (cached == fused bit for bit; low latency to summation order)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert float((outs[0][2] - outs[2][2]).abs().max()) < 1e-10
    assert int(counter[2]) == 0 and int(counter[3]) == 0  # no variance faults on a healthy model


def test_variance_fault_words_count_what_the_reference_turns_into_nan(small):
    """Reference behaviour for a non-positive predictive variance is a silent NaN (sqrt at gpmdm_pf.py:168, log at :189).
    Same values here, plus a count per stage.  A deliberately wrong (inflated) K^-1 makes 1 - k^T K^-1 k negative near the
    training data."""
    from gpmdm_b200 import GPMDM_PF

    spec, wl, model0 = small
    f = orc.precompute_factors(spec)
    C = spec.n_classes
    T = synthetic.markov_matrix(C)
    z = wl.test_trials[0][1][0]
    bad_obs = product_model_from_spec(spec, Ky_inv=f.Ky_inv * 3.0, Kx_inv_blocks=f.Kx_inv_blocks)
    bad_dyn = product_model_from_spec(spec, Ky_inv=f.Ky_inv, Kx_inv_blocks=[b * 3.0 for b in f.Kx_inv_blocks])
    for kw in (dict(low_latency=False, kstar_cache=True), dict(low_latency=False, kstar_cache=False), dict(low_latency=True)):
        # observation faults only: the dynamics stage is healthy, the variances near the data come out negative
        pf = GPMDM_PF(bad_obs, T, 700, seed=3, **kw)
        assert pf.variance_faults() == (0, 0)
        pf.update(z)
        dyn, obs = pf.variance_faults()
        _, var = bad_obs.map_x_to_y(pf.last_pre_resample_states, **kw)
        v = var[:, 0]
        assert dyn == 0 and obs == int((~(v > 0)).sum()) and 0 < obs
        assert int(torch.isnan(pf._log_likelihoods).sum()) == obs     # log of a negative variance, as gpmdm_pf.py:189
        pf.reset()
        assert pf.variance_faults() == (0, 0)
        # dynamics faults: sqrt of a negative variance gives NaN states (gpmdm_pf.py:168), which the observation stage inherits
        pf = GPMDM_PF(bad_dyn, T, 700, seed=3, **kw)
        pf.update(z)
        dyn, obs = pf.variance_faults()
        nan_rows = int(torch.isnan(pf.last_pre_resample_states).any(dim=1).sum())
        assert dyn == nan_rows and dyn > 0 and obs == nan_rows
    good = GPMDM_PF(model0, T, 700, seed=3)
    good.update(z)
    assert good.variance_faults() == (0, 0)
