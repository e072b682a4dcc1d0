"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU filter (gpmdm_b200/sharding.py).
Each rank runs the per-particle stages on its own particle range (the oracle stands in for the CUDA kernels,
which need a GPU), the single exchange step gathers the records, and every rank resolves the global resampling.
The sharded result must equal the unsharded one bit for bit -- the property SURVEY.md section 8e asks for."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpmdm_b200 import sharding, synthetic
from oracle import gpmdm_oracle as orc
from tests.helpers import synthetic_spec


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _one_step(spec, f, T, states, classes, z, E, eps, u, lo, hi, gather):
    """One filter step with particles [lo, hi) computed locally."""
    P, d = states.shape
    c_new = torch.zeros(P, dtype=torch.int64)
    x_new = torch.zeros(P, d, dtype=torch.float64)
    ll = torch.zeros(P, dtype=torch.float64)
    c_new[lo:hi] = orc.transition(classes[lo:hi], T, E[lo:hi])
    x_new[lo:hi], _, _ = orc.dynamics_draw(spec, f, states[lo:hi], c_new[lo:hi], eps[lo:hi])
    mu, var, v = orc.map_x_to_y(spec, f, x_new[lo:hi])
    ll[lo:hi] = orc.log_likelihoods_fused(mu, v, z, spec.y_log_lambdas)
    gather(x_new, c_new, ll, lo, hi)
    lw, w = orc.normalize(ll)
    anc = orc.resample(w, u)
    return x_new[anc], c_new[anc], ll, anc


def _worker(rank, world_size, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    torch.set_num_threads(2)
    spec, wl = synthetic_spec(2, 3, 10, 2, 30, seed=5)
    f = orc.precompute_factors(spec)
    P, C, d = 96, 2, 3
    T = synthetic.markov_matrix(C).to(torch.float64)
    ws, rk = sharding.world()
    lo, hi = sharding.particle_range(P, ws, rk)
    g = torch.Generator().manual_seed(0)
    states = spec.X[torch.randint(0, spec.N, (P,), generator=g)].clone()
    classes = torch.randint(0, C, (P,), generator=g)
    trial = wl.test_trials[0][1]
    for t in range(3):
        E, eps, u = synthetic.raw_draws(P, C, d, 50 + t)
        z = torch.as_tensor(trial[t], dtype=torch.float64)
        states, classes, ll, anc = _one_step(spec, f, T, states, classes, z, E, eps, u, lo, hi,
                                             lambda x, c, l, a, b: sharding.all_gather_particles(x, c, l, a, b))
    if rank == 0:
        torch.save(dict(states=states, classes=classes, ll=ll, anc=anc), out_path)
    # every rank must hold the same replicated state
    chk = [torch.zeros_like(states) for _ in range(world_size)]
    dist.all_gather(chk, states)
    assert all(torch.equal(c, states) for c in chk)
    dist.destroy_process_group()


def test_particle_range_partition():
    for P, G in ((96, 2), (1 << 20, 8), (8, 8)):
        r = [sharding.particle_range(P, G, k) for k in range(G)]
        assert r[0][0] == 0 and r[-1][1] == P and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    with pytest.raises(ValueError):
        sharding.particle_range(10, 4, 0)
    assert sharding.world() == (1, 0)


@pytest.mark.timeout(180)
def test_two_rank_filter_equals_single_rank(tmp_path):
    out = os.path.join(tmp_path, "sharded.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    # unsharded run of the same three steps
    spec, wl = synthetic_spec(2, 3, 10, 2, 30, seed=5)
    f = orc.precompute_factors(spec)
    P, C, d = 96, 2, 3
    T = synthetic.markov_matrix(C).to(torch.float64)
    g = torch.Generator().manual_seed(0)
    states = spec.X[torch.randint(0, spec.N, (P,), generator=g)].clone()
    classes = torch.randint(0, C, (P,), generator=g)
    trial = wl.test_trials[0][1]
    for t in range(3):
        E, eps, u = synthetic.raw_draws(P, C, d, 50 + t)
        z = torch.as_tensor(trial[t], dtype=torch.float64)
        states, classes, ll, anc = _one_step(spec, f, T, states, classes, z, E, eps, u, 0, P, lambda *a: None)
    assert torch.equal(got["anc"], anc) and torch.equal(got["classes"], classes)
    assert torch.equal(got["states"], states) and torch.equal(got["ll"], ll)
