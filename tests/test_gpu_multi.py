"""GPU (needs >= 2 devices; run with `gpurun --gpus 2`): the particle-sharded filter over NCCL reproduces the
single-GPU filter bit for bit on the same seed -- states, classes, log-likelihoods, weights, ancestors and
class posteriors."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_filter(P, steps, seed, **kw):
    from gpmdm_b200 import GPMDM_PF, synthetic
    from tests.helpers import product_model_from_spec, synthetic_spec

    spec, wl = synthetic_spec(3, 3, 16, 3, 40, seed=6)
    model = product_model_from_spec(spec)
    pf = GPMDM_PF(model, synthetic.markov_matrix(3), P, seed=seed, **kw)
    trial = wl.test_trials[0][1]
    probs = []
    for t in range(steps):
        pf.update(trial[t])
        probs.append(pf.class_probabilities().cpu())
    torch.cuda.synchronize()
    return dict(states=pf._particle_states.cpu(), classes=pf._particle_classes.cpu(), ll=pf._log_likelihoods.cpu(),
                w=pf._weights.cpu(), anc=pf.last_ancestors.cpu(), probs=torch.stack(probs),
                mean=pf.current_state_mean().cpu())


def _worker(rank, world_size, port, P, steps, seed, out_path):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    res = _run_filter(P, steps, seed)
    if rank == 0:
        torch.save(res, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("P", [4096, 8192, 20480])
@pytest.mark.parametrize("world_size", [2])
def test_sharded_filter_is_bitwise_equal_to_single_gpu(tmp_path, world_size, P):
    """P = 4096: low-latency mode on both sides (and the single-GPU side takes the small-cloud kernels); P = 8192: 128
    tiles in all, 64 per rank -- the mode is chosen from the TOTAL particle count, so both sides run the fused kernels
    (a per-rank choice would have put the 2-GPU run in low-latency mode, with another summation order over k);
    P = 20 480: several tiles per SM on one GPU, fewer than SMs per rank on two."""
    if torch.cuda.device_count() < world_size:
        pytest.skip(f"needs {world_size} GPUs")
    steps, seed = 4, 21
    out = os.path.join(tmp_path, "multi.pt")
    mp.spawn(_worker, args=(world_size, _free_port(), P, steps, seed, out), nprocs=world_size, join=True)
    multi = torch.load(out)
    single = _run_filter(P, steps, seed, distributed=False)
    for k in single:
        assert torch.equal(single[k], multi[k]), k
