"""CPU: the raw-draw identities that let the same injected arrays drive the reference, the
oracle and the CUDA path: on this torch build the CPU samplers are bit-exactly the formulas below."""
import torch

from oracle import gpmdm_oracle as orc


def test_multinomial_one_draw_is_argmax_of_dist_over_exponential():
    P, C = 4096, 5
    dist = torch.rand(P, C, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) + 0.01
    torch.manual_seed(7)
    ref = torch.multinomial(dist, 1, replacement=True).squeeze(-1)
    torch.manual_seed(7)
    U = torch.rand(P, C, dtype=torch.float64)
    E = -torch.log1p(-U)
    assert torch.equal(ref, torch.argmax(dist / E, dim=-1))


def test_normal_is_randn_times_std_plus_mean():
    mean = torch.randn(1000, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    std = torch.rand(1000, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) + 0.1
    torch.manual_seed(9)
    ref = torch.normal(mean, std)
    torch.manual_seed(9)
    z = torch.randn(1000, 3, dtype=torch.float64)
    assert torch.equal(ref, z * std + mean)


def test_multinomial_resampling_is_search_in_sequential_cdf():
    P = 1 << 16
    w = torch.rand(P, dtype=torch.float64, generator=torch.Generator().manual_seed(4)) ** 8
    w = w / w.sum()
    torch.manual_seed(11)
    ref = torch.multinomial(w, P, replacement=True)
    torch.manual_seed(11)
    u = torch.rand(P, dtype=torch.float64)
    assert torch.equal(ref, orc.resample(w, u))
