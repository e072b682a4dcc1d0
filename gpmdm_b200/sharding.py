"""Particle sharding across ranks (one process per GPU) -- host-side plumbing, backend agnostic.

Particles are split into contiguous, equal index ranges; training data and factors are replicated.  The
only exchange per filter step is `all_gather_particles`: every rank contributes the (state, class,
log-likelihood) records of its own range and receives everybody's, in rank order == particle order, so the
gathered arrays are exactly the arrays a single-GPU run holds.  Weight normalisation, the cdf and the
resampling search then run redundantly on every rank over all P particles in a fixed order
(csrc/pf_stages.cu), which makes a G-GPU run reproduce the 1-GPU run bit for bit.  NCCL on GPUs; the same code
runs over gloo in the CPU tests (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def world(group=None, enabled: Optional[bool] = None) -> Tuple[int, int]:
    """(world_size, rank) of the particle-parallel group; (1, 0) when torch.distributed is not in use."""
    import torch.distributed as dist

    use = (dist.is_available() and dist.is_initialized()) if enabled is None else enabled
    if not use:
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def particle_range(num_particles: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the particles owned by `rank`.  P must divide evenly (as in the BASELINE configs)."""
    if num_particles % world_size != 0:
        raise ValueError(f"num_particles ({num_particles}) must be divisible by the number of ranks ({world_size})")
    per = num_particles // world_size
    return rank * per, (rank + 1) * per


def all_gather_particles(full_states, full_classes, full_ll, lo: int, hi: int, group=None) -> None:
    """In place: fill the [P, d] / [P] / [P] arrays from every rank's [lo, hi) slice (already written locally)."""
    import torch.distributed as dist

    for full in (full_ll, full_states, full_classes):
        dist.all_gather_into_tensor(full, full[lo:hi].clone(), group=group)
