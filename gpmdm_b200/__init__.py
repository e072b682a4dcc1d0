"""gpmdm_b200 -- the GPMDM particle-filter step and GP kernel machinery of Priyanshu4/gpmdm,
re-designed for NVIDIA B200 (sm_100a): PyTorch host code over a C-ABI CUDA library.

    from gpmdm_b200 import GPMDM, GPMDM_PF      # same names as the reference's `gpmdm` package
"""
from .gpmdm import GPMDM
from .gpmdm_pf import GPMDM_PF

__all__ = ["GPMDM", "GPMDM_PF"]
