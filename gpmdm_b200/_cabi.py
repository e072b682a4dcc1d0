"""ctypes binding of libgpmdm_sm100a.so (include/gpmdm_b200.h).

The library is the only implementation of the filter step: if it is missing or a call fails the
caller gets an exception -- there is no CPU or torch fallback behind any of these functions.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import build as _build

ABI_VERSION = 4  # GPMDM_ABI_VERSION of include/gpmdm_b200.h this binding was written against
_i32, _i64, _u64, _f64, _ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p

TILE_P = 64    # GPMDM_TILE_P: particles per predict tile
TILE_N = 256   # GPMDM_TILE_N: row padding of factors / alpha_ld granularity
MAX_LATENT = 8


class GpModel(ctypes.Structure):
    """struct gpmdm_gp_model"""
    _fields_ = [("blocks", _ptr), ("n_blocks", _i32), ("d", _i32), ("dout", _i32), ("alpha_ld", _i32),
                ("kind", _i32), ("tri", _i32), ("lengthscales", _ptr), ("lin_c2", _ptr), ("lambdas", _ptr)]


class GpModelTf32(ctypes.Structure):
    """struct gpmdm_gp_model_tf32"""
    _fields_ = [("coords", _ptr), ("wtiles", _ptr), ("atiles", _ptr), ("n", _i64), ("n_pad", _i64), ("d", _i32),
                ("dout", _i32), ("lengthscales", _ptr), ("lambdas", _ptr)]


class PfSmallIo(ctypes.Structure):
    """struct gpmdm_pf_small_io"""
    _fields_ = [("z_src", _ptr), ("summary_dst", _ptr), ("probs_dst", _ptr), ("frame", _ptr)]


class PfStepArgs(ctypes.Structure):
    """struct gpmdm_pf_step_args"""
    _fields_ = [("dyn", ctypes.POINTER(GpModel)), ("obs", ctypes.POINTER(GpModel)),
                ("P", _i64), ("lo", _i64), ("n_local", _i64), ("C", _i32), ("d", _i32),
                ("generate_draws", _i32), ("systematic", _i32), ("cdf_mode", _i32), ("predict_mode", _i32),
                ("seed", _u64), ("step", _u64), ("T", _ptr), ("z", _ptr), ("ll_const", _f64),
                ("x_prev", _ptr), ("c_prev", _ptr), ("E", _ptr), ("eps", _ptr), ("u", _ptr),
                ("x_new", _ptr), ("c_new", _ptr), ("ll", _ptr),
                ("perm", _ptr), ("tiles", _ptr), ("n_tiles", _ptr), ("tile_counter", _ptr),
                ("workspace", _ptr), ("lowlat_workspace", _ptr), ("obs_n_pad", _i64), ("dyn_max_n_pad", _i64),
                ("obs_seg_chunks", _i32), ("dyn_seg_chunks", _i32),
                ("kstar_workspace", _ptr), ("kstar_workspace_bytes", _i64),
                ("lw", _ptr), ("w", _ptr), ("stats", _ptr), ("cdf", _ptr), ("anc", _ptr), ("x_out", _ptr),
                ("c_out", _ptr)]


_SIGNATURES = {
    "gpmdm_abi_version": (ctypes.c_int, []),
    "gpmdm_last_error": (ctypes.c_char_p, []),
    "gpmdm_quadform_bytes": (_i64, [_i64, ctypes.c_int]),
    "gpmdm_pack_quadform_f64": (ctypes.c_int, [_ptr, _i64, _i64, ctypes.c_int, _ptr, _ptr]),
    "gpmdm_alpha_bytes": (_i64, [_i64, _i32]),
    "gpmdm_pack_alpha_f64": (ctypes.c_int, [_ptr, _i64, _i64, _i32, _i32, _ptr, _ptr]),
    "gpmdm_pf_transition_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _i64, _i32, _ptr, _ptr]),
    "gpmdm_pf_bucket_by_class": (ctypes.c_int, [_ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_bucket_by_class2": (ctypes.c_int, [_ptr, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_dynvar_tc": (ctypes.c_int, [_ptr, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr]),
    "gpmdm_pf_propagate_meanonly_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr,
                                                       _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_propagate_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr,
                                              _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_observe_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _i64, _ptr, _f64, _ptr, _ptr, _ptr,
                                            _ptr, _ptr]),
    "gpmdm_pf_observe_kstar_workspace_bytes": (_i64, [_i64]),
    "gpmdm_pf_observe_cached_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _i64, _ptr, _f64, _ptr, _ptr, _ptr,
                                                   _i64, _ptr, _ptr, _i64, _ptr]),
    "gpmdm_pf_propagate_cached_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr,
                                                     _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr]),
    "gpmdm_pf_loglik_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _i64, _ptr, _f64, _ptr, _ptr, _ptr, _ptr,
                                           _ptr]),
    "gpmdm_predict_lowlat_workspace_bytes": (_i64, [_i64, _i64, _i32, _i32, _i32]),
    "gpmdm_predict_lowlat_pick_segment": (_i32, [_i64, _i64, _i32, _i32]),
    "gpmdm_pf_observe_lowlat_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _i64, _ptr, _f64, _ptr, _ptr, _ptr,
                                                   _ptr, _i64, _i32, _ptr, _ptr, _ptr]),
    "gpmdm_pf_propagate_lowlat_f64": (ctypes.c_int, [ctypes.POINTER(GpModel), _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr,
                                                     _ptr, _ptr, _i64, _i32, _ptr, _ptr, _ptr]),
    "gpmdm_pf_step_local_f64": (ctypes.c_int, [ctypes.POINTER(PfStepArgs), _ptr]),
    "gpmdm_pf_step_global_f64": (ctypes.c_int, [ctypes.POINTER(PfStepArgs), _ptr]),
    "gpmdm_pf_small_max_particles": (_i32, []),
    "gpmdm_pf_step_small_f64": (ctypes.c_int, [ctypes.POINTER(PfStepArgs), _ptr, _ptr, ctypes.POINTER(PfSmallIo), _ptr]),
    "gpmdm_pf_normalize_f64": (ctypes.c_int, [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_cdf_f64": (ctypes.c_int, [_ptr, _i64, _i32, _ptr, _ptr, _ptr]),
    "gpmdm_pf_resample_f64": (ctypes.c_int, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_resample_sorted_f64": (ctypes.c_int, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_pf_summaries_f64": (ctypes.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr, _ptr, _ptr]),
    "gpmdm_pf_draws_philox": (ctypes.c_int, [_u64, _u64, _i64, _i64, _i64, _i32, _i32, _i32, _ptr, _ptr, _ptr,
                                             _ptr]),
    "gpmdm_workspace_bytes": (_i64, [_i64, _i32]),
    "gpmdm_tf32_wtiles_bytes": (_i64, [_i64]),
    "gpmdm_tf32_atiles_bytes": (_i64, [_i64]),
    "gpmdm_pack_whitened_tf32": (ctypes.c_int, [_ptr, _i64, _i64, _ptr, _ptr]),
    "gpmdm_pack_alpha_tf32": (ctypes.c_int, [_ptr, _i64, _i64, _i32, _ptr, _ptr]),
    "gpmdm_pf_observe_tf32": (ctypes.c_int, [ctypes.POINTER(GpModelTf32), _ptr, _i64, _ptr, _f64, _ptr, _ptr, _ptr,
                                             _ptr, _ptr]),
    "gpmdm_kernel_build_f64": (ctypes.c_int, [_ptr, _i64, _i32, _i32, _ptr, _ptr, _f64, _ptr, _i32, _ptr, _ptr]),
    "gpmdm_kernel_grad_f64": (ctypes.c_int, [_ptr, _ptr, _i64, _i32, _i32, _ptr, _ptr, _f64, _ptr, _i32, _ptr, _ptr,
                                             _ptr, _ptr, _ptr, _ptr]),
    "gpmdm_kernel_grad_workspace_bytes": (_i64, [_i64, _i32]),
    "gpmdm_probe_dmma_tflops": (ctypes.c_int, [_i32, ctypes.POINTER(_f64)]),
    "gpmdm_probe_tf32_tflops": (ctypes.c_int, [_i32, ctypes.POINTER(_f64)]),
    "gpmdm_probe_f16_tflops": (ctypes.c_int, [_i32, ctypes.POINTER(_f64)]),
    "gpmdm_f16_wtiles_bytes": (_i64, [_i64]),
    "gpmdm_pack_whitened_f16x2": (ctypes.c_int, [_ptr, _i64, _i64, _ptr, _ptr]),
    "gpmdm_pf_observe_f16x2": (ctypes.c_int, [ctypes.POINTER(GpModelTf32), _ptr, _i64, _ptr, _ptr, _ptr]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def library_path() -> str:
    # GPMDM_LIBRARY: another build of the same sources (tuning experiments); the default is the in-tree build
    return os.environ.get("GPMDM_LIBRARY") or _build.LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (once) the CUDA extension.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m gpmdm_b200.build` (needs nvcc). "
                               "gpmdm_b200 has no CPU or torch fallback for the filter step.")
        handle = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the build is stale
            fn.restype, fn.argtypes = res, args
        if handle.gpmdm_abi_version() != ABI_VERSION:
            raise RuntimeError("libgpmdm_sm100a.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    """Map the C ABI's return convention onto the reference's exception behaviour."""
    if rc == 0:
        return
    msg = lib().gpmdm_last_error().decode(errors="replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (or NULL).  The tensor must be a contiguous CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("gpmdm_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
