"""`GPMDM_PF` -- the particle filter of the reference (`gpmdm/gpmdm_pf.py`) with the same constructor
and step API, every stage of `update(z)` running as CUDA kernels of libgpmdm_sm100a.so.

Reference surface kept (file:line of the reference): ctor :47-85 | update :117 | log_likelihood :215 |
class_probabilities :224 | get_most_likely_class :250 | current_state_mean :256 | reset :264 |
properties :267-285 | state attributes `_particle_states`, `_particle_classes`, `_log_weights`,
`_weights`, `_log_likelihoods` :78-82.

Step pipeline (all asynchronous on the current CUDA stream, no host synchronisation inside):
    draws (Philox on device, or injected)            gpmdm_pf_draws_philox
    class transition                                 gpmdm_pf_transition_f64      gpmdm_pf.py:137-151
    bucket particles by class                        gpmdm_pf_bucket_by_class     gpmdm_pf.py:161
    dynamics GP + draw  (DMMA, fused)                gpmdm_pf_propagate_f64       gpmdm_pf.py:153-168
    observation GP + log-likelihood (DMMA, fused)    gpmdm_pf_observe_f64         gpmdm_pf.py:170-192
    [G > 1: all-gather of (x', c', ll) over NCCL]
    normalise weights                                gpmdm_pf_normalize_f64       gpmdm_pf.py:200-204
    cdf + resampling search + gather                 gpmdm_pf_cdf/resample_f64    gpmdm_pf.py:206-213
    class posteriors / state mean (on demand)        gpmdm_pf_summaries_f64       gpmdm_pf.py:215-262

Randomness.  The reference draws from torch's global CPU generator.  Here the raw draws are explicit:
`update(z)` generates them on the device with Philox keyed by (seed, step, global particle index), and
`update(z, draws=(E, eps, u))` consumes caller-supplied draws with exactly the semantics torch's CPU
samplers have -- the parity tests inject the same arrays into the reference.

Multi-GPU.  With `torch.distributed` initialised (one process per GPU, NCCL) particles are sharded by
contiguous index range; factors are replicated.  The one exchange per step is an all-gather of the
per-particle (state, class, log-likelihood) records; every rank then runs the same fixed-order
normalise / cdf / search over all P particles, so a G-GPU run equals the 1-GPU run bit for bit.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _cabi, sharding
from ._cabi import check, ptr, stream
from .gpmdm import GPMDM

_LOG_2PI = torch.log(torch.tensor(2 * 3.14159265358979323846))  # float32, as gpmdm_pf.py:5


class GPMDM_PF:
    """GPMDM particle filter (reference gpmdm_pf.py:7), B200-native."""

    def __init__(self, gpmdm: GPMDM, markov_switching_model, num_particles: int, *,
                 seed: int = 0, resampling: str = "multinomial", cdf_order: str = "sequential",
                 tri: bool = True, precision: Optional[str] = None, low_latency: Optional[bool] = None,
                 kstar_cache: Optional[bool] = None, native_step: bool = True, init_indices: Optional[Sequence] = None, process_group=None,
                 distributed: Optional[bool] = None, cuda_graph: Optional[bool] = None):
        """
        gpmdm, markov_switching_model [C, C], num_particles: as the reference (:47-50).
        seed            Philox key for device-side draws (the reference has no seed argument)
        resampling      'multinomial' (reference, :211) or 'systematic' (one uniform, comb of P points)
        cdf_order       'sequential' = the reference's running sum, bit-identical cdf; 'blocked' = parallel
                        scan in fixed 1024-element blocks
        tri             use the triangular packing of K^-1 (half the flops of the dense quadratic form)
        precision       'fp64' (exact path), or 'tf32' / 'f16x2': observation GP on tcgen05 tensor cores with error-compensated
                        products (3 x tf32, or 2 x fp16 split at twice the rate) and the whitened variance (~1e-4 relative);
                        dynamics / resampling stay fp64.
                        None = 'fp64' for a float64 model, 'tf32' for a float32 model (reference ctor dtype, gpmdm.py:108)
        low_latency     None = automatic: with fewer 64-particle tiles than SMs the column tiles of each particle tile are
                        split over the SMs (two kernels per GP stage instead of one); True / False to force
        kstar_cache     None = automatic: the fused fp64 observation kernel keeps each particle tile's cross-kernel K* in a
                        per-SM scratch (#SMs x N_pad x 64 doubles) instead of re-evaluating it for every column tile;
                        bit-identical results; True / False to force
        native_step     issue the stage launches of a step from native code (two C-ABI calls per step instead of twelve;
                        matters when a step is launch-latency bound, i.e. with few particles); same kernels, same results
        init_indices    optional per-class index tensors replacing torch.randint in _init_particles (:113)
        cuda_graph      small clouds (P <= 4096, one GPU, low-latency mode): the step is six kernels (csrc/pf_small.cu) with
                        no per-step host parameter; None / True = capture them into a CUDA graph after two warm-up steps
                        and replay it per frame, False = launch them directly.  Same results either way
        """
        self._lib = _cabi.lib()
        self._gpmdm = gpmdm
        self._gpmdm.set_evaluation_mode()
        # cast to the model dtype as gpmdm_pf.py:71 does (a float32 model sees float32 transition probabilities), then
        # to the fp64 the kernels compute in
        self._markov_switching_model = torch.as_tensor(markov_switching_model).type(self.dtype).to(
            device=self.device, dtype=torch.float64).contiguous()
        self._num_particles = int(num_particles)
        if self._gpmdm.n_classes != self._markov_switching_model.size(0):
            raise ValueError("Number of classes in the GPMDM model and the Markov model do not match")
        if resampling not in ("multinomial", "systematic"):
            raise ValueError("resampling must be 'multinomial' or 'systematic'")
        if cdf_order not in ("sequential", "blocked"):
            raise ValueError("cdf_order must be 'sequential' or 'blocked'")
        if gpmdm.dyn_back_step != 1:
            raise ValueError("the particle filter uses the dynamics GP output as the next state: dyn_back_step must be 1")
        self._seed, self._step, self._resets = int(seed), 0, 0
        self._systematic = resampling == "systematic"
        self._cdf_mode = 0 if cdf_order == "sequential" else 1
        self._tri = bool(tri)
        if precision is None:
            precision = "tf32" if gpmdm.dtype == torch.float32 else "fp64"
        if precision not in ("fp64", "tf32", "f16x2"):
            raise ValueError("precision must be 'fp64', 'tf32' or 'f16x2'")
        self._precision = precision

        # one process per GPU; contiguous particle ranges (gpmdm_b200/sharding.py)
        self._pg = process_group
        self._world, self._rank = sharding.world(process_group, distributed)
        self._lo, self._hi = sharding.particle_range(self._num_particles, self._world, self._rank)

        self._packed = gpmdm.packed_models(self._tri, with_obs_L=(precision == "fp64"))
        self._packed_tf32 = gpmdm.packed_model_tf32(precision) if precision != "fp64" else None
        self._tc_dyn = None  # tensor-core variants in fused mode: the dynamics variance runs on tcgen05 as well
        c32 = float(0.5 * self._gpmdm.D * _LOG_2PI)  # fp32 product, as gpmdm_pf.py:191
        self._ll_const = self._packed["ll_const_terms"] - c32
        # the mode is chosen from the TOTAL particle count: the two decompositions sum over k in different orders, so a
        # choice that depended on this rank's share would make a G-GPU run differ from the 1-GPU run
        self._lowlat = gpmdm._use_lowlat(self._num_particles, low_latency)
        self._kstar_cache = (precision == "fp64" and not self._lowlat
                             and gpmdm._use_kstar_cache(self._packed["obs_n_pad"], kstar_cache))
        # ... and the dynamics kernel shares that scratch when its class blocks are large enough for the cache to pay
        # (same rule as csrc/pf_step.cu: >= 4 column panels, no block larger than the observation block)
        self._kstar_cache_dyn = (self._kstar_cache and self._packed["dyn_max_n_pad"] >= 4 * _cabi.TILE_N
                                 and self._packed["dyn_max_n_pad"] <= self._packed["obs_n_pad"])
        self._native_step = bool(native_step) and precision == "fp64"
        if precision != "fp64" and not self._lowlat:
            self._tc_dyn = gpmdm.packed_model_tc_dyn(precision)
        # small-cloud step: draws+transition+bucketing and normalise+cdf+resample+summaries as two single-CTA kernels
        self._small = (self._native_step and self._lowlat and self._world == 1
                       and self._num_particles <= int(self._lib.gpmdm_pf_small_max_particles()))
        self._use_graph = self._small and (cuda_graph is None or bool(cuda_graph))
        self._alloc()
        self._init_particles(init_indices)

    # ---- buffers ------------------------------------------------------------------------------------------
    def _alloc(self):
        P, d, C, dev, f64 = self._num_particles, self.latent_dim, self.num_classes, self.device, torch.float64
        Pl = self._hi - self._lo
        e = lambda *s, dt=f64: torch.empty(*s, dtype=dt, device=dev)
        self._x_new, self._c_new, self._ll_all = e(P, d), e(P, dt=torch.int64), e(P)
        self._cdf, self._anc = e(P), e(P, dt=torch.int64)
        self._states_alt, self._classes_alt = e(P, d), e(P, dt=torch.int64)
        self._perm, self._tiles = e(Pl, dt=torch.int32), e(Pl // _cabi.TILE_P + C + 1, 4, dt=torch.int32)
        self._n_tiles, self._counter = e(1, dt=torch.int32), torch.zeros(4, dtype=torch.int32, device=dev)
        self._E, self._eps, self._u = e(Pl, C), e(Pl, d), e(P)
        self._stats = e(2)
        self._v_buf = e(Pl)
        if self._tc_dyn is not None:  # 128-particle class-homogeneous tiles + the dynamics variances
            self._tiles128, self._n_tiles128 = e(Pl // 128 + C + 1, 4, dt=torch.int32), e(1, dt=torch.int32)
            self._v_dyn = e(Pl)
        self._summary = e(C + d + 1)
        self._summary_step = -1
        if self._small:
            # ping-pong sets: the graph of parity p reads (states, classes)[p] and writes ll / lw / w [p] and
            # (states, classes)[1 - p]; the arrays of the previous step stay intact for one more step
            self._S, self._Cl = [e(P, d), e(P, d)], [e(P, dt=torch.int64), e(P, dt=torch.int64)]
            self._LL, self._LW, self._W = [e(P), e(P)], [e(P), e(P)], [e(P), e(P)]
            self._par, self._graphs, self._small_steps = 0, [None, None], 0
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self._step_dev_host = 0
            self._z_buf = e(self.observation_dim)
            self._z_pin = [torch.empty(self.observation_dim, dtype=f64).pin_memory() for _ in range(4)]
            self._z_ev = [None] * 4
            self._z_slot = 0
            # graph mode: the observation's host-to-device copy and the summaries' device-to-host copy are nodes of the
            # graph (one pinned buffer each; `_graph_done` orders their reuse), so a frame is ONE launch from the host
            self._z_graph_pin = torch.empty(self.observation_dim, dtype=f64).pin_memory()
            self._summary_pin = torch.zeros(C + d + 1, dtype=f64).pin_memory()
            self._z_graph_np, self._summary_np = self._z_graph_pin.numpy(), self._summary_pin.numpy()  # views: no torch op per frame
            self._probs = [e(C), e(C)]  # class posteriors of the last two steps on the device (no clone per query)
            self._graph_done = torch.cuda.Event()
            self._summary_host_step = -1
        ws = int(self._lib.gpmdm_workspace_bytes(P, C))
        self._ws = torch.empty(ws // 8 + 1, dtype=torch.float64, device=dev)
        self._seg_obs = self._seg_dyn = 0
        if self._lowlat:
            # k-segment lengths of the low-latency work items, chosen ONCE from the whole cloud (all ranks agree) so that
            # the item grid fills the SMs in whole waves; a fixed function of (P, C, model sizes)
            tiles = (P + _cabi.TILE_P - 1) // _cabi.TILE_P
            # (from the TOTAL particle count, never from this rank's share: the choice fixes the summation order over k)
            self._seg_obs = int(self._lib.gpmdm_predict_lowlat_pick_segment(tiles, self._packed["obs_n_pad"], _cabi.TILE_N,
                                                                            int(self._tri)))
            # dynamics tiles are class-homogeneous: a cloud of `tiles` full tiles usually leaves one more ragged one (100
            # particles, 70 / 30 over two classes: 64 + 6 + 30), and the initial cloud populates every class
            self._seg_dyn = int(self._lib.gpmdm_predict_lowlat_pick_segment(max(tiles + (1 if C > 1 else 0), min(C, P)),
                                                                            self._packed["dyn_max_n_pad"], _cabi.TILE_N,
                                                                            int(self._tri)))
            need = max(int(self._lib.gpmdm_predict_lowlat_workspace_bytes(Pl, self._packed["obs_n_pad"], self.observation_dim, self._seg_obs, 1)),
                       int(self._lib.gpmdm_predict_lowlat_workspace_bytes(Pl, self._packed["dyn_max_n_pad"], d, self._seg_dyn, C)))
            self._ws_lowlat = torch.empty(need // 8 + 1, dtype=torch.float64, device=dev)
        # scratch is owned by the filter instance (two filters on one model may run on different streams)
        self._ws_kstar = None
        if self._kstar_cache:
            need = int(self._lib.gpmdm_pf_observe_kstar_workspace_bytes(self._packed["obs_n_pad"])) // 8
            self._ws_kstar = torch.empty(need, dtype=f64, device=dev)
        if self._native_step:  # struct gpmdm_pf_step_args: the per-trial constants; per-step fields are set in _update
            a = self._step_args = _cabi.PfStepArgs()
            a.dyn, a.obs = ctypes.pointer(self._packed["dyn"]), ctypes.pointer(self._packed["obs"])
            a.P, a.lo, a.n_local, a.C, a.d = P, self._lo, Pl, C, d
            a.systematic, a.cdf_mode = int(self._systematic), self._cdf_mode
            a.predict_mode = 2 if self._lowlat else (1 if self._kstar_cache else 0)
            a.seed, a.T, a.ll_const = self._seed, ptr(self._markov_switching_model), self._ll_const
            a.u, a.x_new, a.c_new, a.ll = ptr(self._u), ptr(self._x_new), ptr(self._c_new), ptr(self._ll_all)
            a.perm, a.tiles, a.n_tiles, a.tile_counter = ptr(self._perm), ptr(self._tiles), ptr(self._n_tiles), ptr(self._counter)
            a.workspace = ptr(self._ws)
            a.lowlat_workspace = ptr(self._ws_lowlat) if self._lowlat else None
            a.obs_n_pad, a.dyn_max_n_pad = self._packed["obs_n_pad"], self._packed["dyn_max_n_pad"]
            a.obs_seg_chunks, a.dyn_seg_chunks = self._seg_obs, self._seg_dyn
            a.kstar_workspace = ptr(self._ws_kstar) if self._kstar_cache else None
            a.kstar_workspace_bytes = self._ws_kstar.numel() * 8 if self._kstar_cache else 0
            a.stats, a.cdf, a.anc = ptr(self._stats), ptr(self._cdf), ptr(self._anc)

    # ---- initialisation (gpmdm_pf.py:87-115) -------------------------------------------------------------------
    def _init_particles(self, init_indices=None):
        P, C = self._num_particles, self.num_classes
        n_per_class = self._divide_into_n_parts(P, C)
        parts = []
        # seeded so that every rank of a multi-GPU run draws the same (replicated) initial cloud
        gen = torch.Generator(device=self.device)
        gen.manual_seed(self._seed + 0x5EED * (self._resets + 1))
        for i in range(C):
            class_data = self._gpmdm.get_X_for_class(i)
            if init_indices is not None:
                idx = torch.as_tensor(init_indices[i]).to(self.device)
                if idx.numel() != n_per_class[i]:
                    raise ValueError("init_indices[%d] must hold %d indices" % (i, n_per_class[i]))
            else:
                idx = torch.randint(0, class_data.size(0), (n_per_class[i],), device=self.device, generator=gen)
            parts.append(class_data[idx].detach().to(torch.float64))
        self._resets += 1
        self._particle_states = torch.cat(parts, dim=0).contiguous()
        self._particle_classes = torch.repeat_interleave(
            torch.arange(C, dtype=torch.int64, device=self.device),
            torch.tensor(n_per_class, dtype=torch.int64, device=self.device))
        # the particle cloud and its weights are held in fp64 whatever the model's public dtype
        self._log_likelihoods = torch.zeros(P, dtype=torch.float64, device=self.device)
        self._log_weights = torch.zeros(P, dtype=torch.float64, device=self.device)
        self._weights = torch.ones(P, dtype=torch.float64, device=self.device) / P
        self._summary_step = self._summary_host_step = -1

    # ---- the filter step (gpmdm_pf.py:117-135) -----------------------------------------------------------------
    def update(self, z, draws=None):
        """Update the particle filter with a new observation z [D] (numpy / sequence / tensor).
        draws: optional (E [P,C] Exp(1), eps [P,d] N(0,1), u [P] U(0,1)) raw draws for ALL particles."""
        if self._small and not (isinstance(z, torch.Tensor) and z.is_cuda):
            if draws is None and self._use_graph and self._small_steps >= 2 and getattr(self, "_profile_events", None) is None:
                return self._replay_small(z)
            return self._update(self._stage_z(z), draws)
        z = torch.as_tensor(np.asarray(z) if not isinstance(z, torch.Tensor) else z).to(device=self.device, dtype=torch.float64)
        self._update(z.contiguous(), draws)

    @torch.no_grad()
    def update_many(self, Z):
        """Batched multi-frame update -- the caller loop of the reference's notebooks/test_gpmdm_pf.ipynb cell 4
        (`for z in trial: pf.update(z); pf.get_most_likely_class(); pf.class_probabilities()`) as ONE call: the T
        observations are uploaded once, the T steps and their summaries are enqueued back to back with no host
        synchronisation in between, and the per-frame class posteriors [T, C], argmax classes [T] and state means [T, d]
        come back as device tensors.  Identical to T calls of update() + the three queries."""
        Z = torch.as_tensor(np.asarray(Z) if not isinstance(Z, torch.Tensor) else Z).to(device=self.device, dtype=torch.float64)
        if Z.dim() != 2 or Z.shape[1] != self.observation_dim:
            raise ValueError("Z must be [T, D = %d]" % self.observation_dim)
        Z = Z.contiguous()
        T, C, d = Z.shape[0], self.num_classes, self.latent_dim
        if self._small and self._use_graph and getattr(self, "_profile_events", None) is None:
            out = self._replay_many(Z)
        else:
            out = torch.empty(T, C + d + 1, dtype=torch.float64, device=self.device)
            for t in range(T):
                self._update(Z[t])
                out[t].copy_(self._summaries())
        self._summary_host_step = -1
        probs = out[:, :C]
        return probs, torch.argmax(probs, dim=1), out[:, C:C + d].to(self.dtype)

    def _stage_z(self, z):
        """Host observation -> the persistent device buffer the step reads, through a small ring of pinned buffers (a slot
        is reused only after the copy that read it has completed)."""
        i = self._z_slot
        self._z_slot = (i + 1) % len(self._z_pin)
        if self._z_ev[i] is not None:
            self._z_ev[i].synchronize()
        src = z if isinstance(z, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(z))
        if src.numel() != self.observation_dim:
            raise ValueError("z must have D = %d entries" % self.observation_dim)
        self._z_pin[i].copy_(src.reshape(-1))
        self._z_buf.copy_(self._z_pin[i], non_blocking=True)
        if self._z_ev[i] is None:
            self._z_ev[i] = torch.cuda.Event()
        self._z_ev[i].record()
        return self._z_buf

    def _update(self, z, draws=None):
        lib, st = self._lib, stream()
        P, d, C = self._num_particles, self.latent_dim, self.num_classes
        lo, hi = self._lo, self._hi
        Pl = hi - lo
        if z.numel() != self.observation_dim:
            raise ValueError("z must have D = %d entries" % self.observation_dim)
        # the stage-by-stage sequence below is also what csrc/pf_step.cu issues natively (bench.py uses the stages to put
        # CUDA events around the dominant kernel)
        native = self._native_step and getattr(self, "_profile_events", None) is None
        # -- raw draws
        if draws is None:
            if not native:
                check(lib.gpmdm_pf_draws_philox(self._seed, self._step, lo, Pl, P, C, d, 0, ptr(self._E), ptr(self._eps),
                                                None, st), "gpmdm_pf_draws_philox")
                check(lib.gpmdm_pf_draws_philox(self._seed, self._step, 0, P, P, C, d, int(self._systematic), None, None,
                                                ptr(self._u), st), "gpmdm_pf_draws_philox")
            E, eps, u = self._E, self._eps, self._u
        else:
            E, eps, u = (torch.as_tensor(a).to(device=self.device, dtype=torch.float64) for a in draws)
            if E.shape != (P, C) or eps.shape != (P, d) or u.shape != (P,):
                raise ValueError("draws must be (E [P,C], eps [P,d], u [P])")
            E, eps, u = E[lo:hi].contiguous(), eps[lo:hi].contiguous(), u.contiguous()
        x_prev = self._particle_states[lo:hi]
        c_prev = self._particle_classes[lo:hi]
        x_new_l, c_new_l, ll_l = self._x_new[lo:hi], self._c_new[lo:hi], self._ll_all[lo:hi]
        if native and self._small:
            return self._update_small(z, draws is None, E, eps, u)
        if native:
            return self._update_native(z, draws is None, E, eps, u, x_prev, c_prev)
        # -- class transition, bucketing, dynamics draw, observation likelihood (local particles)
        check(lib.gpmdm_pf_transition_f64(ptr(c_prev), ptr(self._markov_switching_model), ptr(E), Pl, C, ptr(c_new_l), st),
              "gpmdm_pf_transition_f64")
        if self._tc_dyn is not None:
            # tensor-core variants: the O(N_c^2) part of the variance, 1 - |W_c k_rbf|^2, on tcgen05 per class block; the
            # low-rank (linear kernel) part, the means and the draw in fp64 on the alpha tiles only
            tc = self._tc_dyn
            check(lib.gpmdm_pf_bucket_by_class2(ptr(c_new_l), Pl, C, ptr(self._perm), ptr(self._tiles), ptr(self._n_tiles),
                                                ptr(self._tiles128), ptr(self._n_tiles128), ptr(self._ws), st),
                  "gpmdm_pf_bucket_by_class2")
            check(lib.gpmdm_pf_dynvar_tc(ptr(tc["table"]), C, d, tc["mode"], ptr(tc["ls"]), ptr(x_prev), ptr(self._perm),
                                         ptr(self._tiles128), ptr(self._n_tiles128), Pl, ptr(self._v_dyn), st),
                  "gpmdm_pf_dynvar_tc")
            check(lib.gpmdm_pf_propagate_meanonly_f64(ctypes.byref(tc["model"]), ptr(tc["H"]), ptr(x_prev), ptr(self._perm),
                                                      ptr(self._tiles), ptr(self._n_tiles), Pl, ptr(eps), ptr(self._v_dyn),
                                                      ptr(x_new_l), None, None, ptr(self._counter), st),
                  "gpmdm_pf_propagate_meanonly_f64")
        else:
            check(lib.gpmdm_pf_bucket_by_class(ptr(c_new_l), Pl, C, ptr(self._perm), ptr(self._tiles), ptr(self._n_tiles),
                                               ptr(self._ws), st), "gpmdm_pf_bucket_by_class")
            if self._lowlat:
                check(lib.gpmdm_pf_propagate_lowlat_f64(ctypes.byref(self._packed["dyn"]), ptr(x_prev), ptr(self._perm),
                                                        ptr(self._tiles), ptr(self._n_tiles), Pl, ptr(eps), ptr(x_new_l), None,
                                                        None, self._packed["dyn_max_n_pad"], self._seg_dyn, ptr(self._counter),
                                                        ptr(self._ws_lowlat), st), "gpmdm_pf_propagate_lowlat_f64")
            elif self._kstar_cache_dyn:
                check(lib.gpmdm_pf_propagate_cached_f64(ctypes.byref(self._packed["dyn"]), ptr(x_prev), ptr(self._perm),
                                                        ptr(self._tiles), ptr(self._n_tiles), Pl, ptr(eps), ptr(x_new_l), None,
                                                        None, self._packed["dyn_max_n_pad"], ptr(self._counter),
                                                        ptr(self._ws_kstar), self._ws_kstar.numel() * 8, st),
                      "gpmdm_pf_propagate_cached_f64")
            else:
                check(lib.gpmdm_pf_propagate_f64(ctypes.byref(self._packed["dyn"]), ptr(x_prev), ptr(self._perm),
                                                 ptr(self._tiles), ptr(self._n_tiles), Pl, ptr(eps), ptr(x_new_l), None, None,
                                                 ptr(self._counter), st), "gpmdm_pf_propagate_f64")
        prof = getattr(self, "_profile_events", None)
        if prof is not None:  # bench.py: CUDA events around the dominant kernel, on the launching stream
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self._precision != "fp64":
            # variances: tcgen05 (3 x tf32 or 2 x fp16 split, whitened form); means + log-likelihood: fp64 on the alpha tile only
            if self._precision == "tf32":
                check(lib.gpmdm_pf_observe_tf32(ctypes.byref(self._packed_tf32["model"]), ptr(x_new_l), Pl, None, 0.0, None,
                                                None, ptr(self._v_buf), ptr(self._counter), st), "gpmdm_pf_observe_tf32")
            else:
                check(lib.gpmdm_pf_observe_f16x2(ctypes.byref(self._packed_tf32["model"]), ptr(x_new_l), Pl, ptr(self._v_buf),
                                                 ptr(self._counter), st), "gpmdm_pf_observe_f16x2")
            if prof is not None:  # the tensor-core kernel alone (the fp64 mean tile follows)
                ev = ev + (torch.cuda.Event(enable_timing=True),)
                ev[2].record()
            check(lib.gpmdm_pf_loglik_f64(ctypes.byref(self._packed["obs"]), ptr(x_new_l), Pl, ptr(z), self._ll_const,
                                          ptr(self._v_buf), ptr(ll_l), None, ptr(self._counter), st),
                  "gpmdm_pf_loglik_f64")
        elif self._lowlat:
            check(lib.gpmdm_pf_observe_lowlat_f64(ctypes.byref(self._packed["obs"]), ptr(x_new_l), Pl, ptr(z),
                                                  self._ll_const, None, ptr(ll_l), None, None, self._packed["obs_n_pad"],
                                                  self._seg_obs, ptr(self._counter), ptr(self._ws_lowlat), st),
                  "gpmdm_pf_observe_lowlat_f64")
        elif self._kstar_cache:
            check(lib.gpmdm_pf_observe_cached_f64(ctypes.byref(self._packed["obs"]), ptr(x_new_l), Pl, ptr(z),
                                                  self._ll_const, ptr(ll_l), None, None, self._packed["obs_n_pad"],
                                                  ptr(self._counter), ptr(self._ws_kstar), self._ws_kstar.numel() * 8, st),
                  "gpmdm_pf_observe_cached_f64")
        else:
            check(lib.gpmdm_pf_observe_f64(ctypes.byref(self._packed["obs"]), ptr(x_new_l), Pl, ptr(z), self._ll_const,
                                           ptr(ll_l), None, None, ptr(self._counter), st), "gpmdm_pf_observe_f64")
        if prof is not None:
            ev[1].record()
            prof.append(ev)
        # -- the one exchange step: every rank gets every particle's (x', c', ll)
        if self._world > 1:
            sharding.all_gather_particles(self._x_new, self._c_new, self._ll_all, lo, hi, self._pg)
        # -- weights, cdf, resampling over all P particles (fixed order => identical on every rank)
        ll_new, lw_new, w_new = self._ll_all, self._log_weights_buf(), self._weights_buf()
        check(lib.gpmdm_pf_normalize_f64(ptr(ll_new), P, ptr(lw_new), ptr(w_new), ptr(self._stats), ptr(self._ws), st),
              "gpmdm_pf_normalize_f64")
        check(lib.gpmdm_pf_cdf_f64(ptr(w_new), P, self._cdf_mode, ptr(self._cdf), ptr(self._ws), st), "gpmdm_pf_cdf_f64")
        # device-generated systematic draws are an ascending comb: windowed search with coalesced accesses
        resample = lib.gpmdm_pf_resample_sorted_f64 if (self._systematic and draws is None) else lib.gpmdm_pf_resample_f64
        check(resample(ptr(self._cdf), P, ptr(u), P, ptr(self._x_new), ptr(self._c_new), d,
                                        ptr(self._anc), ptr(self._states_alt), ptr(self._classes_alt), st),
              "gpmdm_pf_resample_f64")
        # publish (swap buffers; no copies)
        self._particle_states, self._states_alt = self._states_alt, self._particle_states
        self._particle_classes, self._classes_alt = self._classes_alt, self._particle_classes
        self._log_likelihoods = ll_new.clone()
        self._log_weights, self._weights = lw_new, w_new
        self._step += 1

    def _update_native(self, z, generate, E, eps, u, x_prev, c_prev):
        """The same step through gpmdm_pf_step_local_f64 / gpmdm_pf_step_global_f64 (csrc/pf_step.cu)."""
        lib, st, a = self._lib, stream(), self._step_args
        lw_new, w_new = self._log_weights_buf(), self._weights_buf()
        a.generate_draws, a.step, a.z = int(generate), self._step, ptr(z)
        a.x_prev, a.c_prev, a.E, a.eps, a.u = ptr(x_prev), ptr(c_prev), ptr(E), ptr(eps), ptr(u)
        a.lw, a.w, a.x_out, a.c_out = ptr(lw_new), ptr(w_new), ptr(self._states_alt), ptr(self._classes_alt)
        check(lib.gpmdm_pf_step_local_f64(ctypes.byref(a), st), "gpmdm_pf_step_local_f64")
        if self._world > 1:
            sharding.all_gather_particles(self._x_new, self._c_new, self._ll_all, self._lo, self._hi, self._pg)
        check(lib.gpmdm_pf_step_global_f64(ctypes.byref(a), st), "gpmdm_pf_step_global_f64")
        self._particle_states, self._states_alt = self._states_alt, self._particle_states
        self._particle_classes, self._classes_alt = self._classes_alt, self._particle_classes
        self._log_likelihoods = self._ll_all.clone()
        self._log_weights, self._weights = lw_new, w_new
        self._step += 1

    # ---- small clouds: six kernels per step, replayed from a CUDA graph ------------------------------------------------
    def _issue_small(self, par, generate, E, eps, u, io=None):
        a = self._step_args
        a.generate_draws, a.step, a.z = int(generate), self._step, ptr(self._z_buf)
        a.x_prev, a.c_prev, a.E, a.eps, a.u = ptr(self._S[par]), ptr(self._Cl[par]), ptr(E), ptr(eps), ptr(u)
        a.ll, a.lw, a.w = ptr(self._LL[par]), ptr(self._LW[par]), ptr(self._W[par])
        a.x_out, a.c_out = ptr(self._S[1 - par]), ptr(self._Cl[1 - par])
        check(self._lib.gpmdm_pf_step_small_f64(ctypes.byref(a), ptr(self._step_dev), ptr(self._summary),
                                                ctypes.byref(io) if io is not None else None, stream()),
              "gpmdm_pf_step_small_f64")

    def _update_small(self, z, generate, E, eps, u):
        par = self._par
        # the step reads the particle cloud from its own buffers: adopt a cloud that was rebound from outside
        if self._particle_states.data_ptr() != self._S[par].data_ptr():
            self._S[par].copy_(self._particle_states)
        if self._particle_classes.data_ptr() != self._Cl[par].data_ptr():
            self._Cl[par].copy_(self._particle_classes)
        if z.data_ptr() != self._z_buf.data_ptr():
            self._z_buf.copy_(z)
        if self._step_dev_host != self._step:  # keeps the device step key equal to the host's (reset(), injected draws)
            self._step_dev.fill_(self._step)
        self._issue_small(par, generate, E, eps, u)
        self._finish_small(par)

    def _replay_small(self, z):
        """One frame = one graph launch of six kernel nodes: pre (reads z through the mapping of the pinned host buffer) ->
        dynamics GP (items, finalise) -> observation GP (items, finalise) -> post (writes the summaries into the pinned
        host buffer and the class posterior into its device buffer).  No copy or memset nodes."""
        par = self._par
        if self._particle_states.data_ptr() != self._S[par].data_ptr():
            self._S[par].copy_(self._particle_states)
        if self._particle_classes.data_ptr() != self._Cl[par].data_ptr():
            self._Cl[par].copy_(self._particle_classes)
        if self._step_dev_host != self._step:
            self._step_dev.fill_(self._step)
        src = np.asarray(z.detach().cpu() if isinstance(z, torch.Tensor) else z).reshape(-1)
        if src.size != self.observation_dim:
            raise ValueError("z must have D = %d entries" % self.observation_dim)
        self._graph_done.synchronize()  # the previous replay has consumed / produced the pinned buffers
        np.copyto(self._z_graph_np, src, casting="same_kind")
        if self._graphs[par] is None:  # every kernel has run at least once (function attributes are set): capture
            io = _cabi.PfSmallIo(z_src=self._z_graph_pin.data_ptr(), summary_dst=self._summary_pin.data_ptr(),
                                 probs_dst=ptr(self._probs[par]), frame=None)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._issue_small(par, True, self._E, self._eps, self._u, io)
            self._graphs[par] = g
        self._graphs[par].replay()
        self._graph_done.record()
        self._finish_small(par)
        self._summary_host_step, self._probs_par = self._step, par

    def _replay_many(self, Z):
        """update_many for a small cloud: the frames sit in a persistent device buffer, a device frame counter selects the
        row the pre kernel reads and the row of the output the post kernel writes, and the T steps are T replays of the
        same six-kernel graph (one per buffer parity) with no host synchronisation in between."""
        T, ncol = Z.shape[0], self.num_classes + self.latent_dim + 1
        if getattr(self, "_many_cap", 0) < T:  # (re)allocate the frame / output buffers; graphs hold their addresses
            self._many_cap = max(T, 256)
            self._many_z = torch.empty(self._many_cap, self.observation_dim, dtype=torch.float64, device=self.device)
            self._many_out = torch.empty(self._many_cap, ncol, dtype=torch.float64, device=self.device)
            self._many_frame = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._many_graphs = [None, None]
        self._many_z[:T].copy_(Z)
        self._many_frame.zero_()
        for t in range(T):
            par = self._par
            if self._small_steps < 2:  # every kernel has to have run once outside a capture
                self._update(self._many_z[t])
                self._many_out[t].copy_(self._summaries())
                self._many_frame.add_(1)
                continue
            if self._particle_states.data_ptr() != self._S[par].data_ptr():
                self._S[par].copy_(self._particle_states)
            if self._particle_classes.data_ptr() != self._Cl[par].data_ptr():
                self._Cl[par].copy_(self._particle_classes)
            if self._step_dev_host != self._step:
                self._step_dev.fill_(self._step)
            if self._many_graphs[par] is None:
                io = _cabi.PfSmallIo(z_src=ptr(self._many_z), summary_dst=ptr(self._many_out), probs_dst=None,
                                     frame=ptr(self._many_frame))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._issue_small(par, True, self._E, self._eps, self._u, io)
                self._many_graphs[par] = g
            self._many_graphs[par].replay()
            self._finish_small(par)
        return self._many_out[:T].clone()

    def _finish_small(self, par):
        self._small_steps += 1
        self._step += 1
        self._step_dev_host = self._step
        self._particle_states, self._particle_classes = self._S[1 - par], self._Cl[1 - par]
        self._log_likelihoods, self._log_weights, self._weights = self._LL[par], self._LW[par], self._W[par]
        self._summary_step = self._step  # the post kernel has already written the summaries of this state
        self._par = 1 - par

    def _summary_host(self):
        """The summaries on the host: already there after a graph replay (pinned buffer written by the graph's last node)."""
        if self._small and self._summary_host_step == self._step:
            self._graph_done.synchronize()
            return self._summary_np
        return self._summaries().cpu().numpy()

    @property
    def launches_per_step(self) -> int:
        """Kernels of libgpmdm_sm100a.so launched by one update() + one query (device-draw mode)."""
        if self._small:
            return 6  # pre, propagate (items + finalise), observe (items + finalise), post -- the query is a read
        draws, transition, bucket, propagate, normalize, resample, summaries = 2, 1, 3, 1, 5, 1, 4
        observe = 2 if (self._precision != "fp64" or self._lowlat) else 1
        propagate = 2 if self._lowlat else 1
        cdf = 2 if self._cdf_mode == 0 else 3
        return draws + transition + bucket + propagate + observe + normalize + cdf + resample + summaries

    def _log_weights_buf(self):
        return torch.empty(self._num_particles, dtype=torch.float64, device=self.device)

    _weights_buf = _log_weights_buf

    # ---- queries (gpmdm_pf.py:215-262) ---------------------------------------------------------------------------
    def _summaries(self):
        if self._summary_step != self._step:
            check(self._lib.gpmdm_pf_summaries_f64(ptr(self._log_likelihoods), ptr(self._log_weights), ptr(self._weights),
                                                   ptr(self._particle_classes), ptr(self._particle_states),
                                                   self._num_particles, self.num_classes, self.latent_dim,
                                                   ptr(self._summary), ptr(self._ws), stream()), "gpmdm_pf_summaries_f64")
            self._summary_step = self._step
        return self._summary

    def log_likelihood(self) -> float:
        return float(self._summary_host()[self.num_classes + self.latent_dim])

    def class_probabilities(self):
        if self._small and self._summary_host_step == self._step:
            return self._probs[self._probs_par]  # written by the graph; stays valid until the step after next
        return self._summaries()[:self.num_classes].clone()

    def get_most_likely_class(self) -> int:
        # one device-to-host read of the class posteriors; argmax on the host (first maximum, as torch.argmax)
        return int(np.argmax(self._summary_host()[:self.num_classes]))

    def current_state_mean(self):
        C = self.num_classes
        return self._summaries()[C:C + self.latent_dim].to(self.dtype, copy=True)

    def reset(self):
        self._init_particles()
        self._counter[2:].zero_()

    def variance_faults(self):
        """(dynamics, observation): particle updates since construction / reset() whose predictive variance was not a
        positive finite number.  The reference silently produces NaN states / log-likelihoods for them (sqrt at
        gpmdm_pf.py:168, log at :189) and so does this filter; the kernels' epilogues count them in a device word.
        Synchronises the stream."""
        dyn, obs = self._counter[2:4].tolist()
        return int(dyn), int(obs)

    # ---- stage outputs of the last step (for tests / diagnostics) ------------------------------------------------
    @property
    def last_ancestors(self):
        return self._anc

    @property
    def last_pre_resample_states(self):
        return self._x_new

    @property
    def last_pre_resample_classes(self):
        return self._c_new

    # ---- properties (gpmdm_pf.py:267-285) ----------------------------------------------------------------------
    @property
    def latent_dim(self):
        return self._gpmdm.d

    @property
    def observation_dim(self):
        return self._gpmdm.D

    @property
    def num_classes(self):
        return self._gpmdm.n_classes

    @property
    def dtype(self):
        return self._gpmdm.dtype

    @property
    def device(self):
        return self._gpmdm.device

    def _divide_into_n_parts(self, x: int, n: int):
        group, remainder = divmod(x, n)
        return [group + (1 if i < remainder else 0) for i in range(n)]
