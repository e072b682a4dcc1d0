// One filter step issued from native code: the launch sequence of GPMDM_PF._update (reference gpmdm/gpmdm_pf.py:126-135:
// _propogate_markov_switching, _propogate_dynamics, _update_weights, _resample) over the stage entry points of this
// library.  No kernels of its own; it exists because at the reference's own operating point (100 particles) a step is
// bound by launch latency and per-call FFI overhead, not by arithmetic.
#include "common.cuh"

using namespace gpmdm;

#define GPMDM_TRY(call)          \
    do {                         \
        const int rc_ = (call);  \
        if (rc_ != 0) return rc_; \
    } while (0)

extern "C" int gpmdm_pf_step_local_f64(const gpmdm_pf_step_args* a, void* stream) {
    GPMDM_REQUIRE(a && a->dyn && a->obs, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(a->P > 0 && a->lo >= 0 && a->n_local >= 0 && a->lo + a->n_local <= a->P, GPMDM_E_INVALID,
                  "bad particle range [%lld, %lld) of %lld", (long long)a->lo, (long long)(a->lo + a->n_local),
                  (long long)a->P);
    const int64_t P = a->P, lo = a->lo, n = a->n_local;
    const int32_t C = a->C, d = a->d;
    if (a->generate_draws) {
        GPMDM_TRY(gpmdm_pf_draws_philox(a->seed, a->step, lo, n, P, C, d, 0, a->E, a->eps, nullptr, stream));
        GPMDM_TRY(gpmdm_pf_draws_philox(a->seed, a->step, 0, P, P, C, d, a->systematic, nullptr, nullptr, a->u, stream));
    }
    if (n == 0) return 0;
    double* x_new = a->x_new + lo * d;
    int64_t* c_new = a->c_new + lo;
    double* ll = a->ll + lo;
    GPMDM_TRY(gpmdm_pf_transition_f64(a->c_prev, a->T, a->E, n, C, c_new, stream));
    GPMDM_TRY(gpmdm_pf_bucket_by_class(c_new, n, C, a->perm, a->tiles, a->n_tiles, a->workspace, stream));
    if (a->predict_mode == 2) {
        GPMDM_TRY(gpmdm_pf_propagate_lowlat_f64(a->dyn, a->x_prev, a->perm, a->tiles, a->n_tiles, n, a->eps, x_new, nullptr,
                                                nullptr, a->dyn_max_n_pad, a->dyn_seg_chunks, a->tile_counter, a->lowlat_workspace, stream));
        GPMDM_TRY(gpmdm_pf_observe_lowlat_f64(a->obs, x_new, n, a->z, a->ll_const, nullptr, ll, nullptr, nullptr,
                                              a->obs_n_pad, a->obs_seg_chunks, a->tile_counter, a->lowlat_workspace, stream));
    } else {
        // the K* cache pays from ~4 column panels per class block; its scratch is shared with the observation call
        if (a->predict_mode == 1 && a->dyn_max_n_pad >= 4 * GPMDM_TILE_N &&
            (int64_t)a->dyn_max_n_pad <= a->obs_n_pad)
            GPMDM_TRY(gpmdm_pf_propagate_cached_f64(a->dyn, a->x_prev, a->perm, a->tiles, a->n_tiles, n, a->eps, x_new, nullptr,
                                                    nullptr, a->dyn_max_n_pad, a->tile_counter, a->kstar_workspace,
                                                    a->kstar_workspace_bytes, stream));
        else
            GPMDM_TRY(gpmdm_pf_propagate_f64(a->dyn, a->x_prev, a->perm, a->tiles, a->n_tiles, n, a->eps, x_new, nullptr, nullptr,
                                             a->tile_counter, stream));
        if (a->predict_mode == 1)
            GPMDM_TRY(gpmdm_pf_observe_cached_f64(a->obs, x_new, n, a->z, a->ll_const, ll, nullptr, nullptr, a->obs_n_pad,
                                                  a->tile_counter, a->kstar_workspace, a->kstar_workspace_bytes, stream));
        else
            GPMDM_TRY(gpmdm_pf_observe_f64(a->obs, x_new, n, a->z, a->ll_const, ll, nullptr, nullptr, a->tile_counter,
                                           stream));
    }
    return 0;
}

extern "C" int gpmdm_pf_step_global_f64(const gpmdm_pf_step_args* a, void* stream) {
    GPMDM_REQUIRE(a, GPMDM_E_INVALID, "null argument");
    const int64_t P = a->P;
    GPMDM_TRY(gpmdm_pf_normalize_f64(a->ll, P, a->lw, a->w, a->stats, a->workspace, stream));
    GPMDM_TRY(gpmdm_pf_cdf_f64(a->w, P, a->cdf_mode, a->cdf, a->workspace, stream));
    // device-generated systematic draws are an ascending comb; injected draws may be in any order
    if (a->systematic && a->generate_draws)
        GPMDM_TRY(gpmdm_pf_resample_sorted_f64(a->cdf, P, a->u, P, a->x_new, a->c_new, a->d, a->anc, a->x_out, a->c_out,
                                               stream));
    else
        GPMDM_TRY(gpmdm_pf_resample_f64(a->cdf, P, a->u, P, a->x_new, a->c_new, a->d, a->anc, a->x_out, a->c_out, stream));
    return 0;
}
