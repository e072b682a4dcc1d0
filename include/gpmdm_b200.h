/*
 * gpmdm_b200.h -- C ABI of libgpmdm_sm100a.so: the GPMDM particle-filter step and the GP kernel
 * machinery beneath it, as hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * The reference (Priyanshu4/gpmdm) is pure Python on torch and has no FFI; the boundary these entry
 * points replace is the body of the Python methods cited on each function (paths relative to the
 * reference root).  The host side (gpmdm_b200/gpmdm_pf.py, gpmdm_b200/gpmdm.py) keeps the reference's
 * class surface and binds this library with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; arrays are dense row-major;
 *     sizes are int64_t; floating point is IEEE binary64; class ids / ancestors are int64 (as the
 *     reference's torch.int64 tensors, gpmdm_pf.py:101,211).
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); the library never
 *     synchronises, allocates or keeps references to caller memory.  Scratch comes from the caller
 *     (`workspace`, size from gpmdm_workspace_bytes()).
 *   - return value: 0 ok; <0 invalid argument (GPMDM_E_*); >0 a cudaError_t.  gpmdm_last_error()
 *     returns a thread-local message for the last non-zero return.  No exceptions cross the ABI.
 *   - there is no CPU implementation behind any entry point.
 */
#ifndef GPMDM_B200_H_
#define GPMDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPMDM_ABI_VERSION 4

#define GPMDM_E_INVALID (-1)     /* bad size / null pointer / unsupported dimension            */
#define GPMDM_E_UNSUPPORTED (-2) /* valid request outside this build's limits (d > 8, ...)      */
#define GPMDM_E_NODEVICE (-3)    /* no sm_100 device                                            */

#define GPMDM_TILE_P 64    /* particles per tile of the predict kernels                        */
#define GPMDM_TILE_N 256   /* column-tile width: row padding of factors, granularity of alpha_ld */
#define GPMDM_PANEL_LD 260 /* row pitch (doubles) of the L / alpha column panels = shared-memory pitch  */
#define GPMDM_MAX_LATENT 8 /* latent dimension limit of the kernels                             */

/* One GP "block": the observation GP is a single block over all N training frames; the dynamics GP
 * has one block per class over that class' N_c (x_t -> x_t+1) pairs -- the block-diagonal structure
 * the reference expresses with dense 0/1 masks (gpmdm.py:311-378, 1299-1305).
 *
 *   coords [n_pad, rec]  per training row i:  a_i[0..d) , (dynamics only) c_k^2 * x_i[k] for k in [0, d), then
 *                        zero padding to an even width: rec = (d + 1) & ~1 (kind 0), (2d + 1) & ~1 (kind 1);
 *                        a_i = x_i / lengthscale  (gpmdm.py:508-517, 545-548); zero rows pad.
 *   L      quadratic-form matrix Q such that  k^T K^-1 k == k^T Q k :
 *                        tri = 1:  Q[i][j] = Kinv[i][j] + Kinv[j][i] (i > j), Kinv[i][i], 0 (i < j)
 *                        tri = 0:  Q = Kinv.   Zero padded to n_pad.
 *                        Stored as COLUMN PANELS in the order the kernel streams them: panel t holds columns
 *                        [256 t, 256 t + 256) of the rows k >= kb(t) (kb = 256 t for tri = 1 -- the rows above are
 *                        zero and are neither stored nor read -- and 0 for tri = 0), each row padded to
 *                        GPMDM_PANEL_LD = 260 doubles (the shared-memory row pitch that makes the tensor-core
 *                        fragment loads bank-conflict free), panels back to back.  Any 16 consecutive rows of a
 *                        panel are one contiguous 33 280-byte block = one TMA bulk copy.  tri = 1 takes
 *                        ~half the memory of the dense matrix (1.68 GB instead of 3.27 GB at N = 20 000).
 *                        Size: gpmdm_quadform_bytes(n_pad, tri); written by gpmdm_pack_quadform_f64.
 *   alpha  = Kinv^T * targets (Y for the observation GP, Xout_c for dynamics), zero padded to n_pad rows and
 *                        alpha_ld = multiple of GPMDM_TILE_N columns, in the same panel layout:
 *                        [alpha_ld / 256][n_pad][260].  Size: gpmdm_alpha_bytes(n_pad, alpha_ld); written by
 *                        gpmdm_pack_alpha_f64.
 */
typedef struct gpmdm_gp_block {
    const double* coords;
    const double* L;
    const double* alpha;
    int64_t n;     /* real training rows of the block            */
    int64_t n_pad; /* rows padded to a multiple of GPMDM_TILE_N   */
} gpmdm_gp_block;

/* A GP model as consumed by the predict kernels.  `blocks` is a DEVICE array of n_blocks structs. */
typedef struct gpmdm_gp_model {
    const gpmdm_gp_block* blocks; /* device */
    int32_t n_blocks;
    int32_t d;        /* latent dimension (<= GPMDM_MAX_LATENT)                                   */
    int32_t dout;     /* D for the observation GP, d for the dynamics GP                          */
    int32_t alpha_ld; /* columns of alpha (multiple of GPMDM_TILE_N, >= dout)                       */
    int32_t kind;     /* 0 = RBF (gpmdm.py:381 get_y_kernel), 1 = RBF + linear (:408 get_x_kernel) */
    int32_t tri;      /* layout of L, see above                                                   */
    const double* lengthscales; /* device [d]   exp(log_lengthscales)                             */
    const double* lin_c2;       /* device [d+1] exp(x_log_lin_coeff)^2 (kind 1) or NULL           */
    const double* lambdas;      /* device [dout] observation GP: exp(y_log_lambdas)^2 ;
                                   dynamics GP: exp(x_log_lambdas)^-2 (gpmdm.py:960, 1066)         */
} gpmdm_gp_model;

int gpmdm_abi_version(void);
const char* gpmdm_last_error(void);

/* ---- factor packing ------------------------------------------------------------------------------
 * Kinv [n, n] (e.g. the reference's Ky_inv, or a diagonal block of Kx_inv_class[c],
 * gpmdm.py:1289,1305) -> the column panels of L described on gpmdm_gp_block. */
int64_t gpmdm_quadform_bytes(int64_t n_pad, int tri);
int gpmdm_pack_quadform_f64(const double* Kinv, int64_t n, int64_t n_pad, int tri, double* L, void* stream);
/* A [n, dout] row-major (= Kinv^T * targets: K_y^-1 Y of gpmdm.py:957, K_c^-1 Xout of :1064) -> alpha panels. */
int64_t gpmdm_alpha_bytes(int64_t n_pad, int32_t alpha_ld);
int gpmdm_pack_alpha_f64(const double* A, int64_t n, int64_t n_pad, int32_t dout, int32_t alpha_ld, double* alpha,
                         void* stream);

/* ---- filter step ------------------------------------------------------------------------------ */

/* GPMDM_PF._propogate_markov_switching (gpmdm_pf.py:137-151): c_new[p] = argmax_j T[c_prev[p]][j] / E[p][j]
 * where E are Exp(1) draws -- bit-exactly what torch.multinomial(dist, 1) computes from its uniforms
 * (E = -log1p(-U)).  T [C, C] row-major, not normalised (as the reference). */
int gpmdm_pf_transition_f64(const int64_t* c_prev, const double* T, const double* E, int64_t P, int32_t C,
                            int64_t* c_new, void* stream);

/* Groups particles by class so that every GPMDM_TILE_P-particle tile of the dynamics kernel is class
 * homogeneous (the reference's boolean-mask gather, gpmdm_pf.py:161).  Stable counting sort.
 *   perm  [P]            particle indices ordered by (class, index)
 *   tiles [P/64 + C, 4]  {block, first position in perm, count, 0} per tile, int32
 *   n_tiles [1]          int32
 * workspace: gpmdm_workspace_bytes(P, C). */
int gpmdm_pf_bucket_by_class(const int64_t* classes, int64_t P, int32_t C, int32_t* perm, int32_t* tiles,
                             int32_t* n_tiles, void* workspace, void* stream);

/* GPMDM_PF._propogate_dynamics + GPMDM.map_x_dynamics_for_class (gpmdm_pf.py:153-168,
 * gpmdm.py:1032-1068): for particle p of class c (block c of `dyn`)
 *     k_i  = exp(-|(xin_i - x_p)/l|^2) + [xin_i,1] diag(c^2) [x_p,1]^T          (never stored)
 *     mean = k^T alpha_c ; var_k = (1 + [x,1]diag(c^2)[x,1]^T - k^T L_c k) * lambda_k^-2
 *     x_new = eps * sqrt(var) + mean                                   (mul, then add -- :168)
 * perm/tiles/n_tiles from gpmdm_pf_bucket_by_class.  eps [P, d] is indexed by particle.
 * mean_out / var_out ([P, d], may be NULL) expose the GP prediction (map_x_dynamics_for_class).
 * x_new may be NULL when only the prediction is wanted.
 * tile_counter: device int32 [4], shared by all predict entry points.  [0], [1]: tile hand-out and round-synchronisation
 * counters, zeroed by every call.  [2], [3]: STATUS words, only ever incremented by the library (the caller zeroes them
 * when it wants a fresh count): [2] += particles whose dynamics predictive variance was not a positive finite number,
 * [3] += the same for the observation variance.  The results for such particles are what the reference produces --
 * NaN from sqrt / log of the variance (gpmdm_pf.py:168, :189) -- the words make the event observable. */
int gpmdm_pf_propagate_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                           const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                           double* x_new, double* mean_out, double* var_out, int32_t* tile_counter,
                           void* stream);

/* GPMDM_PF._update_weights (first half) + GPMDM.map_x_to_y (gpmdm_pf.py:170-192, gpmdm.py:923-963):
 *     k_i = exp(-|(X_i - x_p)/l_y|^2);  mu = k^T alpha_y;  v = 1 - k^T L k;  var_j = v * lambda_j^-2
 *     ll_p = -(1/(2v)) sum_j lambda_j^2 (z_j - mu_j)^2 - D log v + 2 sum_j log lambda_j - c32
 * which is the reference's per-particle loop (:188-192) in closed form, including its double-counted
 * log-variance term.  ll_const = 2 sum_j log lambda_j - c32 is computed by the host, with
 * c32 = fp32(0.5*D*log(2 pi)) evaluated in float32 exactly as gpmdm_pf.py:5,191 does.
 * z [D] device.  ll [P] may be NULL; mu_out [P, D] and v_out [P] may be NULL. */
int gpmdm_pf_observe_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                         double* ll, double* mu_out, double* v_out, int32_t* tile_counter, void* stream);

/* The same call with a K* cache: each CTA evaluates the cross-kernel of its 64-particle tile against all n_pad training
 * rows once (the N x P_tile slice of get_y_kernel(X, X*), gpmdm.py:955) into its slice of `kstar_ws`, already in
 * tensor-core fragment order, and the contraction re-reads it for every column tile instead of re-evaluating the
 * exponentials.  Results are bit-identical to gpmdm_pf_observe_f64.  n_pad: padded size of the observation block;
 * kstar_ws: device scratch of gpmdm_pf_observe_kstar_workspace_bytes(n_pad) bytes (#SMs x n_pad x 64 doubles),
 * contents undefined before and after. */
int64_t gpmdm_pf_observe_kstar_workspace_bytes(int64_t n_pad);
int gpmdm_pf_observe_cached_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z,
                                double ll_const, double* ll, double* mu_out, double* v_out, int64_t n_pad,
                                int32_t* tile_counter, void* kstar_ws, int64_t kstar_ws_bytes, void* stream);

/* gpmdm_pf_propagate_f64 with the same per-CTA K* cache (max_n_pad = largest padded block of `dyn`; the scratch may be
 * the one passed to gpmdm_pf_observe_cached_f64, the two calls of a step run one after the other).  Bit-identical. */
int gpmdm_pf_propagate_cached_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                                  const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                                  double* x_new, double* mean_out, double* var_out, int64_t max_n_pad,
                                  int32_t* tile_counter, void* kstar_ws, int64_t kstar_ws_bytes, void* stream);

/* Mean and log-likelihood only, with the predictive variances v_in [P] supplied by the caller (the tf32 variant
 * computes them on tcgen05 tensor cores): the N x D mean contraction stays in fp64 because alpha = K^-1 Y cancels
 * heavily; it is 2ND of the 2N^2 + 2ND flops.  The blocks of `obs` may have L == NULL for this call. */
int gpmdm_pf_loglik_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                        const double* v_in, double* ll, double* mu_out, int32_t* tile_counter, void* stream);

/* Low-latency variants for small particle counts (fewer 64-particle tiles than SMs, e.g. the reference's README setup
 * with 100 particles): every (particle tile, 256-column tile, k segment) triple is a separate work item -- the k range
 * of a column tile is cut into (up to) 16 segments, a function of n_pad only -- and the per-item contributions to k^T L k and to
 * the means are added in a fixed order by a second kernel that also runs the epilogue.  Same results contract as the
 * fused calls; only the summation order over k differs (deterministic; independent of P and of the sharding).
 * max_n_pad: largest n_pad over the model's blocks.  seg_chunks: k-segment length in 16-row chunks, 0 = the default rule
 * (a function of max_n_pad only, so results do not depend on the batch); a caller that knows the whole cloud may pass
 * gpmdm_predict_lowlat_pick_segment(...) to fill the SMs in whole waves (results depend on the value only through the
 * summation order over k).  The cross-kernel K* of every particle tile is evaluated once, by a wide first kernel, into the
 * workspace and shared by all the tile's work items; warps of a ragged tile whose 8 particle rows lie beyond its count issue
 * no MMAs (100 particles cost 104 rows of work, not 128).
 * workspace: gpmdm_predict_lowlat_workspace_bytes(P, max_n_pad, dout, seg_chunks, n_blocks of the model). */
int64_t gpmdm_predict_lowlat_workspace_bytes(int64_t P, int64_t max_n_pad, int32_t dout, int32_t seg_chunks,
                                             int32_t n_blocks);
int32_t gpmdm_predict_lowlat_pick_segment(int64_t n_tiles, int64_t n_pad, int32_t alpha_ld, int32_t tri);
int gpmdm_pf_observe_lowlat_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                                const double* v_in, double* ll, double* mu_out, double* v_out, int64_t max_n_pad,
                                int32_t seg_chunks, int32_t* tile_counter, void* workspace, void* stream);
int gpmdm_pf_propagate_lowlat_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                                  const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                                  double* x_new, double* mean_out, double* var_out, int64_t max_n_pad,
                                  int32_t seg_chunks, int32_t* tile_counter, void* workspace, void* stream);

/* GPMDM_PF._update_weights (second half, gpmdm_pf.py:200-204): lw = ll - max(ll); w = exp(lw)/sum.
 * Reductions run in a fixed blocked order independent of the GPU count.  stats_out [2] = {max, sum}. */
int gpmdm_pf_normalize_f64(const double* ll, int64_t P, double* lw, double* w, double* stats_out,
                           void* workspace, void* stream);

/* cdf of torch.multinomial's CPU kernel (gpmdm_pf.py:211): running sum of w in index order, divided
 * by the total, last entry forced to 1.  mode 0: sequential single-thread order (bit-identical to the
 * reference's cumsum); mode 1: blocked parallel scan (fixed 1024-element blocks, deterministic). */
int gpmdm_pf_cdf_f64(const double* w, int64_t P, int32_t mode, double* cdf, void* workspace, void* stream);

/* GPMDM_PF._resample (gpmdm_pf.py:206-213): anc[s] = min{ j : cdf[j] >= u[s] }, then the gathers
 * x_out[s] = x_in[anc[s]], c_out[s] = c_in[anc[s]].  cdf/x_in/c_in span all P particles (all ranks);
 * u/anc/x_out/c_out span the n_out output slots of the calling rank. */
int gpmdm_pf_resample_f64(const double* cdf, int64_t P, const double* u, int64_t n_out, const double* x_in,
                          const int64_t* c_in, int32_t d, int64_t* anc, double* x_out, int64_t* c_out,
                          void* stream);

/* The same search for ASCENDING u (the systematic comb u[s] = (u0 + s) / P): one window of the cdf per 1024 outputs is
 * staged in shared memory, all HBM accesses are coalesced.  Identical ancestors to gpmdm_pf_resample_f64 on the same
 * inputs; undefined for u that is not ascending. */
int gpmdm_pf_resample_sorted_f64(const double* cdf, int64_t P, const double* u, int64_t n_out, const double* x_in,
                                 const int64_t* c_in, int32_t d, int64_t* anc, double* x_out, int64_t* c_out,
                                 void* stream);

/* GPMDM_PF.class_probabilities / current_state_mean / log_likelihood (gpmdm_pf.py:215-262), with the
 * reference's mix of post-resample classes/states and pre-resample weights:
 *   g = ll + lw - max(ll + lw);  class_prob[i] = sum_{c_post=i} exp(g) / sum exp(g)
 *   state_mean = sum_p x_post[p] * w[p];  loglik = sum exp(g)
 * out [C + d + 1] = {class_prob, state_mean, loglik}. */
int gpmdm_pf_summaries_f64(const double* ll, const double* lw, const double* w, const int64_t* c_post,
                           const double* x_post, int64_t P, int32_t C, int32_t d, double* out,
                           void* workspace, void* stream);

/* Device-side raw draws (throughput mode; parity mode injects host draws instead): Philox4x32-10
 * keyed by (seed, step) and counter = global particle index, so results do not depend on how
 * particles are sharded.  E [n, C] Exp(1); eps [n, d] N(0,1); u [n] U(0,1), or the systematic comb
 * u[s] = (u0 + first + s) / P_total when systematic != 0. */
int gpmdm_pf_draws_philox(uint64_t seed, uint64_t step, int64_t first, int64_t n, int64_t P_total, int32_t C,
                          int32_t d, int32_t systematic, double* E, double* eps, double* u, void* stream);

int64_t gpmdm_workspace_bytes(int64_t P, int32_t C);

/* ---- tf32 variant of the observation GP (BASELINE config 4; ~1e-4 relative accuracy) ---------------------------
 * tcgen05 tensor cores with TMEM accumulators, error-compensated tf32 products (3 MMAs per k-step), and the
 * WHITENED variance  v = 1 - |W k|^2,  W = U^-T  with K = U^T U the reference's upper Cholesky factor
 * (gpmdm.py:1287-1288): the explicit-inverse form of gpmdm.py:958-959 is not usable below fp64.
 * The dynamics GP, the draws and the resampling stay on the fp64 path.
 *   coords [n_pad/2, 8, 2] fp32  a_i = sqrt(log2 e) * x_i / lengthscale (so that exp(-|a-b|^2) is one ex2 of the squared
 *                      distance), zero padded to 8 coordinates, rows interleaved in pairs:
 *                      element [i/2][j][i%2] = a_i[j]  (one 64-bit load feeds the packed fp32x2 pipe)
 *   wtiles             tf32 hi/lo tiles of W in tensor-core operand order (gpmdm_pack_whitened_tf32)
 *   atiles             tf32 hi/lo tiles of alpha = K^-T Y                 (gpmdm_pack_alpha_tf32)            */
typedef struct gpmdm_gp_model_tf32 {
    const float* coords;
    const float* wtiles;
    const float* atiles;
    int64_t n;
    int64_t n_pad;              /* multiple of GPMDM_TILE_N                                        */
    int32_t d;
    int32_t dout;               /* <= 256                                                          */
    const double* lengthscales; /* device [d]                                                      */
    const double* lambdas;      /* device [dout] exp(y_log_lambdas)^2                              */
} gpmdm_gp_model_tf32;

int64_t gpmdm_tf32_wtiles_bytes(int64_t n_pad);
int64_t gpmdm_tf32_atiles_bytes(int64_t n_pad);
/* W [n, n] fp64 row-major, lower triangular (W = U^-T). */
int gpmdm_pack_whitened_tf32(const double* W, int64_t n, int64_t n_pad, float* wtiles, void* stream);
/* alpha [n, dout] fp64 row-major. */
int gpmdm_pack_alpha_tf32(const double* alpha, int64_t n, int64_t n_pad, int32_t dout, float* atiles, void* stream);
/* Same contract as gpmdm_pf_observe_f64 (x, z, ll, mu_out, v_out are fp64 arrays; tile_counter: int32 scratch [4],
 * may be NULL). */
int gpmdm_pf_observe_tf32(const gpmdm_gp_model_tf32* obs, const double* x, int64_t P, const double* z, double ll_const,
                          double* ll, double* mu_out, double* v_out, int32_t* tile_counter, void* stream);

/* ---- fp16-split variant of the same kernel (precision "f16x2"): every operand as TWO fp16 pieces, a = hi + 2^-11 lo
 * (the scaling keeps the residual inside fp16's exponent range; requires |W| < 6e4, i.e. noise std above ~2e-5), three
 * kind::f16 MMAs per k-step of 16 into two fp32 TMEM accumulators combined in the epilogue: the same ~22 significant
 * bits per product as the 3 x tf32 scheme at half the tensor-pipe time and half the operand bytes.  Variances only
 * (v_out [P]); means and log-likelihoods come from gpmdm_pf_loglik_f64 as in the hybrid tf32 variant.
 * `obs->wtiles` must point at tiles written by gpmdm_pack_whitened_f16x2 (gpmdm_f16_wtiles_bytes(n_pad) bytes);
 * `obs->atiles` is not used. */
int64_t gpmdm_f16_wtiles_bytes(int64_t n_pad);
int gpmdm_pack_whitened_f16x2(const double* W, int64_t n, int64_t n_pad, void* wtiles, void* stream);
int gpmdm_pf_observe_f16x2(const gpmdm_gp_model_tf32* obs, const double* x, int64_t P, double* v_out,
                           int32_t* tile_counter, void* stream);

/* ---- dynamics GP variance on the tensor cores (precision "tf32" / "f16x2"; gpmdm.py:1032-1068) ----------------------
 * The class block K_c (incl. the 1e-6 jitter, gpmdm.py:1301-1303) = L_c L_c^T, W_c = L_c^-1, and the cross-kernel is
 * k = k_rbf + X~ S x~ (X~ = [X_in, 1], S = diag(c^2), x~ = [x, 1]; gpmdm.py:545-548), so
 *     var = prior - |W_c k|^2 = (1 - |W_c k_rbf|^2)  +  x~^T S x~ - 2 (k^T G_c) x~ + x~^T H_c x~,
 *     G_c = K_c^-1 X~ S  [N_c, d+1],   H_c = S X~^T G_c  [d+1, d+1].
 * Only the first term needs O(N_c^2) work: it runs on tcgen05 with entries k_rbf <= 1 (gpmdm_pf_dynvar_tc, the
 * observation kernel per class block, particles reached through gpmdm_pf_bucket_by_class2's permutation in
 * class-homogeneous tiles of 128).  The low-rank remainder is exact fp64: gpmdm_pf_propagate_meanonly_f64 contracts k with
 * the alpha tile [alpha_c | G_c] (2 N_c (2d+1) flops) and finishes the variance, the mean and the draw.  Keeping the linear
 * kernel (entries ~ |x|^2 c^2, all of one sign) off the tensor cores is what keeps their rounding error relative to 1
 * instead of to the prior.
 *   coords [n_pad/2, 8, 2] fp32: sqrt(log2 e) x_i / lengthscale, rows interleaved in pairs as for gpmdm_gp_model_tf32;
 *   wtiles from gpmdm_pack_whitened_tf32 (mode 0) / gpmdm_pack_whitened_f16x2 (mode 1). */
typedef struct gpmdm_tc_block {
    const float* coords;
    const void* wtiles;
    int64_t n;
    int64_t n_pad;
} gpmdm_tc_block;

/* gpmdm_pf_bucket_by_class plus a second tile table with 128-particle tiles (tiles128 [P/128 + C, 4], n_tiles128 [1]). */
int gpmdm_pf_bucket_by_class2(const int64_t* classes, int64_t P, int32_t C, int32_t* perm, int32_t* tiles,
                              int32_t* n_tiles, int32_t* tiles128, int32_t* n_tiles128, void* workspace, void* stream);
/* u_out [P] (indexed by particle) = 1 - |W_c k_rbf|^2.  blocks: DEVICE array of n_blocks structs; mode 0 = tf32 x 3,
 * 1 = fp16 split. */
int gpmdm_pf_dynvar_tc(const gpmdm_tc_block* blocks, int32_t n_blocks, int32_t d, int32_t mode, const double* lengthscales,
                       const double* x_prev, const int32_t* perm, const int32_t* tiles128, const int32_t* n_tiles128,
                       int64_t P, double* u_out, void* stream);
/* gpmdm_pf_propagate_f64 with the O(N_c^2) part of the variance supplied: u_in [P] from gpmdm_pf_dynvar_tc.  `dyn` is a
 * dynamics model whose alpha matrices carry d + 1 extra columns G_c after the d of alpha_c; lowrank_h = DEVICE
 * [n_blocks, d+1, d+1] doubles (H_c, row major).  Only the alpha tile of each class block is contracted; then
 * x_new = eps * sqrt(var lambda^-2) + mean.  tile_counter word [2] counts non-positive variances. */
int gpmdm_pf_propagate_meanonly_f64(const gpmdm_gp_model* dyn, const double* lowrank_h, const double* x_prev,
                                    const int32_t* perm, const int32_t* tiles, const int32_t* n_tiles, int64_t P,
                                    const double* eps, const double* u_in, double* x_new, double* mean_out,
                                    double* var_out, int32_t* tile_counter, void* stream);

/* ---- one filter step as two calls (GPMDM_PF._update, gpmdm_pf.py:126-135) ----------------------------------
 * The host-side sequence of the stage entry points above, issued from native code so that a step costs two FFI calls
 * instead of ~twelve (with 100 particles the step is launch-latency bound).  `local` = draws, transition, bucketing,
 * dynamics draw, observation log-likelihood for this rank's particles [lo, lo + n_local); then -- multi-GPU only --
 * the caller all-gathers (x_new, c_new, ll); `global` = normalise, cdf, resample over all P particles.
 * All pointers are device pointers except the struct itself; nothing is retained after the call. */
typedef struct gpmdm_pf_step_args {
    const gpmdm_gp_model* dyn;  /* host structs (device block tables inside)                              */
    const gpmdm_gp_model* obs;
    int64_t P, lo, n_local;     /* all particles; this rank's range                                         */
    int32_t C, d;
    int32_t generate_draws;     /* 1: Philox draws into E/eps/u (seed, step); 0: E/eps/u hold injected draws */
    int32_t systematic;         /* resampling comb instead of P independent uniforms                        */
    int32_t cdf_mode;           /* gpmdm_pf_cdf_f64 mode                                                    */
    int32_t predict_mode;       /* 0 fused, 1 fused with K* cache, 2 low latency                            */
    uint64_t seed, step;
    const double* T;            /* [C, C]                                                                   */
    const double* z;            /* [D]                                                                      */
    double ll_const;
    const double* x_prev;       /* [n_local, d] this rank's slice of the particle states                    */
    const int64_t* c_prev;      /* [n_local]                                                                */
    double* E;                  /* [n_local, C]                                                             */
    double* eps;                /* [n_local, d]                                                             */
    double* u;                  /* [P]                                                                      */
    double* x_new;              /* [P, d]  pre-resample states; local rows written by `local`               */
    int64_t* c_new;             /* [P]                                                                      */
    double* ll;                 /* [P]                                                                      */
    int32_t* perm;              /* bucketing scratch, see gpmdm_pf_bucket_by_class                          */
    int32_t* tiles;
    int32_t* n_tiles;
    int32_t* tile_counter;
    void* workspace;            /* gpmdm_workspace_bytes(P, C)                                              */
    void* lowlat_workspace;     /* predict_mode 2                                                           */
    int64_t obs_n_pad, dyn_max_n_pad;
    int32_t obs_seg_chunks, dyn_seg_chunks; /* predict_mode 2: k-segment lengths (0 = default rule)                 */
    void* kstar_workspace;      /* predict_mode 1                                                           */
    int64_t kstar_workspace_bytes;
    double* lw;                 /* [P] outputs of `global`                                                  */
    double* w;
    double* stats;              /* [2] {max ll, sum exp}                                                    */
    double* cdf;
    int64_t* anc;
    double* x_out;              /* [P, d] resampled states                                                  */
    int64_t* c_out;             /* [P]                                                                      */
} gpmdm_pf_step_args;

int gpmdm_pf_step_local_f64(const gpmdm_pf_step_args* a, void* stream);
int gpmdm_pf_step_global_f64(const gpmdm_pf_step_args* a, void* stream);

/* The whole step for a SMALL cloud on one rank (1 <= P <= gpmdm_pf_small_max_particles() = 4096, lo = 0, n_local = P,
 * predict_mode = 2) -- the reference's own operating point is 100 particles (notebooks/test_gpmdm_pf.ipynb cell 3), where a
 * step is bound by launch latency: draws + transition + bucketing run as ONE single-CTA kernel, normalise + cdf +
 * resampling + the class / state summaries as another, around the two low-latency predict calls: 6 kernels instead of ~28.
 * Same device functions and the same fixed reduction order as the staged entry points: bit-identical results.
 *   step_dev  device uint64 (may be NULL): when given, the Philox step key is read from it instead of a->step and it is
 *             advanced by one at the end of the step, so the launch sequence has no per-step host parameter and can be
 *             captured once into a CUDA graph and replayed per frame (gpmdm_b200/gpmdm_pf.py does).
 *   summary   device [C + d + 1] (may be NULL): the outputs of gpmdm_pf_summaries_f64 for the post-step state.
 *   io        (may be NULL) observation source / result destinations handled INSIDE the step's first and last kernel, so
 *             that a frame is six kernel nodes and nothing else (no memset, no copy nodes):
 *               z_src        frames [*, D], device memory or page-locked host memory (read through the mapping): frame
 *                            number *frame (0 if frame is NULL) is copied into a->z by the first kernel
 *               summary_dst  [*, C + d + 1], device or page-locked host memory: row *frame receives the summary
 *               probs_dst    device [C]: copy of the class posterior
 *               frame        device uint64 frame counter, advanced by one at the end of the step
 *             (summary must be non-NULL when summary_dst or probs_dst is given). */
typedef struct gpmdm_pf_small_io {
    const double* z_src;
    double* summary_dst;
    double* probs_dst;
    uint64_t* frame;
} gpmdm_pf_small_io;
int32_t gpmdm_pf_small_max_particles(void);
int gpmdm_pf_step_small_f64(const gpmdm_pf_step_args* a, uint64_t* step_dev, double* summary, const gpmdm_pf_small_io* io,
                            void* stream);

/* ---- training-side kernel matrices (gpmdm.py:381-548, 311-340, 550-628) ------------------------
 * K = exp(-|(x_i-x_j)/l|^2) [+ [x_i,1]diag(c^2)[x_j,1]^T if kind 1] [+ noise2 on the diagonal],
 * multiplied by the class-block mask given as row offsets (class_offsets [n_classes+1], device int64;
 * NULL = no mask) -- `get_x_kernel(Xin,Xin) * self.M` without the dense M. */
int gpmdm_kernel_build_f64(const double* X, int64_t n, int32_t d, int32_t kind, const double* lengthscales,
                           const double* lin_c2, double noise2, const int64_t* class_offsets, int32_t n_classes,
                           double* K, void* stream);

/* Gradient terms of a scalar loss through the kernel build above, given G = dLoss/dK [n, n]
 * (SURVEY.md App. A.5): gX [n, d], g_log_ls [d], g_log_sigma [1], g_log_c [d+1] (kind 1).
 * Deterministic two-stage reduction; workspace >= gpmdm_kernel_grad_workspace_bytes(n, d). */
int gpmdm_kernel_grad_f64(const double* X, const double* G, int64_t n, int32_t d, int32_t kind,
                          const double* lengthscales, const double* lin_c2, double sigma2,
                          const int64_t* class_offsets, int32_t n_classes, double* gX, double* g_log_ls,
                          double* g_log_sigma, double* g_log_c, void* workspace, void* stream);
int64_t gpmdm_kernel_grad_workspace_bytes(int64_t n, int32_t d);

/* ---- measurement helper: sustained fp64 mma.sync (DMMA m8n8k4) rate of this device, TFLOP/s, the
 * denominator of the contraction roofline (bench.py).  Synchronous. */
int gpmdm_probe_dmma_tflops(int32_t iters, double* tflops_host);
/* The same for the tf32 variant: sustained tcgen05.mma kind::tf32 (M = 128, N = 256, K = 8, operands resident in shared
 * memory, pseudo-random values) rate of this device, TFLOP/s.  Synchronous. */
int gpmdm_probe_tf32_tflops(int32_t iters, double* tflops_host);
/* ... and for kind::f16 (M = 128, N = 256, K = 16). */
int gpmdm_probe_f16_tflops(int32_t iters, double* tflops_host);

#ifdef __cplusplus
}
#endif
#endif /* GPMDM_B200_H_ */
