// The filter step for SMALL particle clouds (P <= 4096, one GPU) -- the reference's own operating point is 100
// particles (README; notebooks/test_gpmdm_pf.ipynb cell 3) -- where a step is bound by launch latency, not arithmetic:
// the staged sequence is ~28 dependent launches of a few microseconds of work each.
//
//   pre   (1 CTA)  raw draws + class transition + class bucketing     gpmdm_pf.py:137-151, :161   (6 launches -> 1)
//   GP             gpmdm_pf_propagate_lowlat_f64, gpmdm_pf_observe_lowlat_f64 (unchanged)         gpmdm_pf.py:153-192
//   post  (1 CTA)  normalise + cdf + resampling search / gather + class / state summaries
//                                                                      gpmdm_pf.py:200-262         (15 launches -> 1)
//
// Both kernels are built from the device functions the staged kernels use (pf_stages.cuh), walk the same 1024-element
// blocks in the same order and therefore produce bit-identical results (tested against the staged path).  The step
// counter that keys the Philox draws is read from device memory and advanced by the post kernel, so the whole step is a
// fixed launch sequence: the host captures it once into a CUDA graph and replays it per frame.
#include "common.cuh"
#include "pf_stages.cuh"

namespace gpmdm {

#ifdef GPMDM_TIMELINE  /* diagnostic build (tools/lowlat_timeline.py --frame): start / end of the pre and post kernels */
__device__ unsigned long long g_small_stamps[4];
__device__ __forceinline__ unsigned long long small_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GPMDM_SMALL_STAMP(i) if (threadIdx.x == 0) g_small_stamps[i] = small_now();
#else
#define GPMDM_SMALL_STAMP(i)
#endif

constexpr int SMALL_P_MAX = 4096;  // 4 reduction blocks; one CTA of 1024 threads holds 4 particles per thread
constexpr int PRE_T = 1024;

struct SmallPreArgs {
    unsigned long long seed, step;
    const unsigned long long* step_dev;  // device step counter (overrides `step` when non-null)
    int P, C, d, generate, systematic;
    const double* T;
    const int64_t* c_prev;
    double *E, *eps, *u;
    int64_t* c_new;
    int32_t *perm, *tiles, *n_tiles;
    int32_t* tile_counter;               // words [0..1] zeroed for the low-latency launches that follow
    const double* z_src;                 // frames [*, D] (device or mapped host memory) or NULL
    const unsigned long long* frame;     // device frame counter or NULL (= frame 0)
    double* z;                           // [D] device: where the observation kernels read the frame
    int D;
};

__global__ void __launch_bounds__(PRE_T, 1) small_pre_kernel(const SmallPreArgs a) {
    extern __shared__ double sT[];                    // [C][C]
    __shared__ int cnt[64], cls_start[65], tile_start[65], base[64];
    __shared__ int wc[32 * 64];                       // per-warp class counts of the current 1024-particle sub-block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, C = a.C;
    GPMDM_SMALL_STAMP(0)
    const unsigned long long step = a.step_dev ? *a.step_dev : a.step;
    if (a.z_src) {  // issued first: a read of mapped host memory is one PCIe round trip, hidden behind the rest
        const double* src = a.z_src + (a.frame ? (long long)*a.frame : 0ll) * a.D;
        for (int i = tid; i < a.D; i += PRE_T) a.z[i] = src[i];
    }
    if (tid < 2) a.tile_counter[tid] = 0;
    for (int i = tid; i < C * C; i += PRE_T) sT[i] = a.T[i];
    if (tid < 64) cnt[tid] = base[tid] = 0;
    if (a.generate)
        for (int i = tid; i < P; i += PRE_T)
            philox_draw_one(a.seed, step, (unsigned long long)i, i, P, C, a.d, a.systematic, a.E, a.eps, a.u);
    __syncthreads();  // this CTA's own global writes (E) are visible to all its threads past the barrier
    // class transition (gpmdm_pf.py:137-151); thread t owns particles t, t + 1024, ...
    int myc[SMALL_P_MAX / PRE_T];
#pragma unroll
    for (int sub = 0; sub < SMALL_P_MAX / PRE_T; sub++) {
        const int p = sub * PRE_T + tid;
        myc[sub] = -1;
        if (p < P) {
            myc[sub] = transition_one(sT + a.c_prev[p] * C, a.E + (long long)p * C, C);
            a.c_new[p] = myc[sub];
        }
    }
    // stable counting sort by class (gpmdm_pf.py:161): class counts ...
#pragma unroll
    for (int sub = 0; sub < SMALL_P_MAX / PRE_T; sub++) {
        const unsigned peers = __match_any_sync(0xffffffffu, myc[sub]);
        if (myc[sub] >= 0 && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&cnt[myc[sub]], __popc(peers));
    }
    __syncthreads();
    // ... class starts and the class-homogeneous tile table (same rule as bucket_scan_kernel) ...
    if (tid == 0) {
        int run = 0, t = 0;
        for (int c = 0; c < C; c++) {
            cls_start[c] = run;
            tile_start[c] = t;
            run += cnt[c];
            t += (cnt[c] + GPMDM_TILE_P - 1) / GPMDM_TILE_P;
        }
        cls_start[C] = run;
        tile_start[C] = t;
        a.n_tiles[0] = t;
    }
    __syncthreads();
    for (int c = 0; c < C; c++) {
        const int nt = tile_start[c + 1] - tile_start[c];
        const int n_c = cls_start[c + 1] - cls_start[c];
        for (int t = tid; t < nt; t += PRE_T) {
            int32_t* dst = a.tiles + 4ll * (tile_start[c] + t);
            dst[0] = c;
            dst[1] = cls_start[c] + t * GPMDM_TILE_P;
            dst[2] = min(GPMDM_TILE_P, n_c - t * GPMDM_TILE_P);
            dst[3] = 0;
        }
    }
    // ... and the stable scatter: rank inside the warp, exclusive scan of the warp counts, running base per class
#pragma unroll
    for (int sub = 0; sub < SMALL_P_MAX / PRE_T; sub++) {
        if (sub * PRE_T >= P) break;  // uniform
        for (int i = tid; i < 32 * C; i += PRE_T) wc[i] = 0;
        __syncthreads();
        const int c = myc[sub];
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (c >= 0 && rank == 0) wc[warp * C + c] = __popc(peers);
        __syncthreads();
        for (int k = tid; k < C; k += PRE_T) {
            int run = base[k];
            for (int w = 0; w < 32; w++) {
                const int v = wc[w * C + k];
                wc[w * C + k] = run;
                run += v;
            }
            base[k] = run;
        }
        __syncthreads();
        if (c >= 0) a.perm[cls_start[c] + wc[warp * C + c] + rank] = sub * PRE_T + tid;
        __syncthreads();
    }
    GPMDM_SMALL_STAMP(1)
}

struct SmallPostArgs {
    int P, C, d, cdf_mode;
    const double* ll;       // [P]
    const double* u;        // [P]
    const double* x_new;    // [P, d]
    const int64_t* c_new;   // [P]
    double *lw, *w, *cdf, *stats;
    int64_t* anc;
    double* x_out;
    int64_t* c_out;
    double* summary;        // [C + d + 1] or NULL
    double* ws;             // gpmdm_workspace_bytes(P, C)
    unsigned long long* step_dev;  // advanced by one when non-null
    double* summary_dst;    // [*, C + d + 1] device or mapped host memory, row *frame; or NULL
    double* probs_dst;      // [C] device or NULL
    unsigned long long* frame;  // advanced by one when non-null
};

__global__ void __launch_bounds__(RT, 1) small_post_kernel(const SmallPostArgs a) {
    __shared__ double sh[RT / 32 + 1];
    __shared__ double wsum[RT / 32];
    __shared__ double wtot32[32];
    __shared__ double carry;
    __shared__ double wtot[RT / 32][SUMM_COLS];
    __shared__ double cls[SUMM_COLS];
    __shared__ double sbuf[SMALL_P_MAX + 1];
    const int tid = threadIdx.x;
    const long long P = a.P;
    GPMDM_SMALL_STAMP(2)
    const int nb = (int)((P + RB - 1) / RB);
    double* scal = a.ws;                                                          // [0] max [1] sum [2] cdf total [3] max(ll + lw)
    // per-block partials (at most 4 blocks of 1024 particles) in shared memory: the staged kernels hand them from launch to
    // launch through global memory; inside one CTA every such hand-over would be an L2 round trip
    __shared__ double part_s[(SMALL_P_MAX / RB) * SUMM_COLS];
    double* part = part_s;
    // ---- lw = ll - max, w = exp(lw) / sum (gpmdm_pf.py:200-204) ----
    for (int vb = 0; vb < nb; vb++) {
        const double v = block_max_dev<false>(a.ll, nullptr, P, vb, sh);
        if (tid == 0) part[vb] = v;
    }
    __syncthreads();
    const double m = combine_partials_dev<true, false>(part, nb, sh);
    for (int vb = 0; vb < nb; vb++) {
        const double acc = exp_sum_dev<false>(a.ll, P, m, a.lw, a.w, vb, sh);
        if (tid == 0) part[vb] = acc;
    }
    __syncthreads();
    const double tot = combine_partials_dev<false, false>(part, nb, sh);
    if (tid == 0) {
        scal[0] = m;
        scal[1] = tot;
        if (a.stats) a.stats[0] = m, a.stats[1] = tot;
    }
    for (long long base = tid * 4; base < P; base += RT * 4) divide_dev<false>(a.w, P, tot, base);
    __syncthreads();
    // ---- cdf (gpmdm_pf.py:211: the running sum of torch.multinomial's CPU kernel, or the blocked scan) ----
    if (a.cdf_mode == 0) {
        // the reference's running sum is inherently serial: stage the weights in shared memory (coalesced), let one
        // thread add them there (an L2 round trip per element otherwise: 40 us at P = 100), write back in parallel
        for (long long i = tid; i < P; i += RT) sbuf[i] = __ldcg(a.w + i);
        __syncthreads();
        if (tid == 0) {
            double run = 0.0;
            long long i = 0;
            for (; i + 8 <= P; i += 8) {
                double v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = sbuf[i + k];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    run = __dadd_rn(run, v[k]);
                    sbuf[i + k] = run;
                }
            }
            for (; i < P; i++) {
                run = __dadd_rn(run, sbuf[i]);
                sbuf[i] = run;
            }
            scal[2] = run;
            sbuf[SMALL_P_MAX] = run;
        }
        __syncthreads();
        const double total = sbuf[SMALL_P_MAX];
        for (long long i = tid; i < P; i += RT) {
            const double c = (i == P - 1) ? 1.0 : sbuf[i] / total;  // as cdf_finish_dev
            a.cdf[i] = c;
            sbuf[i] = c;
        }
    } else {
        for (int vb = 0; vb < nb; vb++) {
            const double t = cdf_block_scan_dev<false>(a.w, P, a.cdf, vb, wsum);
            if (tid == RT - 1) part[vb] = t;
        }
        __syncthreads();
        const double total = scan_partials_dev<RT, false>(part, nb, wtot32, &carry);
        __syncthreads();
        for (long long base = tid * 4; base < P; base += RT * 4) cdf_finish_dev<false>(a.cdf, P, part, total, base);
        __syncthreads();
        for (long long i = tid; i < P; i += RT) sbuf[i] = __ldcg(a.cdf + i);
    }
    __syncthreads();
    // ---- ancestors + gathers (gpmdm_pf.py:206-213): first j with cdf[j] >= u, searched in the shared-memory copy ----
    for (long long s = tid; s < P; s += RT) {
        const double us = a.u[s];
        long long lo = 0, hi = P;
        while (hi - lo > 0) {
            const long long mid = lo + (hi - lo) / 2;
            if (sbuf[mid] < us) lo = mid + 1;
            else hi = mid;
        }
        const long long j = lo >= P ? P - 1 : lo;
        a.anc[s] = j;
        for (int k = 0; k < a.d; k++) a.x_out[s * a.d + k] = a.x_new[j * a.d + k];
        a.c_out[s] = a.c_new[j];
    }
    __syncthreads();
    // ---- class posteriors, state mean, likelihood sum (gpmdm_pf.py:215-262), eagerly: the query becomes a read ----
    if (a.summary) {
        for (int vb = 0; vb < nb; vb++) {
            const double v = block_max_dev<false>(a.ll, a.lw, P, vb, sh);
            if (tid == 0) part[vb] = v;
        }
        __syncthreads();
        const double gm = combine_partials_dev<true, false>(part, nb, sh);
        if (tid == 0) scal[3] = gm;
        __syncthreads();
        const int ncol = a.C + a.d + 1;
        for (int vb = 0; vb < nb; vb++)
            summaries_block_dev<false>(a.ll, a.lw, a.w, a.c_out, a.x_out, P, a.C, a.d, gm, part + (long long)vb * ncol, vb, wtot);
        summaries_final_dev<false>(part, nb, a.C, a.d, a.summary, sh, cls);
        if (a.summary_dst || a.probs_dst) {
            __syncthreads();  // a.summary was written by threads of this CTA
            const unsigned long long f = a.frame ? *a.frame : 0ull;
            for (int i = tid; i < ncol; i += RT) {
                const double v = __ldcg(a.summary + i);
                if (a.summary_dst) a.summary_dst[f * ncol + i] = v;
                if (a.probs_dst && i < a.C) a.probs_dst[i] = v;
            }
        }
    }
    __syncthreads();
    if (a.frame && tid == 0) a.frame[0] += 1ull;
    if (a.step_dev && tid == 0) a.step_dev[0] += 1ull;
    GPMDM_SMALL_STAMP(3)
}

}  // namespace gpmdm

using namespace gpmdm;

#define GPMDM_TRY(call)           \
    do {                          \
        const int rc_ = (call);   \
        if (rc_ != 0) return rc_; \
    } while (0)

extern "C" int32_t gpmdm_pf_small_max_particles(void) { return SMALL_P_MAX; }

extern "C" int gpmdm_pf_step_small_f64(const gpmdm_pf_step_args* a, uint64_t* step_dev, double* summary,
                                       const gpmdm_pf_small_io* io, void* stream) {
    GPMDM_REQUIRE(a && a->dyn && a->obs, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(!io || summary || !(io->summary_dst || io->probs_dst), GPMDM_E_INVALID,
                  "summary_dst / probs_dst need the summary buffer");
    GPMDM_REQUIRE(a->P > 0 && a->P <= SMALL_P_MAX && a->lo == 0 && a->n_local == a->P, GPMDM_E_UNSUPPORTED,
                  "the small-cloud step handles 1..%d particles on one rank (P = %lld, local %lld)", SMALL_P_MAX,
                  (long long)a->P, (long long)a->n_local);
    GPMDM_REQUIRE(a->C >= 1 && a->C <= 64 && a->d >= 1 && a->d <= GPMDM_MAX_LATENT, GPMDM_E_UNSUPPORTED,
                  "bad sizes C=%d d=%d", a->C, a->d);
    GPMDM_REQUIRE(a->predict_mode == 2 && a->lowlat_workspace, GPMDM_E_INVALID,
                  "the small-cloud step uses the low-latency predict mode");
    cudaStream_t st = (cudaStream_t)stream;
    const int P = (int)a->P;
    SmallPreArgs pre{};
    pre.seed = a->seed, pre.step = a->step, pre.step_dev = reinterpret_cast<const unsigned long long*>(step_dev);
    pre.P = P, pre.C = a->C, pre.d = a->d, pre.generate = a->generate_draws, pre.systematic = a->systematic;
    pre.T = a->T, pre.c_prev = a->c_prev, pre.E = a->E, pre.eps = a->eps, pre.u = a->u, pre.c_new = a->c_new;
    pre.perm = a->perm, pre.tiles = a->tiles, pre.n_tiles = a->n_tiles;
    pre.tile_counter = a->tile_counter;
    pre.z_src = io ? io->z_src : nullptr;
    pre.frame = io ? reinterpret_cast<const unsigned long long*>(io->frame) : nullptr;
    pre.z = const_cast<double*>(a->z), pre.D = a->obs->dout;
    small_pre_kernel<<<1, PRE_T, (size_t)a->C * a->C * sizeof(double), st>>>(pre);
    GPMDM_TRY(check_launch("small_pre_kernel"));
    // the hand-out counters are zeroed by the pre kernel and again by every finalize kernel: no memset nodes in between
    GPMDM_TRY(propagate_lowlat_impl(a->dyn, a->x_prev, a->perm, a->tiles, a->n_tiles, P, a->eps, a->x_new, nullptr, nullptr,
                                    a->dyn_max_n_pad, a->dyn_seg_chunks, a->tile_counter, a->lowlat_workspace, stream, true));
    GPMDM_TRY(observe_lowlat_impl(a->obs, a->x_new, P, a->z, a->ll_const, nullptr, a->ll, nullptr, nullptr, a->obs_n_pad,
                                  a->obs_seg_chunks, a->tile_counter, a->lowlat_workspace, stream, true));
    SmallPostArgs post{};
    post.P = P, post.C = a->C, post.d = a->d, post.cdf_mode = a->cdf_mode;
    post.ll = a->ll, post.u = a->u, post.x_new = a->x_new, post.c_new = a->c_new;
    post.lw = a->lw, post.w = a->w, post.cdf = a->cdf, post.stats = a->stats, post.anc = a->anc;
    post.x_out = a->x_out, post.c_out = a->c_out, post.summary = summary;
    post.ws = static_cast<double*>(a->workspace);
    post.step_dev = reinterpret_cast<unsigned long long*>(step_dev);
    post.summary_dst = io ? io->summary_dst : nullptr;
    post.probs_dst = io ? io->probs_dst : nullptr;
    post.frame = io ? reinterpret_cast<unsigned long long*>(io->frame) : nullptr;
    small_post_kernel<<<1, RT, 0, st>>>(post);
    return check_launch("small_post_kernel");
}

#ifdef GPMDM_TIMELINE
extern "C" int gpmdm_debug_small_stamps(unsigned long long* host4) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host4, gpmdm::g_small_stamps, sizeof(unsigned long long) * 4);
    return 0;
}
#endif
