// The HBM-bound stages of the filter step: class transition, class bucketing, weight normalisation,
// cdf, resampling search + gather, class/state summaries, and the Philox draw generator.
//
// Replaces (reference paths): gpmdm/gpmdm_pf.py:137-151 (_propogate_markov_switching), :161 (mask
// gather), :200-204 (normalisation), :206-213 (_resample), :215-262 (queries).
//
// Every floating-point reduction here runs in a FIXED order that depends only on the particle count
// (1024-element blocks, 256 threads x 4 consecutive elements, shuffle tree, then a serial pass over
// block partials), never on the grid size or the number of GPUs: a G-GPU run reproduces the 1-GPU
// run bit for bit.
#include <math.h>

#include "common.cuh"
#include "pf_stages.cuh"

namespace gpmdm {

// serial, fixed-order combination of per-block partials by one block (n_part <= a few thousand)
template <bool IS_MAX>
__global__ void __launch_bounds__(RT) combine_partials_kernel(const double* __restrict__ part, long long n_part,
                                                              double* __restrict__ out) {
    __shared__ double sh[RT / 32 + 1];
    const double r = combine_partials_dev<IS_MAX, true>(part, n_part, sh);
    if (threadIdx.x == 0) out[0] = r;
}

// ---- class transition ----------------------------------------------------------------------------------
// The Exp(1) draws are the bulk of the traffic (8 C bytes per particle): each thread streams its row with 16-byte
// loads issued ahead of the (sequential, first-maximum-wins) comparison chain; T lives in shared memory.
__global__ void __launch_bounds__(256) transition_kernel(const int64_t* __restrict__ c_prev,
                                                         const double* __restrict__ T, const double* __restrict__ E,
                                                         long long P, int C, int64_t* __restrict__ c_new) {
    extern __shared__ double sT[];  // [C][C]
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) sT[i] = T[i];
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    c_new[p] = transition_one(sT + c_prev[p] * C, E + p * C, C);
}

// ---- bucket by class (stable counting sort) ---------------------------------------------------------------
// Each CTA handles BSUB consecutive 1024-particle sub-blocks, so that the single-CTA scan between the two passes has
// BSUB x fewer entries to walk.
constexpr int BSUB = 8;
__global__ void __launch_bounds__(RB) bucket_count_kernel(const int64_t* __restrict__ cls, long long P, int C,
                                                          int nb, int32_t* __restrict__ counts /*[C][nb]*/) {
    extern __shared__ int sh_cnt[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < BSUB; sub++) {
        const long long p = ((long long)blockIdx.x * BSUB + sub) * RB + threadIdx.x;
        const int c = p < P ? (int)cls[p] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);  // one shared-memory atomic per class per warp
        if (c >= 0 && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&sh_cnt[c], __popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) counts[(long long)i * nb + blockIdx.x] = sh_cnt[i];
}

// writes the class-homogeneous tile table {block, first, count, 0} for tiles of `tile_p` particles
__device__ __forceinline__ void write_tile_table(const int* cls_start, int* tile_start, int C, int tile_p,
                                                 int32_t* __restrict__ tiles, int32_t* __restrict__ n_tiles) {
    if (threadIdx.x == 0) {
        int t = 0;
        for (int c = 0; c < C; c++) {
            tile_start[c] = t;
            t += (cls_start[c + 1] - cls_start[c] + tile_p - 1) / tile_p;
        }
        tile_start[C] = t;
        n_tiles[0] = t;
    }
    __syncthreads();
    for (int c = 0; c < C; c++) {
        const int nt = tile_start[c + 1] - tile_start[c];
        const int cnt = cls_start[c + 1] - cls_start[c];
        for (int t = threadIdx.x; t < nt; t += blockDim.x) {
            int32_t* d = tiles + 4ll * (tile_start[c] + t);
            d[0] = c;
            d[1] = cls_start[c] + t * tile_p;
            d[2] = min(tile_p, cnt - t * tile_p);
            d[3] = 0;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) bucket_scan_kernel(int32_t* __restrict__ counts /*in: counts, out: offsets*/,
                                                           int C, int nb, int32_t* __restrict__ tiles,
                                                           int32_t* __restrict__ n_tiles, int32_t* __restrict__ tiles2,
                                                           int32_t* __restrict__ n_tiles2, int tile_p2) {
    // exclusive scan over the class-major [C][nb] array, 1024 entries per pass: shuffle scan inside each warp, the 32
    // warp totals scanned by warp 0, running carry across passes
    __shared__ int wtot[32];
    __shared__ int carry;
    extern __shared__ int cls_start[];  // [C + 1] class start offsets, then [C + 1] tile starts
    int* tile_start = cls_start + (C + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const long long total = (long long)C * nb;
    for (long long base = 0; base < total; base += 1024) {
        const long long i = base + threadIdx.x;
        const int v = i < total ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wtot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wtot[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int excl = carry + wtot[warp] + incl - v;
        if (i < total) {
            counts[i] = excl;
            if (i % nb == 0) cls_start[i / nb] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cls_start[C] = carry;
    __syncthreads();
    write_tile_table(cls_start, tile_start, C, GPMDM_TILE_P, tiles, n_tiles);
    if (tiles2) write_tile_table(cls_start, tile_start, C, tile_p2, tiles2, n_tiles2);  // e.g. 128 for the tcgen05 kernels
}

__global__ void __launch_bounds__(RB) bucket_scatter_kernel(const int64_t* __restrict__ cls, long long P, int C,
                                                            int nb, const int32_t* __restrict__ offsets,
                                                            int32_t* __restrict__ perm) {
    extern __shared__ int wc[];  // [32 warps][C], then [C] running base of the CTA's earlier sub-blocks
    int* base = wc + 32 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < C; i += blockDim.x) base[i] = 0;
    for (int sub = 0; sub < BSUB; sub++) {
        for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) wc[i] = 0;
        __syncthreads();
        const long long p = ((long long)blockIdx.x * BSUB + sub) * RB + threadIdx.x;
        const int c = p < P ? (int)cls[p] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (c >= 0 && rank == 0) wc[warp * C + c] = __popc(peers);
        __syncthreads();
        // exclusive scan over the 32 warps' counts, one WARP per class (lane = warp index, five shuffle steps; a single
        // thread walking the 32 counts per class left 1000 threads idle for 32 dependent shared-memory updates),
        // continuing from the earlier sub-blocks
        for (int k = warp; k < C; k += RB / 32) {
            const int v = wc[lane * C + k];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int b = base[k];
            wc[lane * C + k] = b + incl - v;
            if (lane == 31) base[k] = b + incl;
        }
        __syncthreads();
        if (c >= 0) perm[offsets[(long long)c * nb + blockIdx.x] + wc[warp * C + c] + rank] = (int32_t)p;
        __syncthreads();
    }
}

// ---- normalisation -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RT) block_max_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       long long n, double* __restrict__ part) {
    __shared__ double sh[RT / 32 + 1];
    const double v = block_max_dev<true>(a, b, n, blockIdx.x, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

__global__ void __launch_bounds__(RT) exp_sum_kernel(const double* __restrict__ ll, long long n,
                                                     const double* __restrict__ mx, double* __restrict__ lw,
                                                     double* __restrict__ w, double* __restrict__ part) {
    __shared__ double sh[RT / 32 + 1];
    const double acc = exp_sum_dev<true>(ll, n, mx[0], lw, w, blockIdx.x, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void divide_kernel(double* __restrict__ w, long long n, const double* __restrict__ total) {
    divide_dev<true>(w, n, total[0], ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4);
}

// ---- cdf ---------------------------------------------------------------------------------------------------
// mode 0: the reference's order (ATen CPU cumsum: one running sum in index order).
__global__ void cdf_sequential_kernel(const double* __restrict__ w, long long n, double* __restrict__ cdf,
                                      double* __restrict__ total) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double run = 0.0;
    long long i = 0;
    for (; i + 8 <= n; i += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = w[i + k];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            run = __dadd_rn(run, v[k]);
            cdf[i + k] = run;
        }
    }
    for (; i < n; i++) {
        run = __dadd_rn(run, w[i]);
        cdf[i] = run;
    }
    total[0] = run;
}

// mode 1: blocked scan.  Pass A: inclusive scan inside each 1024-element block + block total.
__global__ void __launch_bounds__(RT) cdf_block_scan_kernel(const double* __restrict__ w, long long n,
                                                            double* __restrict__ cdf, double* __restrict__ part) {
    __shared__ double wsum[RT / 32];
    const double tot = cdf_block_scan_dev<true>(w, n, cdf, blockIdx.x, wsum);
    if (threadIdx.x == RT - 1) part[blockIdx.x] = tot;
}
// Pass B: exclusive scan of the block totals by one block (scan_partials_dev).
__global__ void __launch_bounds__(1024) cdf_scan_partials_kernel(double* __restrict__ part, long long nb,
                                                                 double* __restrict__ total) {
    __shared__ double wtot[32];
    __shared__ double carry;
    const double t = scan_partials_dev<1024, true>(part, nb, wtot, &carry);
    if (threadIdx.x == 0) total[0] = t;
}
// Pass C: add the block prefix, divide by the total, force the last entry to 1.
__global__ void cdf_finish_kernel(double* __restrict__ cdf, long long n, const double* __restrict__ part,
                                  const double* __restrict__ total) {
    cdf_finish_dev<true>(cdf, n, part, total[0], ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4);  // 4 | RB: one block prefix
}

// ---- resampling search + gather -----------------------------------------------------------------------------
__global__ void resample_kernel(const double* __restrict__ cdf, long long P, const double* __restrict__ u,
                                long long n_out, const double* __restrict__ x_in, const int64_t* __restrict__ c_in,
                                int d, int64_t* __restrict__ anc, double* __restrict__ x_out,
                                int64_t* __restrict__ c_out) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_out) return;
    const long long left = cdf_search<true>(cdf, P, u[s]);
    if (anc) anc[s] = left;
    if (x_out)
        for (int k = 0; k < d; k++) x_out[s * d + k] = x_in[left * d + k];
    if (c_out) c_out[s] = c_in[left];
}

// Ascending u (the systematic comb): the 1024 outputs of a block fall into one contiguous window [j_lo, j_hi] of the
// cdf, found by two global searches per block; the window is staged in shared memory with coalesced loads and every
// output is a search in shared memory.  The answer is the same index as resample_kernel's (the cdf is monotone, so the
// first j with cdf_j >= u lies inside the window).  Windows wider than the staging buffer (many consecutive particles of
// ~zero weight) fall back to a global search inside the window.
constexpr int RS_OUT = 1024;   // outputs per block
constexpr int RS_CAP = 4096;   // cdf entries staged per block (32 KB)
// first j in [0, P] with cdf[j] >= us (P if none), by a whole warp: 32 probes per step, log32(P) dependent steps
__device__ __forceinline__ long long warp_lower_bound(const double* __restrict__ cdf, long long P, double us) {
    const int lane = threadIdx.x & 31;
    long long lo = 0, hi = P;  // the answer lies in [lo, hi]; entries below lo are < us
    while (hi - lo > 0) {
        const long long len = hi - lo, step = (len + 31) / 32;
        long long probe = lo + (lane + 1) * step - 1;  // last entry of this lane's chunk
        if (probe > hi - 1) probe = hi - 1;
        const unsigned ge = __ballot_sync(0xffffffffu, __ldg(cdf + probe) >= us);
        if (ge == 0) return hi;  // every entry of [lo, hi) is < us
        const int first = __ffs(ge) - 1;
        const long long nlo = lo + first * step;
        long long nhi = lo + (first + 1) * step - 1;  // cdf[nhi] >= us: the answer is in [nlo, nhi]
        if (nhi > hi - 1) nhi = hi - 1;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

__global__ void __launch_bounds__(256) resample_sorted_kernel(const double* __restrict__ cdf, long long P,
                                                              const double* __restrict__ u, long long n_out,
                                                              const double* __restrict__ x_in,
                                                              const int64_t* __restrict__ c_in, int d,
                                                              int64_t* __restrict__ anc, double* __restrict__ x_out,
                                                              int64_t* __restrict__ c_out) {
    __shared__ double win[RS_CAP];
    __shared__ long long bounds[2];
    const long long s0 = (long long)blockIdx.x * RS_OUT;
    const long long s1 = s0 + RS_OUT < n_out ? s0 + RS_OUT : n_out;  // exclusive
    if (threadIdx.x < 64) {  // warp 0: window start, warp 1: window end
        const long long b = warp_lower_bound(cdf, P, u[threadIdx.x < 32 ? s0 : s1 - 1]);
        if ((threadIdx.x & 31) == 0) bounds[threadIdx.x >> 5] = b;
    }
    __syncthreads();
    const long long j_lo = bounds[0], j_hi = bounds[1] < P ? bounds[1] : P - 1;  // answers lie in [j_lo, j_hi]
    const long long span = j_hi - j_lo + 1;
    const bool staged = span <= RS_CAP;
    if (staged)
        for (long long i = threadIdx.x; i < span; i += blockDim.x) win[i] = cdf[j_lo + i];
    __syncthreads();
    for (long long s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
        const double us = u[s];
        long long left = 0, right = span;  // first index in the window with cdf >= u (== span only if u > cdf[j_hi])
        if (staged) {
            while (right - left > 0) {
                const long long mid = left + (right - left) / 2;
                if (win[mid] < us) left = mid + 1;
                else right = mid;
            }
        } else {
            while (right - left > 0) {
                const long long mid = left + (right - left) / 2;
                if (__ldg(cdf + j_lo + mid) < us) left = mid + 1;
                else right = mid;
            }
        }
        long long a = j_lo + left;
        if (a >= P) a = P - 1;
        if (anc) anc[s] = a;
        if (x_out)
            for (int k = 0; k < d; k++) x_out[s * d + k] = x_in[a * d + k];
        if (c_out) c_out[s] = c_in[a];
    }
}

// ---- summaries --------------------------------------------------------------------------------------------
// part [nb][C + d + 1]: per-block class sums of exp(g - max), state-mean partials, total.
__global__ void __launch_bounds__(RT) summaries_block_kernel(const double* __restrict__ ll,
                                                             const double* __restrict__ lw,
                                                             const double* __restrict__ w,
                                                             const int64_t* __restrict__ c_post,
                                                             const double* __restrict__ x_post, long long n, int C,
                                                             int d, const double* __restrict__ gmax,
                                                             double* __restrict__ part) {
    __shared__ double wtot[RT / 32][SUMM_COLS];
    summaries_block_dev<true>(ll, lw, w, c_post, x_post, n, C, d, gmax[0], part + (long long)blockIdx.x * (C + d + 1),
                              blockIdx.x, wtot);
}

__global__ void __launch_bounds__(RT) summaries_final_kernel(const double* __restrict__ part, long long nb, int C,
                                                             int d, double* __restrict__ out) {
    __shared__ double sh[RT / 32 + 1];
    __shared__ double cls[SUMM_COLS];
    summaries_final_dev<true>(part, nb, C, d, out, sh, cls);
}

// ---- Philox4x32-10 draws (generator in pf_stages.cuh) --------------------------------------------------------
__global__ void philox_draws_kernel(unsigned long long seed, unsigned long long step, long long first, long long n,
                                    long long P_total, int C, int d, int systematic, double* __restrict__ E,
                                    double* __restrict__ eps, double* __restrict__ u) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    philox_draw_one(seed, step, (unsigned long long)(first + i), i, P_total, C, d, systematic, E, eps, u);
}

static inline int nblocks(long long n, int per) { return (int)((n + per - 1) / per); }

}  // namespace gpmdm

using namespace gpmdm;

// workspace layout (bytes): [0, 64) scalars {max, sum, ...}; then partials
extern "C" int64_t gpmdm_workspace_bytes(int64_t P, int32_t C) {
    const int64_t nb = (P + RB - 1) / RB;
    const int64_t part = nb * (int64_t)(C + GPMDM_MAX_LATENT + 1) * 8;  // summaries partials (largest user)
    const int64_t counts = (int64_t)C * nb * 4;
    return 256 + round_up(part > counts ? part : counts, 256) + round_up(nb * 8, 256);
}

extern "C" int gpmdm_pf_transition_f64(const int64_t* c_prev, const double* T, const double* E, int64_t P, int32_t C,
                                       int64_t* c_new, void* stream) {
    GPMDM_REQUIRE(P >= 0 && C >= 1 && C <= 64, GPMDM_E_INVALID, "bad sizes P=%lld C=%d (1 <= C <= 64)", (long long)P, C);
    if (P == 0) return 0;
    GPMDM_REQUIRE(c_prev && T && E && c_new, GPMDM_E_INVALID, "null argument");
    transition_kernel<<<nblocks(P, 256), 256, (size_t)C * C * sizeof(double), (cudaStream_t)stream>>>(c_prev, T, E, P, C, c_new);
    return check_launch("transition_kernel");
}

extern "C" int gpmdm_pf_bucket_by_class2(const int64_t* classes, int64_t P, int32_t C, int32_t* perm, int32_t* tiles,
                                         int32_t* n_tiles, int32_t* tiles128, int32_t* n_tiles128, void* workspace,
                                         void* stream) {
    GPMDM_REQUIRE(P > 0 && P < (1ll << 31) && C >= 1 && C <= 1024, GPMDM_E_INVALID, "bad sizes P=%lld C=%d",
                  (long long)P, C);
    GPMDM_REQUIRE(classes && perm && tiles && n_tiles && workspace, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE((tiles128 == nullptr) == (n_tiles128 == nullptr), GPMDM_E_INVALID, "tiles128 and n_tiles128 go together");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = nblocks(P, RB * BSUB);
    int32_t* counts = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + 256);
    bucket_count_kernel<<<nb, RB, C * sizeof(int), st>>>(classes, P, C, nb, counts);
    bucket_scan_kernel<<<1, 1024, 2 * (C + 1) * sizeof(int), st>>>(counts, C, nb, tiles, n_tiles, tiles128, n_tiles128, 128);
    bucket_scatter_kernel<<<nb, RB, 33 * C * sizeof(int), st>>>(classes, P, C, nb, counts, perm);
    return check_launch("bucket_by_class");
}

extern "C" int gpmdm_pf_bucket_by_class(const int64_t* classes, int64_t P, int32_t C, int32_t* perm, int32_t* tiles,
                                        int32_t* n_tiles, void* workspace, void* stream) {
    return gpmdm_pf_bucket_by_class2(classes, P, C, perm, tiles, n_tiles, nullptr, nullptr, workspace, stream);
}

extern "C" int gpmdm_pf_normalize_f64(const double* ll, int64_t P, double* lw, double* w, double* stats_out,
                                      void* workspace, void* stream) {
    GPMDM_REQUIRE(P > 0, GPMDM_E_INVALID, "P must be positive");
    GPMDM_REQUIRE(ll && lw && w && workspace, GPMDM_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = nblocks(P, RB);
    double* scal = static_cast<double*>(workspace);  // [0]=max [1]=sum
    double* part = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    block_max_kernel<<<nb, RT, 0, st>>>(ll, nullptr, P, part);
    combine_partials_kernel<true><<<1, RT, 0, st>>>(part, nb, scal);
    exp_sum_kernel<<<nb, RT, 0, st>>>(ll, P, scal, lw, w, part);
    combine_partials_kernel<false><<<1, RT, 0, st>>>(part, nb, scal + 1);
    divide_kernel<<<nblocks(P, 1024), 256, 0, st>>>(w, P, scal + 1);
    if (stats_out) {
        cudaError_t e = cudaMemcpyAsync(stats_out, scal, 16, cudaMemcpyDeviceToDevice, st);
        GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
    }
    return check_launch("normalize");
}

extern "C" int gpmdm_pf_cdf_f64(const double* w, int64_t P, int32_t mode, double* cdf, void* workspace,
                                void* stream) {
    GPMDM_REQUIRE(P > 0 && (mode == 0 || mode == 1), GPMDM_E_INVALID, "bad arguments P=%lld mode=%d", (long long)P,
                  mode);
    GPMDM_REQUIRE(w && cdf && workspace, GPMDM_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* scal = static_cast<double*>(workspace) + 2;  // total
    double* part = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    if (mode == 0) {
        cdf_sequential_kernel<<<1, 32, 0, st>>>(w, P, cdf, scal);
        cdf_finish_kernel<<<nblocks(P, 1024), 256, 0, st>>>(cdf, P, nullptr, scal);
    } else {
        const int nb = nblocks(P, RB);
        cdf_block_scan_kernel<<<nb, RT, 0, st>>>(w, P, cdf, part);
        cdf_scan_partials_kernel<<<1, 1024, 0, st>>>(part, nb, scal);
        cdf_finish_kernel<<<nblocks(P, 1024), 256, 0, st>>>(cdf, P, part, scal);
    }
    return check_launch("cdf");
}

extern "C" int gpmdm_pf_resample_f64(const double* cdf, int64_t P, const double* u, int64_t n_out,
                                     const double* x_in, const int64_t* c_in, int32_t d, int64_t* anc, double* x_out,
                                     int64_t* c_out, void* stream) {
    GPMDM_REQUIRE(P > 0 && n_out >= 0 && d >= 0, GPMDM_E_INVALID, "bad sizes");
    if (n_out == 0) return 0;
    GPMDM_REQUIRE(cdf && u, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE((x_out == nullptr || x_in != nullptr) && (c_out == nullptr || c_in != nullptr), GPMDM_E_INVALID,
                  "gather output without input");
    resample_kernel<<<nblocks(n_out, 256), 256, 0, (cudaStream_t)stream>>>(cdf, P, u, n_out, x_in, c_in, d, anc,
                                                                           x_out, c_out);
    return check_launch("resample_kernel");
}

extern "C" int gpmdm_pf_resample_sorted_f64(const double* cdf, int64_t P, const double* u, int64_t n_out,
                                            const double* x_in, const int64_t* c_in, int32_t d, int64_t* anc,
                                            double* x_out, int64_t* c_out, void* stream) {
    GPMDM_REQUIRE(P > 0 && n_out >= 0 && d >= 0, GPMDM_E_INVALID, "bad sizes");
    if (n_out == 0) return 0;
    GPMDM_REQUIRE(cdf && u, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE((x_out == nullptr || x_in != nullptr) && (c_out == nullptr || c_in != nullptr), GPMDM_E_INVALID,
                  "gather output without input");
    resample_sorted_kernel<<<nblocks(n_out, RS_OUT), 256, 0, (cudaStream_t)stream>>>(cdf, P, u, n_out, x_in, c_in, d, anc,
                                                                                    x_out, c_out);
    return check_launch("resample_sorted_kernel");
}

extern "C" int gpmdm_pf_summaries_f64(const double* ll, const double* lw, const double* w, const int64_t* c_post,
                                      const double* x_post, int64_t P, int32_t C, int32_t d, double* out,
                                      void* workspace, void* stream) {
    GPMDM_REQUIRE(P > 0 && C >= 1 && C <= 64 && d >= 1 && d <= GPMDM_MAX_LATENT, GPMDM_E_INVALID,
                  "bad sizes P=%lld C=%d d=%d", (long long)P, C, d);
    GPMDM_REQUIRE(ll && lw && w && c_post && x_post && out && workspace, GPMDM_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = nblocks(P, RB);
    double* scal = static_cast<double*>(workspace) + 3;  // max of ll + lw
    double* part = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    block_max_kernel<<<nb, RT, 0, st>>>(ll, lw, P, part);
    combine_partials_kernel<true><<<1, RT, 0, st>>>(part, nb, scal);
    summaries_block_kernel<<<nb, RT, 0, st>>>(ll, lw, w, c_post, x_post, P, C, d, scal, part);
    summaries_final_kernel<<<1, RT, 0, st>>>(part, nb, C, d, out);
    return check_launch("summaries");
}

extern "C" int gpmdm_pf_draws_philox(uint64_t seed, uint64_t step, int64_t first, int64_t n, int64_t P_total,
                                     int32_t C, int32_t d, int32_t systematic, double* E, double* eps, double* u,
                                     void* stream) {
    GPMDM_REQUIRE(n >= 0 && first >= 0 && P_total >= first + n, GPMDM_E_INVALID, "bad particle range");
    if (n == 0) return 0;
    philox_draws_kernel<<<nblocks(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, step, first, n, P_total, C, d,
                                                                          systematic, E, eps, u);
    return check_launch("philox_draws_kernel");
}
