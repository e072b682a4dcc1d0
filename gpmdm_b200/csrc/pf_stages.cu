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

namespace gpmdm {

constexpr int RB = 1024;  // elements per reduction block
constexpr int RT = 256;   // threads per reduction block

// Block-wide sum in a fixed order: lane tree inside each warp, then warp 0 adds the 8 warp totals
// serially.  All threads receive the result.
__device__ __forceinline__ double block_sum_fixed(double v, double* sh /*[RT/32 + 1]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < RT / 32; w++) t += sh[w];
        sh[RT / 32] = t;
    }
    __syncthreads();
    const double out = sh[RT / 32];
    __syncthreads();
    return out;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = sh[0];
        for (int w = 1; w < RT / 32; w++) t = fmax(t, sh[w]);
        sh[RT / 32] = t;
    }
    __syncthreads();
    const double out = sh[RT / 32];
    __syncthreads();
    return out;
}

// 4 consecutive doubles per thread as two 16-byte accesses (arrays come from the caller 16-byte aligned and `base`
// is a multiple of 4); the ragged tail of the array falls back to scalar accesses.
__device__ __forceinline__ void load4(const double* __restrict__ p, long long base, long long n, double (&v)[4],
                                      double fill) {
    if (base + 4 <= n) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(p + base));
        const double2 b = __ldg(reinterpret_cast<const double2*>(p + base + 2));
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = base + k < n ? p[base + k] : fill;
    }
}
__device__ __forceinline__ void store4(double* __restrict__ p, long long base, long long n, const double (&v)[4]) {
    if (base + 4 <= n) {
        *reinterpret_cast<double2*>(p + base) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(p + base + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (base + k < n) p[base + k] = v[k];
    }
}

// serial, fixed-order combination of per-block partials by one block (n_part <= a few thousand)
template <bool IS_MAX>
__global__ void __launch_bounds__(RT) combine_partials_kernel(const double* __restrict__ part, long long n_part,
                                                              double* __restrict__ out) {
    __shared__ double sh[RT / 32 + 1];
    // thread t owns the contiguous slice [t*per, (t+1)*per): order fixed by n_part alone
    const long long per = (n_part + RT - 1) / RT;
    const long long a = threadIdx.x * per, b = (a + per < n_part) ? a + per : n_part;
    double v = IS_MAX ? -INFINITY : 0.0;
    for (long long i = a; i < b; i++) v = IS_MAX ? fmax(v, part[i]) : v + part[i];
    const double r = IS_MAX ? block_max(v, sh) : block_sum_fixed(v, sh);
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
    const double* row = sT + c_prev[p] * C;
    const double* e = E + p * C;
    double best = -INFINITY;
    int arg = 0;
    int j = 0;
    if ((C & 1) == 0) {  // rows are 16-byte aligned
        for (; j + 8 <= C; j += 8) {
            double2 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = __ldcs(reinterpret_cast<const double2*>(e + j) + q);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double q0 = row[j + 2 * q] / v[q].x, q1 = row[j + 2 * q + 1] / v[q].y;  // IEEE division, as torch's `dist / q`
                if (q0 > best || (j + 2 * q == 0)) best = q0, arg = j + 2 * q;  // first maximum wins, as torch.argmax
                if (q1 > best) best = q1, arg = j + 2 * q + 1;
            }
        }
        for (; j + 2 <= C; j += 2) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(e + j));
            const double q0 = row[j] / v.x, q1 = row[j + 1] / v.y;
            if (q0 > best || j == 0) best = q0, arg = j;
            if (q1 > best) best = q1, arg = j + 1;
        }
    }
    for (; j < C; j++) {
        const double q = row[j] / __ldcs(e + j);
        if (q > best || j == 0) best = q, arg = j;
    }
    c_new[p] = arg;
}

// ---- bucket by class (stable counting sort) ---------------------------------------------------------------
// Each CTA handles BSUB consecutive 1024-particle sub-blocks, so that the single-CTA scan between the two passes has
// 4x fewer entries to walk.
constexpr int BSUB = 4;
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

__global__ void __launch_bounds__(1024) bucket_scan_kernel(int32_t* __restrict__ counts /*in: counts, out: offsets*/,
                                                           int C, int nb, int32_t* __restrict__ tiles,
                                                           int32_t* __restrict__ n_tiles) {
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
    if (threadIdx.x == 0) {
        cls_start[C] = carry;
        int t = 0;
        for (int c = 0; c < C; c++) {
            tile_start[c] = t;
            t += (cls_start[c + 1] - cls_start[c] + GPMDM_TILE_P - 1) / GPMDM_TILE_P;
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
            d[1] = cls_start[c] + t * GPMDM_TILE_P;
            d[2] = min(GPMDM_TILE_P, cnt - t * GPMDM_TILE_P);
            d[3] = 0;
        }
    }
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
        // exclusive scan over warps, one thread per class, continuing from the earlier sub-blocks
        for (int k = threadIdx.x; k < C; k += blockDim.x) {
            int run = base[k];
            for (int w = 0; w < 32; w++) {
                const int v = wc[w * C + k];
                wc[w * C + k] = run;
                run += v;
            }
            base[k] = run;
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
    const long long base = (long long)blockIdx.x * RB + threadIdx.x * 4;
    double va[4], vb[4] = {0.0, 0.0, 0.0, 0.0};
    load4(a, base, n, va, -INFINITY);
    if (b) load4(b, base, n, vb, 0.0);
    double v = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; k++) v = fmax(v, b ? va[k] + vb[k] : va[k]);
    v = block_max(v, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

__global__ void __launch_bounds__(RT) exp_sum_kernel(const double* __restrict__ ll, long long n,
                                                     const double* __restrict__ mx, double* __restrict__ lw,
                                                     double* __restrict__ w, double* __restrict__ part) {
    __shared__ double sh[RT / 32 + 1];
    const long long base = (long long)blockIdx.x * RB + threadIdx.x * 4;
    const double m = mx[0];
    double acc = 0.0;
    double vl[4], l[4], e[4];
    load4(ll, base, n, vl, -INFINITY);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        l[k] = vl[k] - m;
        e[k] = base + k < n ? exp(l[k]) : 0.0;
        acc += e[k];
    }
    store4(lw, base, n, l);
    store4(w, base, n, e);
    acc = block_sum_fixed(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void divide_kernel(double* __restrict__ w, long long n, const double* __restrict__ total) {
    const long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (base >= n) return;
    const double t = total[0];
    double v[4];
    load4(w, base, n, v, 0.0);
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = v[k] / t;  // IEEE division, as the reference
    store4(w, base, n, v);
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * RB + threadIdx.x * 4;
    double v[4];
    load4(w, base, n, v, 0.0);
    v[1] += v[0];
    v[2] += v[1];
    v[3] += v[2];
    double incl = v[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    double off = incl - v[3];  // exclusive prefix inside the warp
    for (int k = 0; k < warp; k++) off += wsum[k];
    double o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = off + v[k];
    store4(cdf, base, n, o);
    if (threadIdx.x == RT - 1) part[blockIdx.x] = off + v[3];
}
// Pass B: exclusive scan of the block totals by one block, 1024 entries per pass, in a fixed order (shuffle scan inside
// each warp, the warp totals scanned by warp 0, running carry) that depends only on the number of blocks.
__global__ void __launch_bounds__(1024) cdf_scan_partials_kernel(double* __restrict__ part, long long nb,
                                                                 double* __restrict__ total) {
    __shared__ double wtot[32];
    __shared__ double carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (long long base = 0; base < nb; base += 1024) {
        const long long i = base + threadIdx.x;
        const double v = i < nb ? part[i] : 0.0;
        double incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const double w = wtot[lane];
            double wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wtot[lane] = wi - w;
        }
        __syncthreads();
        const double excl = carry + (wtot[warp] + (incl - v));
        if (i < nb) part[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry;
}
// Pass C: add the block prefix, divide by the total, force the last entry to 1.
__global__ void cdf_finish_kernel(double* __restrict__ cdf, long long n, const double* __restrict__ part,
                                  const double* __restrict__ total) {
    const long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;  // 4 | RB: one block prefix
    if (base >= n) return;
    const double t = total[0], pre = part ? part[base / RB] : 0.0;
    double v[4];
    load4(cdf, base, n, v, 0.0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double x = part ? pre + v[k] : v[k];
        v[k] = (base + k == n - 1) ? 1.0 : x / t;
    }
    store4(cdf, base, n, v);
}

// ---- resampling search + gather -----------------------------------------------------------------------------
__global__ void resample_kernel(const double* __restrict__ cdf, long long P, const double* __restrict__ u,
                                long long n_out, const double* __restrict__ x_in, const int64_t* __restrict__ c_in,
                                int d, int64_t* __restrict__ anc, double* __restrict__ x_out,
                                int64_t* __restrict__ c_out) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_out) return;
    const double us = u[s];
    // first j with cdf[j] >= u -- the answer of ATen's binary search (MultinomialKernel.cpp) for a monotone cdf.
    // Two levels so that most probes hit cache: the last entry of every 1024-element block (a strided, L2-resident
    // subset), then inside one 8 KB block.
    const long long nblk = (P + RB - 1) / RB;
    long long lo = 0, hi = nblk;  // first block whose last entry is >= u
    while (hi - lo > 0) {
        const long long mid = lo + (hi - lo) / 2;
        const long long e = (mid + 1) * RB - 1;
        if (__ldg(cdf + (e < P ? e : P - 1)) < us) lo = mid + 1;
        else hi = mid;
    }
    long long left = lo * RB, right = (lo + 1) * RB < P ? (lo + 1) * RB : P;
    if (lo >= nblk) left = right = P;
    while (right - left > 0) {
        const long long mid = left + (right - left) / 2;
        if (__ldg(cdf + mid) < us) left = mid + 1;
        else right = mid;
    }
    if (left >= P) left = P - 1;
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
    const long long base = (long long)blockIdx.x * RB + threadIdx.x * 4;
    const double m = gmax[0];
    double e[4], ww[4], vll[4], vlw[4];
    int cc[4];
    load4(ll, base, n, vll, 0.0);
    load4(lw, base, n, vlw, 0.0);
    load4(w, base, n, ww, 0.0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool ok = base + k < n;
        e[k] = ok ? exp((vll[k] + vlw[k]) - m) : 0.0;
        cc[k] = ok ? (int)c_post[base + k] : -1;
    }
    // C + d + 1 block sums with two barriers in all: shuffle tree per value inside each warp, then one thread per
    // value adds the 8 warp totals serially -- the same order as block_sum_fixed, value by value.
    __shared__ double wtot[RT / 32][64 + GPMDM_MAX_LATENT + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncol = C + d + 1;
    for (int col = 0; col < ncol; col++) {
        double v = 0.0;
        if (col < C) {
#pragma unroll
            for (int k = 0; k < 4; k++) v += cc[k] == col ? e[k] : 0.0;
        } else if (col < C + d) {
            const int j = col - C;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (base + k < n) v += x_post[(base + k) * d + j] * ww[k];
        } else {
            v = (e[0] + e[1]) + (e[2] + e[3]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) wtot[warp][col] = v;
    }
    __syncthreads();
    if (threadIdx.x < ncol) {
        double t = 0.0;
        for (int w = 0; w < RT / 32; w++) t += wtot[w][threadIdx.x];
        part[(long long)blockIdx.x * ncol + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(RT) summaries_final_kernel(const double* __restrict__ part, long long nb, int C,
                                                             int d, double* __restrict__ out) {
    __shared__ double sh[RT / 32 + 1];
    __shared__ double cls[64 + GPMDM_MAX_LATENT + 1];
    const int ncol = C + d + 1;
    const long long per = (nb + RT - 1) / RT;
    const long long a = threadIdx.x * per, b = (a + per < nb) ? a + per : nb;
    for (int col = 0; col < ncol; col++) {
        double v = 0.0;
        for (long long i = a; i < b; i++) v += part[i * ncol + col];
        v = block_sum_fixed(v, sh);
        if (threadIdx.x == 0) cls[col] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int c = 0; c < C; c++) tot += cls[c];
        for (int c = 0; c < C; c++) out[c] = cls[c] / tot;  // class_likelihoods / sum (gpmdm_pf.py:246)
        for (int j = 0; j < d; j++) out[C + j] = cls[C + j];
        out[C + d] = cls[C + d];
    }
}

// ---- Philox4x32-10 draws ------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0;
        ctr[1] = n1;
        ctr[2] = n2;
        ctr[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// two uniforms in [0,1) with 53 random bits each
__device__ __forceinline__ void philox_u2(unsigned long long seed, unsigned long long step, unsigned long long p,
                                          uint32_t draw, double& a, double& b) {
    uint32_t ctr[4] = {(uint32_t)p, (uint32_t)(p >> 32), (uint32_t)step, draw};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
    const unsigned long long x = ((unsigned long long)ctr[1] << 32) | ctr[0];
    const unsigned long long y = ((unsigned long long)ctr[3] << 32) | ctr[2];
    a = (double)(x >> 11) * 0x1.0p-53;
    b = (double)(y >> 11) * 0x1.0p-53;
}

__global__ void philox_draws_kernel(unsigned long long seed, unsigned long long step, long long first, long long n,
                                    long long P_total, int C, int d, int systematic, double* __restrict__ E,
                                    double* __restrict__ eps, double* __restrict__ u) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long p = (unsigned long long)(first + i);
    uint32_t draw = 0;
    double a, b;
    if (E) {
        for (int j = 0; j < C; j += 2) {
            philox_u2(seed, step, p, draw++, a, b);
            E[i * C + j] = -log1p(-a);  // Exp(1), as torch's exponential_()
            if (j + 1 < C) E[i * C + j + 1] = -log1p(-b);
        }
    }
    draw = 0x1000;
    if (eps) {
        for (int j = 0; j < d; j += 2) {
            philox_u2(seed, step, p, draw++, a, b);
            const double rad = sqrt(-2.0 * log(1.0 - a));  // 1 - a in (0, 1]
            double sn, cs;
            sincospi(2.0 * b, &sn, &cs);
            eps[i * d + j] = rad * cs;
            if (j + 1 < d) eps[i * d + j + 1] = rad * sn;
        }
    }
    if (u) {
        if (systematic) {
            philox_u2(seed, step, 0xffffffffffffffffull, 0x2000, a, b);  // one shared offset u0
            u[i] = (a + (double)p) / (double)P_total;
        } else {
            philox_u2(seed, step, p, 0x2000, a, b);
            u[i] = a;
        }
    }
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

extern "C" int gpmdm_pf_bucket_by_class(const int64_t* classes, int64_t P, int32_t C, int32_t* perm, int32_t* tiles,
                                        int32_t* n_tiles, void* workspace, void* stream) {
    GPMDM_REQUIRE(P > 0 && P < (1ll << 31) && C >= 1 && C <= 1024, GPMDM_E_INVALID, "bad sizes P=%lld C=%d",
                  (long long)P, C);
    GPMDM_REQUIRE(classes && perm && tiles && n_tiles && workspace, GPMDM_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = nblocks(P, RB * BSUB);
    int32_t* counts = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + 256);
    bucket_count_kernel<<<nb, RB, C * sizeof(int), st>>>(classes, P, C, nb, counts);
    bucket_scan_kernel<<<1, 1024, 2 * (C + 1) * sizeof(int), st>>>(counts, C, nb, tiles, n_tiles);
    bucket_scatter_kernel<<<nb, RB, 33 * C * sizeof(int), st>>>(classes, P, C, nb, counts, perm);
    return check_launch("bucket_by_class");
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
