// Device code shared by the stage kernels (pf_stages.cu) and the fused small-cloud kernels (pf_small.cu): the fixed-order
// block reductions, vector accessors, the class-transition rule and the Philox generator.  Every floating-point
// reduction has ONE definition here, so the fused kernels reproduce the staged results bit for bit.
#pragma once
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

// NC = true: the array is read-only for the whole kernel (non-coherent path, ld.global.nc).  NC = false: it may have been
// written earlier in the SAME kernel by the SAME thread block -- the fused single-CTA kernels of pf_small.cu, the only
// users: ordinary generic loads, which the block's own earlier stores are visible to after a __syncthreads (one SM, one
// L1) and which also accept shared-memory pointers (the block partials of the small post kernel live there).
template <bool NC>
__device__ __forceinline__ double ld1(const double* p) {
    if (NC) return __ldg(p);
    return *p;
}
template <bool NC>
__device__ __forceinline__ double2 ld2(const double* p) {
    if (NC) return __ldg(reinterpret_cast<const double2*>(p));
    return *reinterpret_cast<const double2*>(p);
}
// 4 consecutive doubles per thread as two 16-byte accesses (arrays come from the caller 16-byte aligned and `base`
// is a multiple of 4); the ragged tail of the array falls back to scalar accesses.
template <bool NC = true>
__device__ __forceinline__ void load4(const double* p, long long base, long long n, double (&v)[4], double fill) {
    if (base + 4 <= n) {
        const double2 a = ld2<NC>(p + base);
        const double2 b = ld2<NC>(p + base + 2);
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = base + k < n ? ld1<NC>(p + base + k) : fill;
    }
}
__device__ __forceinline__ void store4(double* p, long long base, long long n, const double (&v)[4]) {
    if (base + 4 <= n) {
        *reinterpret_cast<double2*>(p + base) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(p + base + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (base + k < n) p[base + k] = v[k];
    }
}

// one particle: `row` = T[c_prev[p]] (shared memory), `e` = the particle's C Exp(1) draws
__device__ __forceinline__ int transition_one(const double* __restrict__ row, const double* __restrict__ e, int C) {
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
    return arg;
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


// ---- block-level bodies of the stage kernels: `vb` is the (virtual) 1024-element block, blockDim.x == RT ----------------
template <bool IS_MAX, bool NC>
__device__ __forceinline__ double combine_partials_dev(const double* part, long long n_part, double* sh) {
    // thread t owns the contiguous slice [t*per, (t+1)*per): order fixed by n_part alone
    const long long per = (n_part + RT - 1) / RT;
    const long long a = threadIdx.x * per, b = (a + per < n_part) ? a + per : n_part;
    double v = IS_MAX ? -INFINITY : 0.0;
    for (long long i = a; i < b; i++) v = IS_MAX ? fmax(v, ld1<NC>(part + i)) : v + ld1<NC>(part + i);
    return IS_MAX ? block_max(v, sh) : block_sum_fixed(v, sh);
}

template <bool NC>
__device__ __forceinline__ double block_max_dev(const double* a, const double* b, long long n, long long vb, double* sh) {
    const long long base = vb * RB + threadIdx.x * 4;
    double va[4], vbv[4] = {0.0, 0.0, 0.0, 0.0};
    load4<NC>(a, base, n, va, -INFINITY);
    if (b) load4<NC>(b, base, n, vbv, 0.0);
    double v = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; k++) v = fmax(v, b ? va[k] + vbv[k] : va[k]);
    return block_max(v, sh);
}

template <bool NC>
__device__ __forceinline__ double exp_sum_dev(const double* ll, long long n, double m, double* lw, double* w, long long vb,
                                              double* sh) {
    const long long base = vb * RB + threadIdx.x * 4;
    double acc = 0.0;
    double vl[4], l[4], e[4];
    load4<NC>(ll, base, n, vl, -INFINITY);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        l[k] = vl[k] - m;
        e[k] = base + k < n ? exp(l[k]) : 0.0;
        acc += e[k];
    }
    store4(lw, base, n, l);
    store4(w, base, n, e);
    return block_sum_fixed(acc, sh);
}

template <bool NC>
__device__ __forceinline__ void divide_dev(double* w, long long n, double t, long long base) {
    if (base >= n) return;
    double v[4];
    load4<NC>(w, base, n, v, 0.0);
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = v[k] / t;  // IEEE division, as the reference
    store4(w, base, n, v);
}

// inclusive scan inside one 1024-element block; returns the block total to thread RT - 1 (other threads: undefined)
template <bool NC>
__device__ __forceinline__ double cdf_block_scan_dev(const double* w, long long n, double* cdf, long long vb, double* wsum /*[RT/32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = vb * RB + threadIdx.x * 4;
    double v[4];
    load4<NC>(w, base, n, v, 0.0);
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
    __syncthreads();  // wsum is reused by the next block of a fused caller
    return off + v[3];
}

// exclusive scan of the block totals by one block of NT threads, NT entries per pass, in a fixed order (shuffle scan inside
// each warp, the warp totals scanned by warp 0, running carry) that depends only on the number of blocks -- and not on NT
// as long as nb <= NT (one pass): thread i sees the same warp / lane structure.  Returns the grand total to thread 0.
template <int NT, bool NC>
__device__ __forceinline__ double scan_partials_dev(double* part, long long nb, double* wtot /*[32]*/, double* carry_sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh[0] = 0.0;
    if (threadIdx.x < 32) wtot[threadIdx.x] = 0.0;
    __syncthreads();
    for (long long base = 0; base < nb; base += NT) {
        const long long i = base + threadIdx.x;
        const double v = i < nb ? ld1<NC>(part + i) : 0.0;
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
        const double excl = carry_sh[0] + (wtot[warp] + (incl - v));
        if (i < nb) part[i] = excl;
        __syncthreads();
        if (threadIdx.x == NT - 1) carry_sh[0] = excl + v;
        if (threadIdx.x < 32 && threadIdx.x >= NT / 32) wtot[threadIdx.x] = 0.0;
        __syncthreads();
    }
    return carry_sh[0];
}

// add the block prefix, divide by the total, force the last entry to 1 (4 elements from `base`)
template <bool NC>
__device__ __forceinline__ void cdf_finish_dev(double* cdf, long long n, const double* part, double t, long long base) {
    if (base >= n) return;
    const double pre = part ? ld1<NC>(part + base / RB) : 0.0;
    double v[4];
    load4<NC>(cdf, base, n, v, 0.0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double x = part ? pre + v[k] : v[k];
        v[k] = (base + k == n - 1) ? 1.0 : x / t;
    }
    store4(cdf, base, n, v);
}

// first j with cdf[j] >= u -- the answer of ATen's binary search (MultinomialKernel.cpp) for a monotone cdf.  Two levels
// so that most probes hit cache: the last entry of every 1024-element block (a strided, L2-resident subset), then inside
// one 8 KB block.
template <bool NC>
__device__ __forceinline__ long long cdf_search(const double* cdf, long long P, double us) {
    const long long nblk = (P + RB - 1) / RB;
    long long lo = 0, hi = nblk;  // first block whose last entry is >= u
    while (hi - lo > 0) {
        const long long mid = lo + (hi - lo) / 2;
        const long long e = (mid + 1) * RB - 1;
        if (ld1<NC>(cdf + (e < P ? e : P - 1)) < us) lo = mid + 1;
        else hi = mid;
    }
    long long left = lo * RB, right = (lo + 1) * RB < P ? (lo + 1) * RB : P;
    if (lo >= nblk) left = right = P;
    while (right - left > 0) {
        const long long mid = left + (right - left) / 2;
        if (ld1<NC>(cdf + mid) < us) left = mid + 1;
        else right = mid;
    }
    return left >= P ? P - 1 : left;
}

// per-block class sums of exp(g - max), state-mean partials, total: C + d + 1 block sums with two barriers in all --
// shuffle tree per value inside each warp, then one thread per value adds the 8 warp totals serially (the order of
// block_sum_fixed, value by value).  wtot [RT/32][64 + GPMDM_MAX_LATENT + 1] shared.
constexpr int SUMM_COLS = 64 + GPMDM_MAX_LATENT + 1;
template <bool NC>
__device__ __forceinline__ void summaries_block_dev(const double* ll, const double* lw, const double* w,
                                                    const int64_t* c_post, const double* x_post, long long n, int C, int d,
                                                    double m, double* part_row, long long vb, double (*wtot)[SUMM_COLS]) {
    const long long base = vb * RB + threadIdx.x * 4;
    double e[4], ww[4], vll[4], vlw[4];
    int cc[4];
    load4<NC>(ll, base, n, vll, 0.0);
    load4<NC>(lw, base, n, vlw, 0.0);
    load4<NC>(w, base, n, ww, 0.0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool ok = base + k < n;
        e[k] = ok ? exp((vll[k] + vlw[k]) - m) : 0.0;
        cc[k] = ok ? (int)(NC ? __ldg(c_post + base + k) : c_post[base + k]) : -1;
    }
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
                if (base + k < n) v += ld1<NC>(x_post + (base + k) * d + j) * ww[k];
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
        for (int wi = 0; wi < RT / 32; wi++) t += wtot[wi][threadIdx.x];
        part_row[threadIdx.x] = t;
    }
    __syncthreads();
}

template <bool NC>
__device__ __forceinline__ void summaries_final_dev(const double* part, long long nb, int C, int d, double* out, double* sh,
                                                    double* cls /*[SUMM_COLS]*/) {
    const int ncol = C + d + 1;
    const long long per = (nb + RT - 1) / RT;
    const long long a = threadIdx.x * per, b = (a + per < nb) ? a + per : nb;
    for (int col = 0; col < ncol; col++) {
        double v = 0.0;
        for (long long i = a; i < b; i++) v += ld1<NC>(part + i * ncol + col);
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

// one particle's raw draws (Philox keyed by the global particle index p; i = index into the local arrays)
__device__ __forceinline__ void philox_draw_one(unsigned long long seed, unsigned long long step, unsigned long long p,
                                                long long i, long long P_total, int C, int d, int systematic, double* E,
                                                double* eps, double* u) {
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

}  // namespace gpmdm
