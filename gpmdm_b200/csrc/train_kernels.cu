// Training-side kernel-matrix construction and the NLL gradient terms through it.
//
// Replaces (reference paths): gpmdm/gpmdm.py:381-548 (get_y_kernel / get_x_kernel / get_rbf_kernel /
// get_weighted_distances / get_lin_kernel), the dense 0/1 class masks of :311-378 (`* self.M`,
// :616, :1292) and the autograd backward through those ops.  The O(N^3) factorisation stays with
// torch.linalg (cuSOLVER); these kernels are HBM-bound: the build writes 8 N^2 bytes, the gradient
// reads G once (8 N^2 bytes: each pair of mirror-image tiles is visited by one block, row-wise and as coalesced
// 256-byte column segments).
#include <math.h>

#include "common.cuh"
#include "fast_exp.cuh"

namespace gpmdm {

constexpr int MAXD_T = GPMDM_MAX_LATENT;

__device__ __forceinline__ int class_of_row(long long i, const int64_t* __restrict__ offs, int n_classes) {
    int c = 0;
    while (c + 1 < n_classes && i >= offs[c + 1]) c++;
    return c;
}

__constant__ double c_exp_table_train[64] = {GPMDM_EXP_TABLE_VALUES};

// K is symmetric: one block computes a 64 x 64 tile (I <= J) once and writes it twice -- rows of tile (I, J) directly and
// tile (J, I) through a shared-memory transpose -- so every exponential is evaluated once and both writes are coalesced.
// Off-class tiles are zero-filled without any math.  256 threads, 16 elements each.  The latent dimension and the kernel
// kind are compile-time (ncu on the run-time-d version: 85 % issue-slot utilisation, i.e. instruction bound, not HBM
// bound): the thread's column record stays in registers and the d-loops unroll.
constexpr int BT = 64;
template <int DL, int KIND>
__global__ void __launch_bounds__(256) kernel_build_kernel(const double* __restrict__ X, long long n,
                                                           const double* __restrict__ ls,
                                                           const double* __restrict__ lin_c2, double noise2,
                                                           const int64_t* __restrict__ offs, int n_classes,
                                                           double* __restrict__ K) {
    constexpr int d = DL;
    constexpr int AW = (DL + 2) & ~1;  // a_i[0..d), |a_i|^2, padded to an even width (16-byte shared loads)
    const int I = blockIdx.y, J = blockIdx.x;
    if (J < I) return;
    __shared__ __align__(16) double ai[BT][AW], aj[BT][AW];  // x / l and |x/l|^2
    __shared__ __align__(16) double xi[BT][KIND == 1 ? ((DL + 1) & ~1) : 2];  // c_k^2 x_ik
    __shared__ int ci[BT], cj[BT];
    __shared__ double tile[BT][BT + 1];
    __shared__ double exptab[64];
    const long long i0 = (long long)I * BT, j0 = (long long)J * BT;
    const int t = threadIdx.x;
    if (t < 64) exptab[t] = c_exp_table_train[t];
    if (t < 2 * BT) {
        const bool is_i = t < BT;
        const int r = t & (BT - 1);
        const long long row = (is_i ? i0 : j0) + r;
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < d; k++) {
            const double x = row < n ? X[row * d + k] : 0.0;
            const double a = x / ls[k];
            (is_i ? ai : aj)[r][k] = a;
            if (KIND == 1 && is_i) xi[r][k] = lin_c2[k] * x;
            n2 = fma(a, a, n2);
        }
        (is_i ? ai : aj)[r][d] = n2;
        (is_i ? ci : cj)[r] = (offs && row < n) ? class_of_row(row, offs, n_classes) : 0;
    }
    __syncthreads();
    const int tx = t & 63, ty = t >> 6;  // column within the tile, 4 row groups
    const double c2last = KIND == 1 ? lin_c2[d] : 0.0;
    const long long j = j0 + tx;
    double ajr[DL + 1], xraw[KIND == 1 ? DL : 1];
#pragma unroll
    for (int k = 0; k <= d; k++) ajr[k] = aj[tx][k];
    if (KIND == 1) {
#pragma unroll
        for (int k = 0; k < d; k++) xraw[k] = j < n ? X[j * d + k] : 0.0;
    }
    const int cjx = cj[tx];
    double* const kcol = K + j;
#pragma unroll 4
    for (int r = ty; r < BT; r += 4) {
        const long long i = i0 + r;
        double v = 0.0;
        if (i < n && j < n && (!offs || ci[r] == cjx)) {
            double air[AW];
#pragma unroll
            for (int q = 0; q < AW; q += 2) {
                const double2 w = *reinterpret_cast<const double2*>(&ai[r][q]);
                air[q] = w.x, air[q + 1] = w.y;
            }
            double dot = 0.0;
#pragma unroll
            for (int k = 0; k < d; k++) dot = fma(air[k], ajr[k], dot);
            v = fast_exp(-(air[d] + ajr[d] - 2.0 * dot), exptab);  // gpmdm.py:515-517 expansion form
            if (i == j) v += noise2;
            if (KIND == 1) {
                double lin = c2last;
#pragma unroll
                for (int k = 0; k < d; k++) lin = fma(xi[r][k], xraw[k], lin);
                v += lin;
            }
        }
        tile[r][tx] = v;
        if (i < n && j < n) kcol[i * n] = v;
    }
    if (I == J) return;
    __syncthreads();
    const long long ii = i0 + tx;
#pragma unroll 4
    for (int r = ty; r < BT; r += 4) {  // transposed copy: row (j0 + r), columns i0 + tx
        const long long jj = j0 + r;
        if (jj < n && ii < n) K[jj * n + ii] = tile[tx][r];
    }
}

// Gradient terms.  With S = G^ + G^^T (G^ = G o M) every UNORDERED pair (i, j) contributes to row i and to row j:
//     dL/dx_i += S_ij [ k_rbf (-2 D_ij / l^2) + c^2 x_j ]        dL/dx_j += S_ij [ k_rbf (+2 D_ij / l^2) + c^2 x_i ]
//     dL/dlog l_k += S_ij k_rbf 2 D_ijk^2 / l_k^2                  dL/dlog c_k += 2 c_k^2 S_ij x~_ik x~_jk   (App. A.5)
// so block (b, s) owns the 32-row strip b and walks only the column tiles J >= b of split s of its (class-restricted)
// column range: every tile PAIR {(I, J), (J, I)} of G is read exactly once -- 8 N^2 bytes in all, where the row-wise
// kernel this replaces read 16 N^2 -- and every exponential is evaluated once instead of twice.
//   row side  -> registers over the walk -> partX [NS][n][d]      (per-split row sums)
//   col side  -> per tile, reduced over the 8 warps in shared memory -> partC[I][j][d], j >= 32 I (each entry is written
//                by exactly one block); the second kernel adds the strips I <= j/32 in a fixed order
//   part [nblk * NS][2 d + 2]: per-block partials of g_log_ls[d], tr(G^), g_log_c[d+1].
// offset (in rows of d doubles) of strip I in partC: rows j = 32 I .. 32 nblk - 1
__host__ __device__ __forceinline__ long long partc_offset(long long I, long long nblk) {
    return 32 * (I * nblk - I * (I - 1) / 2);
}

template <int DL>
__global__ void __launch_bounds__(256) kernel_grad_kernel(const double* __restrict__ X, const double* __restrict__ G,
                                                          long long n, int kind, const double* __restrict__ ls,
                                                          const double* __restrict__ lin_c2,
                                                          const int64_t* __restrict__ offs, int n_classes,
                                                          double* __restrict__ partX, double* __restrict__ partC,
                                                          double* __restrict__ part) {
    __shared__ double ai[32][MAXD_T], xi[32][MAXD_T];  // row side: read as broadcasts
    __shared__ double aj[MAXD_T][32], xj[MAXD_T][32];  // column side: k-major, lane = column (conflict-free)
    __shared__ double GT[32][33];
    __shared__ double colred[8][DL][32];               // column-side partial sums of the 8 warps
    __shared__ int ci[32], cj[32];
    __shared__ double red[8][2 * MAXD_T + 2];
    __shared__ double exptab[64];
    __shared__ double w2[MAXD_T];                      // 2 / l_k
    constexpr int d = DL;  // compile-time latent dimension: the per-thread accumulators below stay in registers
    const long long I = blockIdx.x, i0 = I * 32, nblk = gridDim.x;
    const int NS = gridDim.y, split = blockIdx.y;
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 32 + tx;
    if (t < 64) exptab[t] = c_exp_table_train[t];
    if (t < d) w2[t] = 2.0 / ls[t];
    if (t < 32) {
        const long long row = i0 + t;
        for (int k = 0; k < d; k++) {
            const double x = row < n ? X[row * d + k] : 0.0;
            xi[t][k] = x;
            ai[t][k] = x * (1.0 / ls[k]);
        }
        ci[t] = (offs && row < n) ? class_of_row(row, offs, n_classes) : 0;
    }
    double gx[4][DL], gl[DL], gc[DL + 1], tr = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < d; k++) gx[q][k] = 0.0;
#pragma unroll
    for (int k = 0; k < d; k++) gl[k] = gc[k] = 0.0;
    gc[d] = 0.0;

    // the walk: column tiles from the strip's own diagonal tile to the end of its class range (class-masked gradients
    // vanish outside the class block; rows of one strip may straddle classes: walk up to the last one's end)
    long long jbeg = i0, jend = n;
    __syncthreads();
    if (offs) {
        const long long last = (i0 + 31 < n ? i0 + 31 : n - 1);
        jend = offs[class_of_row(last, offs, n_classes) + 1];
    }
    {
        const long long ntile = (jend - jbeg + 31) / 32, per = (ntile + NS - 1) / NS;
        const long long a0 = jbeg + split * per * 32;
        const long long a1 = a0 + per * 32;
        jbeg = a0;
        jend = a1 < jend ? a1 : jend;
    }
    double* pc = partC + partc_offset(I, nblk) * d - i0 * d;  // pc[j * d + k], j >= i0
    // G values of the current tile live in registers; the next tile's are prefetched while this one is computed, so
    // the HBM latency of the two 8 KB tile reads (G[i][j] row-wise, G[j][i] as 256-byte column segments) is hidden.
    double gcur[4], gtc[4];
    auto fetch = [&](long long j0, double (&gv)[4], double (&gt)[4]) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = ty + 8 * q;
            const long long i = i0 + r, j = j0 + tx;
            gv[q] = (j0 < jend && i < n && j < n) ? G[i * n + j] : 0.0;
            const long long jj = j0 + r, ii = i0 + tx;
            gt[q] = (j0 < jend && jj < n && ii < n) ? G[jj * n + ii] : 0.0;
        }
    };
    fetch(jbeg, gcur, gtc);
    for (long long j0 = jbeg; j0 < jend; j0 += 32) {
        __syncthreads();
        if (t < 32) {
            const long long row = j0 + t;
            for (int k = 0; k < d; k++) {
                const double x = row < n ? X[row * d + k] : 0.0;
                xj[k][t] = x;
                aj[k][t] = x * (0.5 * w2[k]);
            }
            cj[t] = (offs && row < n) ? class_of_row(row, offs, n_classes) : 0;
        }
        // transposed tile: GT[r][c] = G[j0 + r][i0 + c]
#pragma unroll
        for (int q = 0; q < 4; q++) GT[ty + 8 * q][tx] = gtc[q];
        __syncthreads();
        double gnext[4], gtn[4];
        fetch(j0 + 32, gnext, gtn);
        const long long j = j0 + tx;
        const bool diag_tile = j0 == i0;
        double ajr[DL], xjr[DL], gcol[DL];  // this lane's column, shared by its four rows
#pragma unroll
        for (int k = 0; k < d; k++) {
            ajr[k] = aj[k][tx];
            xjr[k] = kind == 1 ? xj[k][tx] : 0.0;
            gcol[k] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = ty + 8 * q;
            const long long i = i0 + r;
            if (i >= n || j >= n) continue;
            if (offs && ci[r] != cj[tx]) continue;
            if (diag_tile && i >= j) {
                if (i == j) {  // diagonal element: S_ii = 2 G_ii, k_rbf = 1, D = 0
                    const double g = gcur[q];
                    tr += g;
                    if (kind == 1) {
#pragma unroll
                        for (int k = 0; k < d; k++) {
                            gx[q][k] = fma(2.0 * g * lin_c2[k], xjr[k], gx[q][k]);
                            gc[k] = fma(g, xjr[k] * xjr[k], gc[k]);
                        }
                        gc[d] += g;
                    }
                }
                continue;  // pairs below the diagonal are handled from their mirror image
            }
            const double s = gcur[q] + GT[tx][r];  // S_ij = G^_ij + G^_ji
            double dist = 0.0;
            double dk[DL];
#pragma unroll
            for (int k = 0; k < d; k++) {
                dk[k] = ai[r][k] - ajr[k];
                dist = fma(dk[k], dk[k], dist);
            }
            const double skr = s * fast_exp(-dist, exptab);
#pragma unroll
            for (int k = 0; k < d; k++) {
                // d k_rbf / d x_ik = k_rbf * (-2 (x_ik - x_jk) / l_k^2) = -d k_rbf / d x_jk
                const double u = skr * dk[k] * w2[k];
                gx[q][k] -= u;
                gcol[k] += u;
                gl[k] = fma(skr * dk[k], 2.0 * dk[k], gl[k]);
                if (kind == 1) {
                    const double sc = s * lin_c2[k];
                    gx[q][k] = fma(sc, xjr[k], gx[q][k]);
                    gcol[k] = fma(sc, xi[r][k], gcol[k]);
                    gc[k] = fma(s, xi[r][k] * xjr[k], gc[k]);
                }
            }
            if (kind == 1) gc[d] += s;
        }
        // column sums of this tile: the 8 warps' partials added in warp order by one thread per (column, k)
#pragma unroll
        for (int k = 0; k < d; k++) colred[ty][k][tx] = gcol[k];
        __syncthreads();
        if (t < 32 * d) {
            const int c = t & 31, k = t >> 5;
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) v += colred[w][k][c];
            if (j0 + c < n) pc[(j0 + c) * d + k] = v;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            gcur[q] = gnext[q];
            gtc[q] = gtn[q];
        }
    }
    // row sums: lanes of a warp share ty, i.e. the same four rows
    double* px = partX + (long long)split * n * d;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const long long i = i0 + ty + 8 * q;
#pragma unroll
        for (int k = 0; k < d; k++) {
            const double v = warp_sum(gx[q][k]);
            if (tx == 0 && i < n) px[i * d + k] = v;
        }
    }
    // scalar partials: warp tree, then serial over the 8 warps
    const int ncol = 2 * d + 2;
    for (int k = 0; k < d; k++) {
        const double v = warp_sum(gl[k]);
        if (tx == 0) red[ty][k] = v;
    }
    {
        const double v = warp_sum(tr);
        if (tx == 0) red[ty][d] = v;
    }
    for (int k = 0; k <= d; k++) {
        const double v = warp_sum(gc[k]);
        if (tx == 0) red[ty][d + 1 + k] = v;
    }
    __syncthreads();
    if (t < ncol) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += red[w][t];
        part[((long long)blockIdx.x * NS + split) * ncol + t] = v;
    }
}

// fixed-order reductions of the partials:
//   gX[i][k] = sum_s partX[s][i][k]  +  sum over the strips I that hold rows of i's class, I <= i / 32, of partC[I][i][k]
__global__ void kernel_grad_reduce_x_kernel(const double* __restrict__ partX, const double* __restrict__ partC, long long n,
                                            int d, int NS, long long nblk, const int64_t* __restrict__ offs, int n_classes,
                                            double* __restrict__ gX) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const long long i = idx / d;
    const int k = (int)(idx - i * d);
    double v = 0.0;
    for (int s = 0; s < NS; s++) v += partX[(long long)s * n * d + idx];
    const long long I_lo = offs ? offs[class_of_row(i, offs, n_classes)] / 32 : 0, I_hi = i / 32;
    for (long long I = I_lo; I <= I_hi; I++) v += partC[(partc_offset(I, nblk) + (i - 32 * I)) * d + k];
    gX[idx] = v;
}

__global__ void kernel_grad_final_kernel(const double* __restrict__ part, long long nblk, int d, int kind,
                                         const double* __restrict__ lin_c2, double sigma2, double* g_log_ls,
                                         double* g_log_sigma, double* g_log_c) {
    const int ncol = 2 * d + 2;
    const int col = threadIdx.x;
    if (col >= ncol) return;
    double v = 0.0;
    for (long long b = 0; b < nblk; b++) v += part[b * ncol + col];
    if (col < d) {
        if (g_log_ls) g_log_ls[col] = v;  // sum S k 2 dist_k over unordered pairs (dk already divided by l_k)
    } else if (col == d) {
        if (g_log_sigma) g_log_sigma[0] = 2.0 * sigma2 * v;
    } else if (kind == 1 && g_log_c) {
        const int k = col - d - 1;
        g_log_c[k] = 2.0 * lin_c2[k] * v;
    }
}

constexpr int GRAD_SPLITS = 8;

}  // namespace gpmdm

using namespace gpmdm;

extern "C" int gpmdm_kernel_build_f64(const double* X, int64_t n, int32_t d, int32_t kind, const double* lengthscales,
                                      const double* lin_c2, double noise2, const int64_t* class_offsets,
                                      int32_t n_classes, double* K, void* stream) {
    GPMDM_REQUIRE(X && lengthscales && K, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && d >= 1 && d <= MAXD_T, GPMDM_E_INVALID, "bad sizes n=%lld d=%d", (long long)n, d);
    GPMDM_REQUIRE(kind == 0 || (kind == 1 && lin_c2), GPMDM_E_INVALID, "kind 1 needs lin_c2");
    GPMDM_REQUIRE(class_offsets == nullptr || n_classes >= 1, GPMDM_E_INVALID, "bad class offsets");
    const unsigned g = (unsigned)((n + BT - 1) / BT);
    const dim3 grid(g, g);
    cudaStream_t st = (cudaStream_t)stream;
#define GPMDM_BUILD_CASE(DL)                                                                                          \
    case DL:                                                                                                          \
        if (kind == 1)                                                                                                \
            kernel_build_kernel<DL, 1><<<grid, 256, 0, st>>>(X, n, lengthscales, lin_c2, noise2, class_offsets, n_classes, K); \
        else                                                                                                          \
            kernel_build_kernel<DL, 0><<<grid, 256, 0, st>>>(X, n, lengthscales, lin_c2, noise2, class_offsets, n_classes, K); \
        break;
    switch (d) {
        GPMDM_BUILD_CASE(1) GPMDM_BUILD_CASE(2) GPMDM_BUILD_CASE(3) GPMDM_BUILD_CASE(4)
        GPMDM_BUILD_CASE(5) GPMDM_BUILD_CASE(6) GPMDM_BUILD_CASE(7) GPMDM_BUILD_CASE(8)
    }
#undef GPMDM_BUILD_CASE
    return check_launch("kernel_build_kernel");
}

extern "C" int64_t gpmdm_kernel_grad_workspace_bytes(int64_t n, int32_t d) {
    const int64_t nblk = (n + 31) / 32;
    return nblk * GRAD_SPLITS * (int64_t)(2 * d + 2) * 8 + (int64_t)GRAD_SPLITS * n * d * 8 +
           partc_offset(nblk, nblk) * d * 8;
}

extern "C" int gpmdm_kernel_grad_f64(const double* X, const double* G, int64_t n, int32_t d, int32_t kind,
                                     const double* lengthscales, const double* lin_c2, double sigma2,
                                     const int64_t* class_offsets, int32_t n_classes, double* gX, double* g_log_ls,
                                     double* g_log_sigma, double* g_log_c, void* workspace, void* stream) {
    GPMDM_REQUIRE(X && G && lengthscales && gX && workspace, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && d >= 1 && d <= MAXD_T, GPMDM_E_INVALID, "bad sizes n=%lld d=%d", (long long)n, d);
    GPMDM_REQUIRE(kind == 0 || (kind == 1 && lin_c2), GPMDM_E_INVALID, "kind 1 needs lin_c2");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nblk = (n + 31) / 32;
    double* part = static_cast<double*>(workspace);
    double* partX = part + nblk * GRAD_SPLITS * (2 * d + 2);
    double* partC = partX + (long long)GRAD_SPLITS * n * d;
    const dim3 grid((unsigned)nblk, GRAD_SPLITS), block(32, 8);
#define GPMDM_GRAD_CASE(DL)                                                                                        \
    case DL:                                                                                                       \
        kernel_grad_kernel<DL><<<grid, block, 0, st>>>(X, G, n, kind, lengthscales, lin_c2, class_offsets, n_classes, \
                                                       partX, partC, part);                                        \
        break;
    switch (d) {
        GPMDM_GRAD_CASE(1) GPMDM_GRAD_CASE(2) GPMDM_GRAD_CASE(3) GPMDM_GRAD_CASE(4)
        GPMDM_GRAD_CASE(5) GPMDM_GRAD_CASE(6) GPMDM_GRAD_CASE(7) GPMDM_GRAD_CASE(8)
    }
#undef GPMDM_GRAD_CASE
    kernel_grad_reduce_x_kernel<<<(unsigned)((n * d + 255) / 256), 256, 0, st>>>(partX, partC, n, d, GRAD_SPLITS, nblk,
                                                                                class_offsets, n_classes, gX);
    kernel_grad_final_kernel<<<1, 32, 0, st>>>(part, nblk * GRAD_SPLITS, d, kind, lin_c2, sigma2, g_log_ls, g_log_sigma,
                                               g_log_c);
    return check_launch("kernel_grad_kernel");
}
