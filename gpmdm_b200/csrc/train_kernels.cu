// Training-side kernel-matrix construction and the NLL gradient terms through it.
//
// Replaces (reference paths): gpmdm/gpmdm.py:381-548 (get_y_kernel / get_x_kernel / get_rbf_kernel /
// get_weighted_distances / get_lin_kernel), the dense 0/1 class masks of :311-378 (`* self.M`,
// :616, :1292) and the autograd backward through those ops.  The O(N^3) factorisation stays with
// torch.linalg (cuSOLVER); these kernels are HBM-bound: the build writes 8 N^2 bytes, the gradient
// reads G twice (16 N^2 bytes: once row-wise, once as coalesced 256-byte column segments).
#include <math.h>

#include "common.cuh"

namespace gpmdm {

constexpr int MAXD_T = GPMDM_MAX_LATENT;

__device__ __forceinline__ int class_of_row(long long i, const int64_t* __restrict__ offs, int n_classes) {
    int c = 0;
    while (c + 1 < n_classes && i >= offs[c + 1]) c++;
    return c;
}

// One 32 x 32 tile per block (32 x 8 threads); masked-out tiles are written as zeros without any math.
__global__ void __launch_bounds__(256) kernel_build_kernel(const double* __restrict__ X, long long n, int d, int kind,
                                                           const double* __restrict__ ls,
                                                           const double* __restrict__ lin_c2, double noise2,
                                                           const int64_t* __restrict__ offs, int n_classes,
                                                           double* __restrict__ K) {
    __shared__ double ai[32][MAXD_T + 1], aj[32][MAXD_T + 1];  // x / l and |x/l|^2
    __shared__ double xi[32][MAXD_T], xj[32][MAXD_T];
    __shared__ int ci[32], cj[32];
    const long long i0 = (long long)blockIdx.y * 32, j0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 32 + tx;
    if (t < 64) {
        const bool is_i = t < 32;
        const int r = t & 31;
        const long long row = (is_i ? i0 : j0) + r;
        double n2 = 0.0;
        for (int k = 0; k < d; k++) {
            const double x = row < n ? X[row * d + k] : 0.0;
            const double a = x / ls[k];
            (is_i ? ai : aj)[r][k] = a;
            (is_i ? xi : xj)[r][k] = x;
            n2 = fma(a, a, n2);
        }
        (is_i ? ai : aj)[r][d] = n2;
        (is_i ? ci : cj)[r] = (offs && row < n) ? class_of_row(row, offs, n_classes) : 0;
    }
    __syncthreads();
    const long long j = j0 + tx;
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r;
        if (i >= n || j >= n) continue;
        double v = 0.0;
        if (!offs || ci[r] == cj[tx]) {
            double dot = 0.0;
            for (int k = 0; k < d; k++) dot = fma(ai[r][k], aj[tx][k], dot);
            v = exp(-(ai[r][d] + aj[tx][d] - 2.0 * dot));  // gpmdm.py:515-517 expansion form
            if (i == j) v += noise2;
            if (kind == 1) {
                double lin = lin_c2[d];
                for (int k = 0; k < d; k++) lin = fma(lin_c2[k] * xi[r][k], xj[tx][k], lin);
                v += lin;
            }
        }
        K[i * n + j] = v;
    }
}

// Gradient terms.  Block b owns rows [32 b, 32 b + 32) and walks all column tiles.
//   part [nblk][2 d + 2] : per-block partials of g_log_ls[d], tr(G^), g_log_c[d+1]
__global__ void __launch_bounds__(256) kernel_grad_kernel(const double* __restrict__ X, const double* __restrict__ G,
                                                          long long n, int d, int kind, const double* __restrict__ ls,
                                                          const double* __restrict__ lin_c2,
                                                          const int64_t* __restrict__ offs, int n_classes,
                                                          double* __restrict__ gX, double* __restrict__ part) {
    __shared__ double ai[32][MAXD_T], aj[32][MAXD_T];
    __shared__ double xi[32][MAXD_T], xj[32][MAXD_T];
    __shared__ double GT[32][33];
    __shared__ int ci[32], cj[32];
    __shared__ double red[8][2 * MAXD_T + 2];
    const long long i0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 32 + tx;
    if (t < 32) {
        const long long row = i0 + t;
        for (int k = 0; k < d; k++) {
            const double x = row < n ? X[row * d + k] : 0.0;
            xi[t][k] = x;
            ai[t][k] = x / ls[k];
        }
        ci[t] = (offs && row < n) ? class_of_row(row, offs, n_classes) : 0;
    }
    double gx[4][MAXD_T];
    double gl[MAXD_T], gc[MAXD_T + 1], tr = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < MAXD_T; k++) gx[q][k] = 0.0;
#pragma unroll
    for (int k = 0; k < MAXD_T; k++) gl[k] = gc[k] = 0.0;
    gc[MAXD_T] = 0.0;

    // class-masked gradients vanish outside the class block: restrict the column walk to it
    long long jbeg = 0, jend = n;
    __syncthreads();
    if (offs) {
        // rows of one block may straddle two classes; walk the union of their column ranges
        const long long last = (i0 + 31 < n ? i0 + 31 : n - 1);
        jbeg = offs[class_of_row(i0, offs, n_classes)];
        jend = offs[class_of_row(last, offs, n_classes) + 1];
        jbeg = jbeg / 32 * 32;
    }
    for (long long j0 = jbeg; j0 < jend; j0 += 32) {
        __syncthreads();
        if (t < 32) {
            const long long row = j0 + t;
            for (int k = 0; k < d; k++) {
                const double x = row < n ? X[row * d + k] : 0.0;
                xj[t][k] = x;
                aj[t][k] = x / ls[k];
            }
            cj[t] = (offs && row < n) ? class_of_row(row, offs, n_classes) : 0;
        }
        // transposed tile: GT[r][c] = G[j0 + r][i0 + c]  (32 rows of 256 contiguous bytes)
        for (int r = ty; r < 32; r += 8) {
            const long long jj = j0 + r, ii = i0 + tx;
            GT[r][tx] = (jj < n && ii < n) ? G[jj * n + ii] : 0.0;
        }
        __syncthreads();
        const long long j = j0 + tx;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int r = ty + 8 * q;
            const long long i = i0 + r;
            if (i >= n || j >= n) continue;
            if (offs && ci[r] != cj[tx]) continue;
            const double g = G[i * n + j];  // G^_ij
            const double s = g + GT[tx][r];  // S_ij = G^_ij + G^_ji
            double dist = 0.0;
            double dk[MAXD_T];
#pragma unroll
            for (int k = 0; k < MAXD_T; k++)
                if (k < d) {
                    dk[k] = ai[r][k] - aj[tx][k];
                    dist = fma(dk[k], dk[k], dist);
                }
            const double kr = exp(-dist);
#pragma unroll
            for (int k = 0; k < MAXD_T; k++)
                if (k < d) {
                    // d k_rbf / d x_ik = k_rbf * (-2 (x_ik - x_jk) / l_k^2)
                    double term = kr * (-2.0 * dk[k] / ls[k]);
                    if (kind == 1) term = fma(lin_c2[k], xj[tx][k], term);
                    gx[q][k] = fma(s, term, gx[q][k]);
                    gl[k] = fma(g * kr, 2.0 * dk[k] * dk[k], gl[k]);
                    if (kind == 1) gc[k] = fma(g, xi[r][k] * xj[tx][k], gc[k]);
                }
            if (kind == 1) gc[d] += g;
            if (i == j) tr += g;
        }
    }
    // row sums: lanes of a warp share ty, i.e. the same four rows
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const long long i = i0 + ty + 8 * q;
#pragma unroll
        for (int k = 0; k < MAXD_T; k++)
            if (k < d) {
                const double v = warp_sum(gx[q][k]);
                if (tx == 0 && i < n) gX[i * d + k] = v;
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
        part[(long long)blockIdx.x * ncol + t] = v;
    }
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
        if (g_log_ls) g_log_ls[col] = v;  // sum G^ k 2 dist_k  (dk already divided by l_k)
    } else if (col == d) {
        if (g_log_sigma) g_log_sigma[0] = 2.0 * sigma2 * v;
    } else if (kind == 1 && g_log_c) {
        const int k = col - d - 1;
        g_log_c[k] = 2.0 * lin_c2[k] * v;
    }
}

}  // namespace gpmdm

using namespace gpmdm;

extern "C" int gpmdm_kernel_build_f64(const double* X, int64_t n, int32_t d, int32_t kind, const double* lengthscales,
                                      const double* lin_c2, double noise2, const int64_t* class_offsets,
                                      int32_t n_classes, double* K, void* stream) {
    GPMDM_REQUIRE(X && lengthscales && K, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && d >= 1 && d <= MAXD_T, GPMDM_E_INVALID, "bad sizes n=%lld d=%d", (long long)n, d);
    GPMDM_REQUIRE(kind == 0 || (kind == 1 && lin_c2), GPMDM_E_INVALID, "kind 1 needs lin_c2");
    GPMDM_REQUIRE(class_offsets == nullptr || n_classes >= 1, GPMDM_E_INVALID, "bad class offsets");
    const unsigned g = (unsigned)((n + 31) / 32);
    kernel_build_kernel<<<dim3(g, g), dim3(32, 8), 0, (cudaStream_t)stream>>>(X, n, d, kind, lengthscales, lin_c2,
                                                                             noise2, class_offsets, n_classes, K);
    return check_launch("kernel_build_kernel");
}

extern "C" int64_t gpmdm_kernel_grad_workspace_bytes(int64_t n, int32_t d) {
    return ((n + 31) / 32) * (int64_t)(2 * d + 2) * 8;
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
    kernel_grad_kernel<<<(unsigned)nblk, dim3(32, 8), 0, st>>>(X, G, n, d, kind, lengthscales, lin_c2, class_offsets,
                                                                n_classes, gX, part);
    kernel_grad_final_kernel<<<1, 32, 0, st>>>(part, nblk, d, kind, lin_c2, sigma2, g_log_ls, g_log_sigma, g_log_c);
    return check_launch("kernel_grad_kernel");
}
