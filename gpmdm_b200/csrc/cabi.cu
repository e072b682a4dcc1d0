// Error plumbing, version, factor packing and the DMMA peak probe of libgpmdm_sm100a.so.
#include <stdarg.h>

#include "common.cuh"

namespace gpmdm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// Q[i][j] = Kinv[i][j] + Kinv[j][i] (i > j) | Kinv[i][i] (i == j) | 0 (i < j)   -- tri
// Q[i][j] = Kinv[i][j]                                                         -- dense
// written into the column-panel layout of gpmdm_gp_block (the buffer is zero-filled by the caller: padding, zero
// upper parts of the diagonal blocks).  32 x 32 tiles through shared memory so that both the row and the transposed
// read coalesce.
__global__ void pack_quadform_kernel(const double* __restrict__ Kinv, long long n, long long n_pad, int tri,
                                     double* __restrict__ L) {
    __shared__ double tT[32][33];
    const long long i0 = (long long)blockIdx.y * 32, j0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    if (tri && j0 > i0 + 31) return;               // strictly upper tile: zeros (already there, or not stored at all)
    if (tri) {
        for (int r = ty; r < 32; r += 8) {  // transposed tile: rows j0.., cols i0..
            const long long jj = j0 + r, ii = i0 + tx;
            tT[r][tx] = (jj < n && ii < n) ? Kinv[jj * n + ii] : 0.0;
        }
        __syncthreads();
    }
    const int t = (int)(j0 / GPMDM_TILE_N);  // column panel (a 32-wide tile never straddles two panels)
    const long long kb = panel_first_row(t, tri);
    double* panel = L + panel_row_offset(t, n_pad, tri) * GPMDM_PANEL_LD;
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r, j = j0 + tx;
        if (i < kb || i >= n || j >= n) continue;
        const double a = Kinv[i * n + j];
        double v = 0.0;
        if (!tri) v = a;
        else if (i > j) v = a + tT[tx][r];
        else if (i == j) v = a;
        panel[(i - kb) * GPMDM_PANEL_LD + (j - (long long)t * GPMDM_TILE_N)] = v;
    }
}

// alpha panels [alpha_ld / 256][n_pad][260] from A [n, dout] row-major (buffer zero-filled by the caller)
__global__ void pack_alpha_kernel(const double* __restrict__ A, long long n, long long n_pad, int dout,
                                  double* __restrict__ alpha) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * dout) return;
    const long long i = idx / dout;
    const int j = (int)(idx - i * dout), t = j / GPMDM_TILE_N;
    alpha[((long long)t * n_pad + i) * GPMDM_PANEL_LD + (j - t * GPMDM_TILE_N)] = A[idx];
}

// ---- DMMA peak probe --------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) dmma_probe_kernel(double* out, int iters, double seed) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = 0.0;
    const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma_m8n8k4(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

}  // namespace gpmdm

using namespace gpmdm;

extern "C" int gpmdm_abi_version(void) { return GPMDM_ABI_VERSION; }
extern "C" const char* gpmdm_last_error(void) { return g_err; }

extern "C" int64_t gpmdm_quadform_bytes(int64_t n_pad, int tri) {
    return panel_row_offset(n_pad / GPMDM_TILE_N, n_pad, tri) * GPMDM_PANEL_LD * 8;
}

extern "C" int gpmdm_pack_quadform_f64(const double* Kinv, int64_t n, int64_t n_pad, int tri, double* L,
                                       void* stream) {
    GPMDM_REQUIRE(Kinv && L, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && n_pad >= n && n_pad % GPMDM_TILE_N == 0, GPMDM_E_INVALID,
                  "n_pad %lld must be a multiple of %d and >= n %lld", (long long)n_pad, GPMDM_TILE_N, (long long)n);
    cudaError_t e = cudaMemsetAsync(L, 0, (size_t)gpmdm_quadform_bytes(n_pad, tri), (cudaStream_t)stream);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)(n_pad / 32), (unsigned)(n_pad / 32)), block(32, 8);
    pack_quadform_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(Kinv, n, n_pad, tri, L);
    return check_launch("pack_quadform_kernel");
}

extern "C" int64_t gpmdm_alpha_bytes(int64_t n_pad, int32_t alpha_ld) {
    return (int64_t)(alpha_ld / GPMDM_TILE_N) * n_pad * GPMDM_PANEL_LD * 8;
}

extern "C" int gpmdm_pack_alpha_f64(const double* A, int64_t n, int64_t n_pad, int32_t dout, int32_t alpha_ld,
                                    double* alpha, void* stream) {
    GPMDM_REQUIRE(A && alpha, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && n_pad >= n && n_pad % GPMDM_TILE_N == 0 && dout >= 1 && alpha_ld >= dout &&
                      alpha_ld % GPMDM_TILE_N == 0,
                  GPMDM_E_INVALID, "bad sizes n=%lld n_pad=%lld dout=%d alpha_ld=%d", (long long)n, (long long)n_pad,
                  dout, alpha_ld);
    cudaError_t e = cudaMemsetAsync(alpha, 0, (size_t)gpmdm_alpha_bytes(n_pad, alpha_ld), (cudaStream_t)stream);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const long long total = n * dout;
    pack_alpha_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, n, n_pad, dout, alpha);
    return check_launch("pack_alpha_kernel");
}

extern "C" int gpmdm_probe_dmma_tflops(int32_t iters, double* tflops_host) {
    GPMDM_REQUIRE(tflops_host && iters > 0, GPMDM_E_INVALID, "bad argument");
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    GPMDM_REQUIRE(e == cudaSuccess, GPMDM_E_NODEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out = nullptr;
    e = cudaMalloc(&out, 8);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMalloc: %s", cudaGetErrorString(e));
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int blocks = sms * 2, threads = 512;  // 32 warps per SM
    dmma_probe_kernel<<<blocks, threads>>>(out, iters / 4 + 1, 1.0);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(t0);
        dmma_probe_kernel<<<blocks, threads>>>(out, iters, 1.0);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(out);
    int rc = check_launch("dmma_probe_kernel");
    if (rc) return rc;
    const double flops = 2.0 * 256.0 * 8.0 * (double)iters * (threads / 32.0) * blocks;
    *tflops_host = flops / (best * 1e-3) / 1e12;
    return 0;
}
