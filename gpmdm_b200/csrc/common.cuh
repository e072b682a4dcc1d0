// Shared device/host helpers for libgpmdm_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gpmdm_b200.h"

namespace gpmdm {

// ---- error plumbing (no exceptions across the C ABI) ----------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

// the low-latency predict launches (gp_predict.cu) for callers inside the library that issue a fixed launch sequence:
// counters_zero = tile_counter[0..1] are already zero on the stream (every finalize kernel leaves them zeroed)
int observe_lowlat_impl(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                        const double* v_in, double* ll, double* mu_out, double* v_out, int64_t max_n_pad, int32_t seg_chunks,
                        int32_t* tile_counter, void* workspace, void* stream, bool counters_zero);
int propagate_lowlat_impl(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm, const int32_t* tiles,
                          const int32_t* n_tiles, int64_t P, const double* eps, double* x_new, double* mean_out,
                          double* var_out, int64_t max_n_pad, int32_t seg_chunks, int32_t* tile_counter, void* workspace,
                          void* stream, bool counters_zero);

#define GPMDM_REQUIRE(cond, code, ...)      \
    do {                                    \
        if (!(cond)) {                      \
            gpmdm::set_error(__VA_ARGS__);  \
            return (code);                  \
        }                                   \
    } while (0)

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// non-blocking probe of a phase: the predicate can be consumed much later, so the probe's latency hides under other work
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// warpgroup register re-allocation (sm_90a+): producers give registers to the MMA warpgroups
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// fp64 tensor-core MMA (DMMA), D[8x8] += A[8x4] * B[4x8].
//   a : A[row = lane>>2][col = lane&3]      b : B[k = lane&3][n = lane>>2]
//   c0,c1 : C[row = lane>>2][col = 2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// first stored row of column panel t of L, and the row offset of that panel in the packed array (gpmdm_gp_block)
__host__ __device__ __forceinline__ long long panel_first_row(int t, int tri) { return tri ? (long long)t * GPMDM_TILE_N : 0; }
__host__ __device__ __forceinline__ long long panel_row_offset(long long t, long long n_pad, int tri) {
    return tri ? t * n_pad - (GPMDM_TILE_N / 2) * t * (t - 1) : t * n_pad;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace gpmdm
