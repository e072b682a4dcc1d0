// exp(x) for x <= ~0 in double precision, shaped for the K* prologue of the predict kernel, where every
// non-MMA fp64 instruction steals a slot of the one fp64 datapath the DMMAs run on:
//   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128
//   exp(x) = 2^m * 2^(j/64) * exp(r);   2^(j/64) from a 64-entry table (shared memory), exp(r) by a degree-6
//   Taylor polynomial (truncation 1.4e-19 relative), 2^m by adding m to the exponent field of the table entry
//   (integer pipe).  11 fp64-pipe instructions instead of ~20 for a table-free evaluation.
// Results below 2^-1021 are flushed to 0; arguments below -800 are clamped on the integer pipe.
// Max error vs libm on [-745, 1e-8]: 2 ulp (tools/check_fast_exp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpmdm {

struct ExpConst {
    static constexpr double INV = 0x1.71547652b82fep+6;       // 64 / ln 2
    static constexpr double LN2_64_HI = 0x1.62e42ff000000p-7;  // 32 significant bits: n * HI is exact
    static constexpr double LN2_64_LO = -0x1.718432a1b0e26p-41;
    static constexpr double MAGIC = 6755399441055744.0;        // 1.5 * 2^52
};

#define GPMDM_EXP_TABLE_VALUES                                                                              \
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,               \
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,               \
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,               \
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,               \
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,               \
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,               \
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,               \
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,               \
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,               \
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,               \
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,               \
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,               \
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,               \
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,               \
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,               \
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0

__host__ __device__ __forceinline__ int64_t exp_bits(double v) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(v);
#else
    union { double d; int64_t i; } u; u.d = v; return u.i;
#endif
}
__host__ __device__ __forceinline__ double exp_from_bits(int64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    union { double d; int64_t i; } u; u.i = b; return u.d;
#endif
}

// Stage 1: range reduction (4 fp64 instructions).  Returns r, writes n = 64 m + j.
__host__ __device__ __forceinline__ double exp_reduce(double x, int& n) {
    // clamp x >= -800 with integer compares: for negative doubles a larger high word means a smaller value
    const uint32_t hi = (uint32_t)(exp_bits(x) >> 32);
    // (up to and including -inf = 0xFFF00000:00000000; NaNs lie above it and must propagate, as they do through the
    // reference's torch.exp: a NaN particle state gives NaN predictions, gpmdm_pf.py:168 -> :189)
    if (hi > 0xC0890000u && hi <= 0xFFF00000u) x = -800.0;
    const double tn = fma(x, ExpConst::INV, ExpConst::MAGIC);
    n = (int)(uint32_t)exp_bits(tn);
    const double nf = tn - ExpConst::MAGIC;
    double r = fma(nf, -ExpConst::LN2_64_HI, x);
    r = fma(nf, -ExpConst::LN2_64_LO, r);
    return r;
}

// Stage 2: exp(r), |r| <= 0.0055 (6 fp64 instructions)
__host__ __device__ __forceinline__ double exp_poly(double r) {
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p;
}

// Stage 3: 2^m * table[j] * p (1 fp64 instruction)
__host__ __device__ __forceinline__ double exp_scale(double p, int n, const double* __restrict__ table) {
    const int m = n >> 6;
    const double t = exp_from_bits(exp_bits(table[n & 63]) + ((int64_t)m << 52));
    return m < -1021 ? 0.0 : t * p;
}

__host__ __device__ __forceinline__ double fast_exp(double x, const double* __restrict__ table) {
    int n;
    const double r = exp_reduce(x, n);
    return exp_scale(exp_poly(r), n, table);
}

}  // namespace gpmdm
