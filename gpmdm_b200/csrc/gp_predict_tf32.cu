// tf32 variant of the observation-GP kernel (BASELINE config 4): tcgen05 tensor cores, TMEM accumulators.
//
// Replaces the same reference lines as the fp64 kernel (gpmdm/gpmdm.py:923-963 + gpmdm_pf.py:188-192) for
// callers that accept ~1e-4 relative accuracy.  It is NOT the fp64 kernel with narrower types:
//   * the explicit-inverse form 1 - k^T K^-1 k is unusable in fp32 (K^-1 entries reach 1/sigma_n^2, rounding
//     exceeds the variance), so the variance uses the whitened form  v = 1 - |W k|^2,  W = U^-T from the
//     reference's upper Cholesky factor K = U^T U (gpmdm.py:1287-1288) -- positive terms bounded by the prior;
//   * plain tf32 inputs (10-bit mantissa) would cap accuracy at ~5e-4, so every product is error-compensated
//     ("3xTF32"):  a b ~ a_hi b_hi + a_lo b_hi + a_hi b_lo  with a_hi = tf32(a), a_lo = tf32(a - a_hi): three
//     tcgen05.mma per k-step into one fp32 TMEM accumulator, ~2^-21 relative per product.
//
// Shape: per 128-particle tile and 256-column tile J of W (lower triangular: k < 256 (J+1)),
//     U[p, n] = sum_k K*[p, k] W[n, k]      (tcgen05.mma M=128, N=256, K=8, kind::tf32, A and B from shared memory)
//     q[p]   += sum_n U[p, n]^2             (epilogue: tcgen05.ld, one particle row per thread)
// then one more tile with B = alpha^T for the mean and the fused log-likelihood.
//
// Warp roles (512 threads, one CTA per SM):
//   warp 0       TMA producer: one cp.async.bulk (32 KB) per k-chunk of pre-packed W tiles -> B ring (3 stages)
//   warp 1       MMA issuer (one elected lane): 6 tcgen05.mma per chunk, tcgen05.commit -> mbarriers; owns TMEM
//   warps 4-11   K* generators: exp2-based RBF in fp32 from latent coordinates, hi/lo split, written straight into
//                the UMMA canonical (core-matrix, no-swizzle, K-major) layout of the A ring (4 stages)
//   warps 12-15  epilogue: TMEM -> registers, sum of squares / mean / log-likelihood; double-buffered accumulator
// The operand tiles are packed on the host side once (gpmdm_pack_whitened_tf32) in exactly the byte order the
// tensor core reads, so a 1-D bulk copy lands them ready to use (no tensor maps, no swizzle).
//
// MODE_F16X2 -- the same kernel with every operand split into TWO fp16 pieces instead of the error-compensated tf32
// triple:  a = a_hi + 2^-11 a_lo,  a_hi = fp16(a),  a_lo = fp16((a - a_hi) 2^11)  (the scaling keeps the residual inside
// fp16's narrow exponent range: K* lies in [0, 1], |W| <= 1/sigma_n),
//     a b ~ a_hi b_hi + 2^-11 (a_lo b_hi + a_hi b_lo)          -- 22 significant bits, as 3xTF32
// i.e. three kind::f16 MMAs per k-step of 16 (twice the k of a tf32 step, at the same cycle count): half the tensor-pipe
// time, half the operand bytes per product.  The two terms go to separate fp32 TMEM accumulators (D1, D2 = 2 x 256
// columns, so the accumulator is single-buffered) and are combined in the epilogue.  Variance only; means and the
// log-likelihood stay with the fp64 mean tile (gpmdm_pf_loglik_f64), as in the hybrid tf32 variant.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace gpmdm {
namespace tf32 {

constexpr int MODE_TF32X3 = 0, MODE_F16X2 = 1;
constexpr int TM = 128;      // particles per CTA tile (UMMA M)
constexpr int TN = 256;      // columns per column tile (UMMA N)
#ifndef GPMDM_TC_ASTAGES
#define GPMDM_TC_ASTAGES 4
#endif
#ifndef GPMDM_TC_BSTAGES
#define GPMDM_TC_BSTAGES 3
#endif
constexpr int ASTAGES = GPMDM_TC_ASTAGES;  // ring depths of the single-CTA variant (tuning experiments: -DGPMDM_TC_ASTAGES=..)
constexpr int BSTAGES = GPMDM_TC_BSTAGES;
constexpr int A_HALF_BYTES = TM * 16 * 4;   // bytes in the hi (or lo) part of an A stage: 8 KB in both modes
constexpr int B_HALF_BYTES = TN * 16 * 4;   // bytes in the hi (or lo) part of a B stage / packed tile: 16 KB
constexpr int NGEN = 256;                // generator threads
constexpr int NEPI = 128;                // epilogue threads
constexpr int NTHREADS = 512;
constexpr int GEN_WARP0 = 4, EPI_WARP0 = 12;
constexpr int CREC = 8;                  // floats per training record (a_i padded to 8)
// k per chunk: 16 tf32 (2 MMA k-steps of 8) or 32 fp16 (2 k-steps of 16) -- 64 operand bytes per row either way, i.e.
// four 16-byte core-matrix rows along K, so stages, tiles and descriptors have the same byte geometry in both modes
template <int MODE> struct Geo {
    static constexpr int ESIZE = MODE == MODE_F16X2 ? 2 : 4;
    static constexpr int KC = 64 / ESIZE;
    static constexpr int KSTEP = 32 / ESIZE;   // k per MMA
    static constexpr int CM = 16 / ESIZE;      // elements per core-matrix row
};
constexpr int KC = Geo<MODE_TF32X3>::KC;       // (tf32 names kept for the packers and the probe below)
constexpr int A_HALF = A_HALF_BYTES / 4, B_HALF = B_HALF_BYTES / 4;
constexpr uint32_t LBO = 128, SBO = 4 * 128;  // core-matrix strides (bytes) along K and along M/N

// CG = CTAs per tensor-core tile group: 1, or 2 = a cluster of two CTAs (SM pair) issuing cta_group::2 MMAs -- each CTA
// supplies its own 128 particles (A) and HALF of every W tile (B rows 0-127 / 128-255): half the B bytes per SM from L2 and
// from shared memory, the two resources the single-CTA kernel runs out of (profiles/ncu_observe_f16x2_kernel_r02.txt).
// Ring depths: the pair variant's producer -> consumer path crosses the cluster (local barrier -> relay -> the leader's
// barrier, and the multicast commit back), a round trip of several chunk times: with the single-CTA depths (4 / 3) the
// pair ran 8-19 % SLOWER than the single CTAs; the half-size B stages leave room for 8 / 5.
template <int CG>
struct Ring {
    static constexpr int A = CG == 2 ? 8 : ASTAGES;
    static constexpr int B = CG == 2 ? 5 : BSTAGES;
};
template <int CG>
struct __align__(1024) SmemT {
    unsigned char A[Ring<CG>::A][2][A_HALF_BYTES];        // [stage][hi|lo]
    unsigned char B[Ring<CG>::B][2][B_HALF_BYTES / CG];   // [stage][hi|lo], this CTA's share of the 256 rows
    float coords[2][32 * CREC];                           // the generators' double buffer: one chunk of training records
    uint64_t a_full[Ring<CG>::A], a_empty[Ring<CG>::A], b_full[Ring<CG>::B], b_empty[Ring<CG>::B], t_full[2], t_empty[2];
    uint32_t tmem_base;
};
using Smem = SmemT<1>;

struct Params {
    const float* coords;   // [n_pad / 2, CREC, 2]: pairs of training rows interleaved per coordinate, x sqrt(log2 e) / l
    const float* wtiles;   // packed W tiles
    const float* atiles;   // packed alpha tiles
    int n_pad, d, dout;
    const double* ls;      // [d]
    const double* lam2;    // [dout]
    const double* x;       // [P, d]
    long long P;
    const double* z;
    double ll_const;
    double* ll;
    double* mu_out;
    double* v_out;
    int32_t* round_counter;  // device scratch (zeroed by the call), NULL = no round synchronisation
    int32_t* status;         // particles whose variance was not a positive finite number (NULL = not counted)
    // KIND 1 (dynamics GP): class-homogeneous 128-particle tiles over per-class blocks
    const gpmdm_tc_block* dblocks;  // device, one per class
    const int32_t* perm;            // particles ordered by (class, index)
    const int32_t* tiles;           // {block, first position in perm, count, 0} per tile
    const int32_t* n_tiles_dev;     // device int
};

// element (row, k) of a [rows x KC] operand tile in the canonical no-swizzle K-major layout, in elements
template <int MODE = MODE_TF32X3>
__host__ __device__ __forceinline__ int tile_index(int row, int k) {
    constexpr int CM = Geo<MODE>::CM, ES = Geo<MODE>::ESIZE;
    return (row >> 3) * (SBO / ES) + (k / CM) * (LBO / ES) + (row & 7) * CM + (k % CM);
}

__device__ __forceinline__ uint64_t smem_desc(const void* p) {
    // SM100 shared-memory matrix descriptor: start address, leading (K) / stride (M,N) byte offsets in 16-byte
    // units, version 1, no swizzle
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(p) >> 4) & 0x3fff);
    d |= (uint64_t)((LBO >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = f32 (bits 4-5 = 1), A / B format (bits 7-9 / 10-12: tf32 = 2 for kind::tf32, f16 = 0 for
// kind::f16), both K-major, N = 256 (bits 17-22 = N >> 3), M = 128 (bits 24-28 = M >> 4)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC_F16), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// ---- cta_group::2 (CTA pair) variants ---------------------------------------------------------------------------------
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
constexpr uint32_t IDESC2_F16 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
template <bool F16>
__device__ __forceinline__ void umma_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    if (F16)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(IDESC2_F16), "r"(accumulate), "r"(0u)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(IDESC2), "r"(accumulate), "r"(0u)
            : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at this offset in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 2^x, one MUFU.EX2 (results below 2^-126 flush to zero: a cross-kernel entry of 1e-38 is zero for every purpose here)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// number of W tiles before column tile J, and in total (KC = k per chunk / per packed tile)
__host__ __device__ __forceinline__ long long wtile_offset_kc(int J, int kc) { return (long long)(TN / kc) * J * (J + 1) / 2; }
__host__ __device__ __forceinline__ long long wtile_offset(int J) { return wtile_offset_kc(J, KC); }

// bounded mbarrier wait for the MODE_F16X2 instance: a protocol bug traps (launch failure) instead of hanging the GPU.
// try_wait suspends the thread for a hardware-defined interval per attempt, so waiting warps do not take issue slots.
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// What one work unit (a 128-particle tile) runs against: the observation GP's single block, or -- KIND 1, the dynamics GP
// (gpmdm.py:1032-1068) -- the block of the tile's class, with the particles reached through the class-sorted permutation.
struct Unit {
    const float* coords;
    const unsigned char* wt;
    int nq, nkc, nct, first, count;
};

template <int DL, int MODE, int CG, int KIND = 0>
__global__ void __launch_bounds__(NTHREADS, 1) observe_tf32_kernel(const Params prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemT<CG>& s = *reinterpret_cast<SmemT<CG>*>(smem_raw);
    using G = Geo<MODE>;
    constexpr int KCm = G::KC;                               // k per chunk in this mode
    constexpr bool F16 = MODE == MODE_F16X2;
    constexpr int AST = Ring<CG>::A, BST = Ring<CG>::B;         // ring depths
    constexpr int BH = B_HALF_BYTES / CG;                    // bytes of one piece (hi or lo) of this CTA's share of a W tile
    // accumulator buffers: two in tf32 mode; one in F16X2 mode, where D1 | D2 fill all 512 TMEM columns
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    static_assert(KIND == 0 || CG == 1, "the two tiles of a CTA pair would belong to different class blocks");
    // + one alpha tile (dout <= 256) when a mean is wanted (tf32 observation mode only: the variants compute variances)
    const int n_tiles = KIND == 1 ? __ldg(prm.n_tiles_dev) : (int)((prm.P + TM - 1) / TM);
    auto unit = [&](int t) {
        Unit un;
        if (KIND == 1) {
            const gpmdm_tc_block b = prm.dblocks[prm.tiles[4 * t]];
            un.coords = b.coords, un.wt = reinterpret_cast<const unsigned char*>(b.wtiles);
            un.nq = (int)(b.n_pad / TN), un.nkc = (int)(b.n_pad / KCm), un.nct = un.nq;
            un.first = prm.tiles[4 * t + 1], un.count = prm.tiles[4 * t + 2];
        } else {
            un.coords = prm.coords, un.wt = reinterpret_cast<const unsigned char*>(prm.wtiles);
            un.nq = prm.n_pad / TN, un.nkc = prm.n_pad / KCm;
            un.nct = un.nq + ((!F16 && (prm.ll || prm.mu_out)) ? 1 : 0);
            un.first = t * TM, un.count = TM;
        }
        return un;
    };
    // work units: particle tiles (CG = 1), or PAIRS of particle tiles 2u, 2u+1 handled by the two CTAs of a cluster
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int n_units = (n_tiles + CG - 1) / CG;
    const int unit0 = (int)blockIdx.x / CG, unit_stride = (int)gridDim.x / CG;
    auto tile_of = [&](int u) { return CG * u + (int)rank; };  // may be == n_tiles for the second CTA of the last pair

    if (tid == 0) {
        for (int i = 0; i < AST; i++) {
            mbar_init(&s.a_full[i], NGEN + ((CG == 2 && leader) ? 1 : 0));  // + the relay of the peer's generators
            mbar_init(&s.a_empty[i], 1);
        }
        for (int i = 0; i < BST; i++) {
            mbar_init(&s.b_full[i], (CG == 2 && leader) ? 2 : 1);           // own TMA + the relay of the peer's TMA
            mbar_init(&s.b_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.t_full[i], 1);
            mbar_init(&s.t_empty[i], NEPI * CG);                            // both CTAs' epilogue warps (leader's copy)
        }
        mbar_fence_init();
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 256-column fp32 accumulators)
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)),
                         "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)),
                         "r"(512u)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();  // the peer's barriers exist before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    auto wait = [](uint64_t* bar, uint32_t parity) {
        if (F16 || CG == 2) mbar_wait_b(bar, parity);
        else mbar_wait(bar, parity);
    };

    // chunk count of column tile ct: W tiles are lower triangular (k < 256 (J+1)); the alpha tile needs all k
    auto chunks_of = [&](const Unit& un, int ct) { return ct < un.nq ? (ct + 1) * (TN / KCm) : un.nkc; };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t g = 0;
            int rounds = 0;
            const unsigned char* at = reinterpret_cast<const unsigned char*>(prm.atiles);
            for (int u = unit0; u < n_units; u += unit_stride) {
                const Unit un = unit(tile_of(u) < n_tiles ? tile_of(u) : n_tiles - 1);
                for (int ct = 0; ct < un.nct; ct++) {
                    const unsigned char* base = ct < un.nq ? un.wt + wtile_offset_kc(ct, KCm) * (2 * B_HALF_BYTES) : at;
                    const int nch = chunks_of(un, ct);
                    for (int kc = 0; kc < nch; kc++, g++) {
                        const int st = g % BST;
                        wait(&s.b_empty[st], ((g / BST) & 1) ^ 1);
                        mbar_expect_tx(&s.b_full[st], 2 * BH);
                        const unsigned char* src = base + (long long)kc * (2 * B_HALF_BYTES);
                        if (CG == 1) {
                            bulk_g2s(&s.B[st][0][0], src, 2 * B_HALF_BYTES, &s.b_full[st]);
                        } else {  // this CTA's 128 of the 256 rows of the hi and of the lo piece: two contiguous ranges
                            bulk_g2s(&s.B[st][0][0], src + rank * BH, BH, &s.b_full[st]);
                            bulk_g2s(&s.B[st][1][0], src + B_HALF_BYTES + rank * BH, BH, &s.b_full[st]);
                        }
                    }
                }
                // Round synchronisation of the producers (bounded polling, never a hang): the 148 CTAs stream the same
                // W tiles; aligned, one HBM read serves all of them out of L2, adrift they would each pull 32 KB per
                // chunk from HBM (148 x 32 KB per ~0.4 us is beyond HBM bandwidth).
                if (prm.round_counter) {
                    rounds++;
                    atomicAdd(prm.round_counter, 1);
                    const int full_rounds = n_units / unit_stride;  // rounds in which every CTA has work
                    if (rounds <= full_rounds) {
                        const int target = rounds * (int)gridDim.x;
                        for (int spin = 0; spin < 200000; spin++) {
                            if (*reinterpret_cast<volatile int*>(prm.round_counter) >= target) break;
                            __nanosleep(100);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (the pair's leader CTA issues for both) =====================
        if (lane == 0 && leader) {
            uint32_t g = 0, tcount = 0;
            for (int u = unit0; u < n_units; u += unit_stride) {
                const Unit un = unit(tile_of(u) < n_tiles ? tile_of(u) : n_tiles - 1);
                for (int ct = 0; ct < un.nct; ct++, tcount++) {
                    const int acc = F16 ? 0 : (tcount & 1);
                    const uint32_t use = F16 ? tcount : (tcount >> 1);     // how often this buffer has been used before
                    wait(&s.t_empty[acc], (use & 1) ^ 1);  // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem + (uint32_t)acc * TN;
                    const int nch = chunks_of(un, ct);
                    for (int kc = 0; kc < nch; kc++, g++) {
                        const int sa = g % AST, sb = g % BST;
                        wait(&s.a_full[sa], (g / AST) & 1);
                        wait(&s.b_full[sb], (g / BST) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 2; ks++) {  // two MMA k-steps per chunk (2 x 8 tf32 / 2 x 16 fp16)
                            const uint32_t koff = ks * 2 * LBO;  // two core matrices (32 bytes of every row) along K per MMA
                            const uint64_t a_hi = smem_desc(&s.A[sa][0][0] + koff);
                            const uint64_t a_lo = smem_desc(&s.A[sa][1][0] + koff);
                            const uint64_t b_hi = smem_desc(&s.B[sb][0][0] + koff);
                            const uint64_t b_lo = smem_desc(&s.B[sb][1][0] + koff);
                            const uint32_t first = (kc | ks) != 0;
                            if (CG == 2) {
                                if (F16) {
                                    umma_2cta<true>(d_tmem, a_hi, b_hi, first);
                                    umma_2cta<true>(d_tmem + TN, a_lo, b_hi, first);
                                    umma_2cta<true>(d_tmem + TN, a_hi, b_lo, 1);
                                } else {
                                    umma_2cta<false>(d_tmem, a_hi, b_hi, first);
                                    umma_2cta<false>(d_tmem, a_lo, b_hi, 1);
                                    umma_2cta<false>(d_tmem, a_hi, b_lo, 1);
                                }
                            } else if (F16) {  // D1 += a_hi b_hi ;  D2 += a_lo b_hi + a_hi b_lo  (scaled by 2^-11 in the epilogue)
                                umma_f16(d_tmem, a_hi, b_hi, first);
                                umma_f16(d_tmem + TN, a_lo, b_hi, first);
                                umma_f16(d_tmem + TN, a_hi, b_lo, 1);
                            } else {
                                umma_tf32(d_tmem, a_hi, b_hi, first);
                                umma_tf32(d_tmem, a_lo, b_hi, 1);
                                umma_tf32(d_tmem, a_hi, b_lo, 1);
                            }
                        }
                        if (CG == 2) {
                            umma_commit_2cta(&s.a_empty[sa]);  // frees the stage in both CTAs
                            umma_commit_2cta(&s.b_empty[sb]);
                        } else {
                            umma_commit(&s.a_empty[sa]);  // arrives when the MMAs above have read the operands
                            umma_commit(&s.b_empty[sb]);
                        }
                    }
                    if (CG == 2) umma_commit_2cta(&s.t_full[acc]);
                    else umma_commit(&s.t_full[acc]);  // accumulator complete
                }
            }
        }
    } else if (CG == 2 && !leader && (warp == 2 || warp == 3)) {
        // ===================== relays (second CTA of a pair): local "full" -> the leader's barrier =====================
        // The leader's MMA thread waits on ITS a_full / b_full only; this CTA's generators and TMA complete on the local
        // copies, and one thread per ring forwards each completion with a single remote arrive.
        if (lane == 0) {
            uint32_t g = 0;
            for (int u = unit0; u < n_units; u += unit_stride) {
                const Unit un = unit(tile_of(u) < n_tiles ? tile_of(u) : n_tiles - 1);
                for (int ct = 0; ct < un.nct; ct++) {
                    const int nch = chunks_of(un, ct);
                    for (int kc = 0; kc < nch; kc++, g++) {
                        if (warp == 2) {
                            const int sa = g % AST;
                            wait(&s.a_full[sa], (g / AST) & 1);
                            mbar_arrive_remote(&s.a_full[sa], 0);
                        } else {
                            const int sb = g % BST;
                            wait(&s.b_full[sb], (g / BST) & 1);
                            mbar_arrive_remote(&s.b_full[sb], 0);
                        }
                    }
                }
            }
        }
    } else if (warp >= GEN_WARP0 && warp < EPI_WARP0) {
        // ===================== K* generators =====================
        const int gt = tid - GEN_WARP0 * 32;
        // Each thread generates TWO particle rows (row0, row0 + 64) x a quarter of the chunk's k: a training record read from
        // shared memory (a warp-wide broadcast: one wavefront for 16 useful bytes) then serves two K* entries, which halves
        // the generators' record loads -- at d = 8 they were as many shared-memory wavefronts as the tensor core's own operand
        // reads (27 G vs 26 G, profiles/ncu_observe_f16x2_kernel_cfg4_r02.txt), on the pipe that bounds the kernel.
        const int row0 = gt & 63, kq = gt >> 6;
        // exp(-|a - b|^2) = 2^(-|s a - s b|^2), s = sqrt(log2 e): the training coordinates arrive pre-scaled by s
        // (GPMDM.packed_model_tf32) and the particle's are scaled here, so a K* entry is the distance + ONE ex2.approx
        constexpr double SQRT_LOG2E = 1.2011224087864498;
        constexpr int KPT = KCm / 4;                     // k per thread, row and chunk: 4 (tf32) / 8 (fp16)
        // The chunk's KCm training records (KCm x 8 floats, contiguous) are staged in shared memory by the 256 generator
        // threads themselves -- one coalesced 4-byte load per thread, issued a whole chunk ahead -- and read back as
        // warp-wide broadcasts.  (Every thread used to pull its 8-16 records through L1 with 16-byte loads: at d = 8 the
        // 12 % of them that missed L1 were the kernel's largest stall, profiles/ncu_observe_f16x2_kernel_r02.txt.)
        constexpr int CHUNK_FLOATS = KCm * CREC;  // 256 (fp16) / 128 (tf32)
        uint32_t g = 0;
        if (unit0 < n_units && gt < CHUNK_FLOATS) {  // chunk 0 of the first unit
            const Unit un0 = unit(tile_of(unit0) < n_tiles ? tile_of(unit0) : n_tiles - 1);
            s.coords[0][gt] = __ldg(un0.coords + gt);
        }
        named_bar_sync(1, NGEN);
        for (int u = unit0; u < n_units; u += unit_stride) {
            const Unit un = unit(tile_of(u) < n_tiles ? tile_of(u) : n_tiles - 1);
            // the records that follow this unit's last chunk: chunk 0 of the next unit's block
            const float* nxt_c = un.coords;
            if (KIND == 1 && u + unit_stride < n_units) {
                const Unit nu = unit(tile_of(u + unit_stride) < n_tiles ? tile_of(u + unit_stride) : n_tiles - 1);
                nxt_c = nu.coords;
            }
            // particle coordinates, negated and duplicated into both halves of a packed f32x2 register
            float2 nb[2][DL];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int row = row0 + 64 * r;
                long long p;
                if (KIND == 1) p = prm.perm[un.first + (row < un.count ? row : un.count - 1)];
                else {
                    p = (long long)tile_of(u) * TM + row;
                    if (p >= prm.P) p = prm.P - 1;
                }
#pragma unroll
                for (int j = 0; j < DL; j++) {
                    const double xj = prm.x[p * DL + j];
                    const float bj = (float)(xj / prm.ls[j] * SQRT_LOG2E);
                    nb[r][j] = make_float2(-bj, -bj);
                }
            }
            for (int ct = 0; ct < un.nct; ct++) {
                const int nch = chunks_of(un, ct);
                for (int kc = 0; kc < nch; kc++, g++) {
                    const int sa = g % AST;
                    // the next chunk of this thread's walk (every column tile starts again at k = 0)
                    const bool last = kc + 1 == nch && ct + 1 == un.nct;
                    const long long noff = (long long)(kc + 1 < nch ? kc + 1 : 0) * CHUNK_FLOATS + gt;
                    float cnext = 0.f;
                    if (gt < CHUNK_FLOATS) cnext = __ldg((last ? nxt_c : un.coords) + noff);
                    float kv[2][KPT];
                    // records are stored per PAIR of training rows as [j][2] (a_k[j], a_k+1[j]): one 64-bit element feeds
                    // the packed fp32x2 pipe (sm_100 FADD2 / FFMA2), two K* entries per instruction
                    const float2* rec = reinterpret_cast<const float2*>(s.coords[g & 1]) + (kq * KPT) / 2 * CREC;
#pragma unroll
                    for (int kk = 0; kk < KPT; kk += 2) {
                        float2 a[CREC];
#pragma unroll
                        for (int q = 0; q < (DL + 1) / 2; q++) {
                            const float4 r4 = *reinterpret_cast<const float4*>(rec + (kk / 2) * CREC + 2 * q);
                            a[2 * q] = make_float2(r4.x, r4.y);
                            a[2 * q + 1] = make_float2(r4.z, r4.w);
                        }
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            float2 dist = make_float2(0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < DL; j++) {
                                const float2 tdiff = __fadd2_rn(a[j], nb[r][j]);
                                dist = __ffma2_rn(tdiff, tdiff, dist);
                            }
                            kv[r][kk] = ex2_approx(-dist.x);
                            kv[r][kk + 1] = ex2_approx(-dist.y);
                        }
                    }
                    if (gt < CHUNK_FLOATS) s.coords[(g + 1) & 1][gt] = cnext;
                    named_bar_sync(1, NGEN);  // next chunk's records visible; nobody still reads the buffer written next time
                    wait(&s.a_empty[sa], ((g / AST) & 1) ^ 1);
                    if (F16) {
                        // a = hi + 2^-11 lo: 8 consecutive k of one row = one 16-byte core-matrix row per piece
                        static_assert(!F16 || KPT == 8, "one core-matrix row per thread, row and piece");
                        __half* ah = reinterpret_cast<__half*>(&s.A[sa][0][0]);
                        __half* al = reinterpret_cast<__half*>(&s.A[sa][1][0]);
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            // packed conversions (two values per F2FP): scalar fp32 <-> fp16 conversions run at a
                            // fraction of the rate
                            __half2 hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const float k0 = kv[r][2 * i], k1 = kv[r][2 * i + 1];
                                hi[i] = __floats2half2_rn(k0, k1);
                                const float2 hb = __half22float2(hi[i]);
                                lo[i] = __floats2half2_rn(fmaf(hb.x, -2048.f, k0 * 2048.f), fmaf(hb.y, -2048.f, k1 * 2048.f));
                            }
                            const int idx = tile_index<MODE>(row0 + 64 * r, kq * KPT);
                            *reinterpret_cast<uint4*>(ah + idx) = make_uint4(*reinterpret_cast<uint32_t*>(&hi[0]), *reinterpret_cast<uint32_t*>(&hi[1]),
                                                                             *reinterpret_cast<uint32_t*>(&hi[2]), *reinterpret_cast<uint32_t*>(&hi[3]));
                            *reinterpret_cast<uint4*>(al + idx) = make_uint4(*reinterpret_cast<uint32_t*>(&lo[0]), *reinterpret_cast<uint32_t*>(&lo[1]),
                                                                             *reinterpret_cast<uint32_t*>(&lo[2]), *reinterpret_cast<uint32_t*>(&lo[3]));
                        }
                    } else {
                        static_assert(F16 || KPT == 4, "one core-matrix row per thread, row and piece");
                        float* ah = reinterpret_cast<float*>(&s.A[sa][0][0]);
                        float* al = reinterpret_cast<float*>(&s.A[sa][1][0]);
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            float hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                hi[i] = to_tf32(kv[r][i]);
                                lo[i] = to_tf32(kv[r][i] - hi[i]);
                            }
                            const int idx = tile_index<MODE>(row0 + 64 * r, kq * KPT);
                            *reinterpret_cast<float4*>(ah + idx) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<float4*>(al + idx) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                    mbar_arrive(&s.a_full[sa]);
                }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue =====================
        const int et = tid - EPI_WARP0 * 32;  // particle row; this warp owns TMEM lanes 32 (warp % 4) ..
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t tcount = 0;
        for (int u = unit0; u < n_units; u += unit_stride) {
            const Unit un = unit(tile_of(u) < n_tiles ? tile_of(u) : n_tiles - 1);
            long long p;
            bool valid;
            if (KIND == 1) {
                valid = et < un.count;
                p = prm.perm[un.first + (valid ? et : un.count - 1)];
            } else {
                p = (long long)tile_of(u) * TM + et;
                valid = p < prm.P;
            }
            float q = 0.f;
            double S = 0.0;
            for (int ct = 0; ct < un.nct; ct++, tcount++) {
                const int acc = F16 ? 0 : (tcount & 1);
                const uint32_t use = F16 ? tcount : (tcount >> 1);
                wait(&s.t_full[acc], use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem + lane_base + (uint32_t)acc * TN;
                if (ct < un.nq) {
#pragma unroll 1
                    for (int cb = 0; cb < TN; cb += 32) {
                        float v[32];
                        tmem_ld32(taddr + cb, v);
                        float part = 0.f;
                        if (F16) {
                            float w[32];
                            tmem_ld32(taddr + TN + cb, w);
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                const float uu = fmaf(w[i], 1.0f / 2048.f, v[i]);
                                part = fmaf(uu, uu, part);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; i++) part = fmaf(v[i], v[i], part);
                        }
                        q += part;
                    }
                } else {
#pragma unroll 1
                    for (int cb = 0; cb < TN && cb < prm.dout; cb += 32) {
                        float v[32];
                        tmem_ld32(taddr + cb, v);
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const int col = cb + i;
                            if (col < prm.dout) {
                                if (prm.z) {
                                    const double dz = prm.z[col] - (double)v[i];
                                    S = fma(prm.lam2[col] * dz, dz, S);
                                }
                                if (prm.mu_out && valid) prm.mu_out[p * prm.dout + col] = (double)v[i];
                            }
                        }
                    }
                }
                tc_fence_before();
                if (CG == 2 && !leader) mbar_arrive_remote(&s.t_empty[acc], 0);  // the leader's MMA thread owns the count
                else mbar_arrive(&s.t_empty[acc]);
            }
            if (valid) {
                // KIND 1: the RBF part 1 - |W_c k_rbf|^2 only; the linear-kernel part of the class-block variance is
                // low rank and is added exactly, in fp64, by gpmdm_pf_propagate_meanonly_f64 (which also counts faults)
                const double v = 1.0 - (double)q;
                if (KIND == 0 && prm.status && !(v > 0.0 && v < INFINITY)) atomicAdd(prm.status, 1);
                if (prm.ll) prm.ll[p] = -0.5 * S / v - (double)prm.dout * log(v) + prm.ll_const;
                if (prm.v_out) prm.v_out[p] = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();  // neither CTA leaves (or frees TMEM) while the pair's MMAs may still read its shared memory
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- packing ------------------------------------------------------------------------------------------------
// W [n, n] (fp64, lower triangular, W = U^-T) -> tf32 hi/lo tiles in the UMMA canonical layout.
//   tile (J, kc), kc < 16 (J+1):  B[row = n - 256 J][k - 16 kc] = W[n][k]
__global__ void pack_whitened_kernel(const double* __restrict__ W, long long n, int n_pad, float* __restrict__ out) {
    const int nq = n_pad / TN;
    const long long total_tiles = wtile_offset(nq);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_tiles * B_HALF) return;
    const long long tile = i / B_HALF;
    const int e = (int)(i % B_HALF);
    // invert wtile_offset: largest J with offset(J) <= tile
    int J = (int)((sqrt(1.0 + 8.0 * (double)tile / (TN / KC)) - 1.0) * 0.5);
    while (wtile_offset(J + 1) <= tile) J++;
    while (wtile_offset(J) > tile) J--;
    const int kc = (int)(tile - wtile_offset(J));
    const int row = e / KC, k = e % KC;
    const long long nn = (long long)J * TN + row, kk = (long long)kc * KC + k;
    const double w = (nn < n && kk < n && kk <= nn) ? W[nn * n + kk] : 0.0;
    const float hi = __uint_as_float(__float_as_uint((float)w) & 0xffffe000u);  // exact tf32 (truncation)
    const float lo = (float)(w - (double)hi);
    float* dst = out + tile * (2 * B_HALF);
    const int idx = tile_index(row, k);
    dst[idx] = hi;
    dst[B_HALF + idx] = lo;
}

// alpha [n, dout] (fp64) -> tiles (kc):  B[row = output column j][k - 16 kc] = alpha[k][j]
__global__ void pack_alpha_kernel(const double* __restrict__ alpha, long long n, int n_pad, int dout,
                                  float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)(n_pad / KC) * B_HALF;
    if (i >= total) return;
    const long long kc = i / B_HALF;
    const int e = (int)(i % B_HALF);
    const int row = e / KC, k = e % KC;
    const long long kk = kc * KC + k;
    const double a = (row < dout && kk < n) ? alpha[kk * dout + row] : 0.0;
    const float hi = __uint_as_float(__float_as_uint((float)a) & 0xffffe000u);
    const float lo = (float)(a - (double)hi);
    float* dst = out + kc * (2 * B_HALF);
    const int idx = tile_index(row, k);
    dst[idx] = hi;
    dst[B_HALF + idx] = lo;
}

// W [n, n] (fp64, lower triangular) -> fp16 hi / scaled-lo tiles (MODE_F16X2): w = hi + 2^-11 lo.
//   tile (J, kc), kc < 8 (J+1):  B[row = n - 256 J][k - 32 kc] = W[n][k]
__global__ void pack_whitened_f16_kernel(const double* __restrict__ W, long long n, int n_pad, __half* __restrict__ out) {
    constexpr int KCh = Geo<MODE_F16X2>::KC;
    constexpr int TILE_ELEMS = B_HALF_BYTES / 2;  // halves in one piece of a packed tile: 256 x 32
    const int nq = n_pad / TN;
    const long long total_tiles = wtile_offset_kc(nq, KCh);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_tiles * TILE_ELEMS) return;
    const long long tile = i / TILE_ELEMS;
    const int e = (int)(i % TILE_ELEMS);
    int J = (int)((sqrt(1.0 + 8.0 * (double)tile / (TN / KCh)) - 1.0) * 0.5);
    while (wtile_offset_kc(J + 1, KCh) <= tile) J++;
    while (wtile_offset_kc(J, KCh) > tile) J--;
    const int kc = (int)(tile - wtile_offset_kc(J, KCh));
    const int row = e / KCh, k = e % KCh;
    const long long nn = (long long)J * TN + row, kk = (long long)kc * KCh + k;
    const double w = (nn < n && kk < n && kk <= nn) ? W[nn * n + kk] : 0.0;
    const __half hi = __double2half(w);
    const __half lo = __double2half((w - (double)__half2float(hi)) * 2048.0);
    __half* dst = out + tile * (2 * TILE_ELEMS);
    const int idx = tile_index<MODE_F16X2>(row, k);
    dst[idx] = hi;
    dst[TILE_ELEMS + idx] = lo;
}

// CTA pairs (cta_group::2) only on request (GPMDM_TC_CLUSTER=1).  Measured on B200 (profiles/tensor_core_variants_r02.md):
// bit-identical results, but never faster than independent CTAs -- 3 % slower (fp16 split) to 22 % slower (tf32) at
// N = 20 k, level / 11 % slower at N = 50 k: what the pair saves in W-tile bytes it loses on the cross-CTA hand-offs
// (generators -> relay -> leader barrier, multicast commits back), even with 8- / 5-deep rings.
static bool use_cluster(long long tiles) {
    const char* e = getenv("GPMDM_TC_CLUSTER");  // read per call: tests compare the two variants in one process
    return e && e[0] == '1' && tiles >= 4;
}

template <int DL, int MODE, int CG, int KIND = 0>
static int launch_cg(const Params& prm, int grid, cudaStream_t st) {
    // the opt-in to > 48 KB of dynamic shared memory is per function AND per device
    static bool configured[64] = {};  // per instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    auto kern = observe_tf32_kernel<DL, MODE, CG, KIND>;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<CG>));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(observe_tf32): %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured[dev] = true;
    }
    if (CG == 1) {
        kern<<<grid, NTHREADS, sizeof(SmemT<CG>), st>>>(prm);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = sizeof(SmemT<CG>);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, prm);
        if (e != cudaSuccess) {
            set_error("cudaLaunchKernelEx(observe_tf32, cluster 2): %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    return check_launch("observe_tf32_kernel");
}

// grid: one CTA per 128-particle tile up to one per SM; in pairs when the cluster variant is used
template <int DL, int MODE = MODE_TF32X3>
static int launch(const Params& prm, long long tiles, int sms, cudaStream_t st) {
    if (use_cluster(tiles)) {
        const long long units = (tiles + 1) / 2, clusters = units < sms / 2 ? units : sms / 2;
        return launch_cg<DL, MODE, 2>(prm, (int)(2 * clusters), st);
    }
    return launch_cg<DL, MODE, 1>(prm, (int)(tiles < sms ? tiles : sms), st);
}

// ---- tf32 tensor-core peak probe (the denominator of the tf32 roofline in bench.py) --------------------------------
// One CTA per SM; one thread issues `iters` x 2 tcgen05.mma (M = 128, N = 256, K = 8, kind::tf32) back to back on a
// resident pair of operand tiles holding pseudo-random values (realistic datapath toggling, i.e. realistic power),
// alternating between the two TMEM accumulators exactly as observe_tf32_kernel does.  No loads, no epilogue: this is
// the rate the tensor pipe sustains when nothing else limits it.
struct ProbeSmem {
    float A[A_HALF];
    float B[B_HALF];
    uint64_t done;
    uint32_t tmem_base;
};

template <bool F16>
__global__ void __launch_bounds__(128, 1) tf32_probe_kernel(int iters, float* __restrict__ sink) {
    __shared__ __align__(1024) ProbeSmem s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t h = 0x9E3779B9u * (uint32_t)(tid + 1);
    for (int i = tid; i < A_HALF + B_HALF; i += 128) {
        h = h * 1664525u + 1013904223u;
        const float u = (float)(h >> 8) * (1.0f / 16777216.0f);  // [0, 1)
        // the same bytes are read as 16 x tf32 or 32 x fp16 per row; for fp16 pack two modest values per word
        const float a = i < A_HALF ? u : (u - 0.5f) * 1e-3f;
        float word = to_tf32(a);
        if (F16) {
            const __half2 h2 = __floats2half2_rn(a, a * 0.75f);
            word = *reinterpret_cast<const float*>(&h2);
        }
        if (i < A_HALF) s.A[i] = word;
        else s.B[i - A_HALF] = word;
    }
    if (tid == 0) {
        mbar_init(&s.done, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    if (tid == 0) {
        for (int it = 0; it < iters; it++) {
            const uint32_t d_tmem = tmem + (uint32_t)(it & 1) * TN;
#pragma unroll
            for (int k8 = 0; k8 < KC / 8; k8++) {
                const uint32_t koff = k8 * 2 * LBO;
                if (F16)
                    umma_f16(d_tmem, smem_desc(reinterpret_cast<const char*>(s.A) + koff),
                             smem_desc(reinterpret_cast<const char*>(s.B) + koff), (it > 1) || k8);
                else
                    umma_tf32(d_tmem, smem_desc(reinterpret_cast<const char*>(s.A) + koff),
                              smem_desc(reinterpret_cast<const char*>(s.B) + koff), (it > 1) || k8);
            }
        }
        umma_commit(&s.done);
        mbar_wait(&s.done, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        float v[32];
        tmem_ld32(tmem, v);  // keep the accumulator observable
        if (v[0] == 123.456f && sink) sink[0] = v[1];
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace tf32
}  // namespace gpmdm

using namespace gpmdm;

extern "C" int64_t gpmdm_tf32_wtiles_bytes(int64_t n_pad) {
    return tf32::wtile_offset((int)(n_pad / tf32::TN)) * (2 * tf32::B_HALF) * 4;
}
extern "C" int64_t gpmdm_tf32_atiles_bytes(int64_t n_pad) { return (n_pad / tf32::KC) * (2 * tf32::B_HALF) * 4; }

extern "C" int gpmdm_pack_whitened_tf32(const double* W, int64_t n, int64_t n_pad, float* wtiles, void* stream) {
    GPMDM_REQUIRE(W && wtiles, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && n_pad >= n && n_pad % GPMDM_TILE_N == 0, GPMDM_E_INVALID, "bad sizes");
    const long long total = tf32::wtile_offset((int)(n_pad / tf32::TN)) * tf32::B_HALF;
    tf32::pack_whitened_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, n, (int)n_pad, wtiles);
    return check_launch("pack_whitened_kernel");
}

extern "C" int gpmdm_pack_alpha_tf32(const double* alpha, int64_t n, int64_t n_pad, int32_t dout, float* atiles,
                                     void* stream) {
    GPMDM_REQUIRE(alpha && atiles, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && n_pad >= n && n_pad % GPMDM_TILE_N == 0 && dout >= 1 && dout <= tf32::TN, GPMDM_E_INVALID,
                  "bad sizes (dout must be <= %d)", tf32::TN);
    const long long total = (n_pad / tf32::KC) * tf32::B_HALF;
    tf32::pack_alpha_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(alpha, n, (int)n_pad, dout,
                                                                                              atiles);
    return check_launch("pack_alpha_kernel");
}

extern "C" int gpmdm_pf_observe_tf32(const gpmdm_gp_model_tf32* m, const double* x, int64_t P, const double* z,
                                     double ll_const, double* ll, double* mu_out, double* v_out, int32_t* tile_counter,
                                     void* stream) {
    GPMDM_REQUIRE(m && m->coords && m->wtiles && m->atiles && m->lengthscales && m->lambdas, GPMDM_E_INVALID,
                  "null model field");
    GPMDM_REQUIRE(m->d >= 1 && m->d <= GPMDM_MAX_LATENT, GPMDM_E_UNSUPPORTED, "latent dimension %d outside [1, %d]", m->d,
                  GPMDM_MAX_LATENT);
    GPMDM_REQUIRE(m->dout >= 1 && m->dout <= tf32::TN, GPMDM_E_UNSUPPORTED, "dout %d outside [1, %d]", m->dout, tf32::TN);
    GPMDM_REQUIRE(m->n_pad > 0 && m->n_pad % GPMDM_TILE_N == 0, GPMDM_E_INVALID, "n_pad must be a multiple of %d",
                  GPMDM_TILE_N);
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P out of range");
    if (P == 0) return 0;
    GPMDM_REQUIRE(x, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(ll == nullptr || z != nullptr, GPMDM_E_INVALID, "ll requested without z");
    tf32::Params prm{};
    prm.coords = m->coords;
    prm.wtiles = m->wtiles;
    prm.atiles = m->atiles;
    prm.n_pad = (int)m->n_pad;
    prm.d = m->d;
    prm.dout = m->dout;
    prm.ls = m->lengthscales;
    prm.lam2 = m->lambdas;
    prm.x = x;
    prm.P = P;
    prm.z = z;
    prm.ll_const = ll_const;
    prm.ll = ll;
    prm.mu_out = mu_out;
    prm.v_out = v_out;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (P + tf32::TM - 1) / tf32::TM;
    cudaStream_t st = (cudaStream_t)stream;
    prm.status = tile_counter ? tile_counter + 3 : nullptr;
    if (tile_counter && tiles >= 2ll * sms) {
        cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int32_t), st);
        GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        prm.round_counter = tile_counter;
    }
    switch (m->d) {
        case 1: return tf32::launch<1>(prm, tiles, sms, st);
        case 2: return tf32::launch<2>(prm, tiles, sms, st);
        case 3: return tf32::launch<3>(prm, tiles, sms, st);
        case 4: return tf32::launch<4>(prm, tiles, sms, st);
        case 5: return tf32::launch<5>(prm, tiles, sms, st);
        case 6: return tf32::launch<6>(prm, tiles, sms, st);
        case 7: return tf32::launch<7>(prm, tiles, sms, st);
        case 8: return tf32::launch<8>(prm, tiles, sms, st);
    }
    return GPMDM_E_UNSUPPORTED;
}

template <bool F16>
static int probe_impl(int32_t iters, double* tflops_host) {
    GPMDM_REQUIRE(tflops_host && iters > 0, GPMDM_E_INVALID, "bad argument");
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    GPMDM_REQUIRE(e == cudaSuccess, GPMDM_E_NODEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    tf32::tf32_probe_kernel<F16><<<sms, 128>>>(iters / 4 + 1, nullptr);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(t0);
        tf32::tf32_probe_kernel<F16><<<sms, 128>>>(iters, nullptr);
        cudaEventRecord(t1);
        cudaEventSynchronize(t1);
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    int rc = check_launch("tf32_probe_kernel");
    if (rc) return rc;
    // per iteration: two MMAs of M = 128, N = 256 and K = 8 (tf32) / 16 (fp16)
    const double flops = 2.0 * tf32::TM * tf32::TN * (F16 ? 32.0 : 16.0) * (double)iters * sms;
    *tflops_host = flops / (best * 1e-3) / 1e12;
    return 0;
}

extern "C" int gpmdm_probe_tf32_tflops(int32_t iters, double* tflops_host) { return probe_impl<false>(iters, tflops_host); }
extern "C" int gpmdm_probe_f16_tflops(int32_t iters, double* tflops_host) { return probe_impl<true>(iters, tflops_host); }

extern "C" int64_t gpmdm_f16_wtiles_bytes(int64_t n_pad) {
    return tf32::wtile_offset_kc((int)(n_pad / tf32::TN), tf32::Geo<tf32::MODE_F16X2>::KC) * (2 * tf32::B_HALF_BYTES);
}

extern "C" int gpmdm_pack_whitened_f16x2(const double* W, int64_t n, int64_t n_pad, void* wtiles, void* stream) {
    GPMDM_REQUIRE(W && wtiles, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(n > 0 && n_pad >= n && n_pad % GPMDM_TILE_N == 0, GPMDM_E_INVALID, "bad sizes");
    const long long total = gpmdm_f16_wtiles_bytes(n_pad) / 4;  // one thread per (hi, lo) pair of halves
    tf32::pack_whitened_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        W, n, (int)n_pad, static_cast<__half*>(wtiles));
    return check_launch("pack_whitened_f16_kernel");
}

extern "C" int gpmdm_pf_observe_f16x2(const gpmdm_gp_model_tf32* m, const double* x, int64_t P, double* v_out,
                                      int32_t* tile_counter, void* stream) {
    GPMDM_REQUIRE(m && m->coords && m->wtiles && m->lengthscales, GPMDM_E_INVALID, "null model field");
    GPMDM_REQUIRE(m->d >= 1 && m->d <= GPMDM_MAX_LATENT, GPMDM_E_UNSUPPORTED, "latent dimension %d outside [1, %d]", m->d,
                  GPMDM_MAX_LATENT);
    GPMDM_REQUIRE(m->n_pad > 0 && m->n_pad % GPMDM_TILE_N == 0, GPMDM_E_INVALID, "n_pad must be a multiple of %d",
                  GPMDM_TILE_N);
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P out of range");
    if (P == 0) return 0;
    GPMDM_REQUIRE(x && v_out, GPMDM_E_INVALID, "null argument");
    tf32::Params prm{};
    prm.coords = m->coords;
    prm.wtiles = m->wtiles;
    prm.n_pad = (int)m->n_pad;
    prm.d = m->d;
    prm.dout = m->dout;
    prm.ls = m->lengthscales;
    prm.lam2 = m->lambdas;
    prm.x = x;
    prm.P = P;
    prm.v_out = v_out;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (P + tf32::TM - 1) / tf32::TM;
    cudaStream_t st = (cudaStream_t)stream;
    prm.status = tile_counter ? tile_counter + 3 : nullptr;
    if (tile_counter && tiles >= 2ll * sms) {
        cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int32_t), st);
        GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        prm.round_counter = tile_counter;
    }
    switch (m->d) {
        case 1: return tf32::launch<1, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 2: return tf32::launch<2, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 3: return tf32::launch<3, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 4: return tf32::launch<4, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 5: return tf32::launch<5, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 6: return tf32::launch<6, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 7: return tf32::launch<7, tf32::MODE_F16X2>(prm, tiles, sms, st);
        case 8: return tf32::launch<8, tf32::MODE_F16X2>(prm, tiles, sms, st);
    }
    return GPMDM_E_UNSUPPORTED;
}

extern "C" int gpmdm_pf_dynvar_tc(const gpmdm_tc_block* blocks, int32_t n_blocks, int32_t d, int32_t mode,
                                  const double* lengthscales, const double* x_prev, const int32_t* perm,
                                  const int32_t* tiles128, const int32_t* n_tiles128, int64_t P, double* u_out,
                                  void* stream) {
    GPMDM_REQUIRE(blocks && lengthscales, GPMDM_E_INVALID, "null model field");
    GPMDM_REQUIRE(n_blocks >= 1 && d >= 1 && d <= GPMDM_MAX_LATENT && (mode == 0 || mode == 1), GPMDM_E_UNSUPPORTED,
                  "bad sizes n_blocks=%d d=%d mode=%d", n_blocks, d, mode);
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P out of range");
    if (P == 0) return 0;
    GPMDM_REQUIRE(x_prev && perm && tiles128 && n_tiles128 && u_out, GPMDM_E_INVALID, "null argument");
    tf32::Params prm{};
    prm.dblocks = blocks;
    prm.d = d;
    prm.ls = lengthscales;
    prm.x = x_prev;
    prm.P = P;
    prm.perm = perm;
    prm.tiles = tiles128;
    prm.n_tiles_dev = n_tiles128;
    prm.v_out = u_out;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long max_tiles = (P + tf32::TM - 1) / tf32::TM + n_blocks;
    const int grid = (int)(max_tiles < sms ? max_tiles : sms);
    cudaStream_t st = (cudaStream_t)stream;
#define GPMDM_DYN_CASE(DL)                                                                          \
    case DL:                                                                                        \
        return mode == 1 ? tf32::launch_cg<DL, tf32::MODE_F16X2, 1, 1>(prm, grid, st)               \
                         : tf32::launch_cg<DL, tf32::MODE_TF32X3, 1, 1>(prm, grid, st);
    switch (d) {
        GPMDM_DYN_CASE(1) GPMDM_DYN_CASE(2) GPMDM_DYN_CASE(3) GPMDM_DYN_CASE(4)
        GPMDM_DYN_CASE(5) GPMDM_DYN_CASE(6) GPMDM_DYN_CASE(7) GPMDM_DYN_CASE(8)
    }
#undef GPMDM_DYN_CASE
    return GPMDM_E_UNSUPPORTED;
}
