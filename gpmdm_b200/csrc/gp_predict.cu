// GP predictive mean / quadratic form for every particle without materialising the P x N
// cross-covariance -- the contraction kernel of the filter step.
//
// Replaces (reference paths): gpmdm/gpmdm.py:923-963 (map_x_to_y), :1032-1068
// (map_x_dynamics_for_class), and the per-particle Python loop of gpmdm/gpmdm_pf.py:188-192 plus
// the draw of :167-168, which are fused here as epilogues.
//
// Shape of the computation, for one 128-particle tile (rows p) of one GP block (N_pad training rows):
//     C[p, n]  = sum_k  K*[p, k] * B[k, n]          B = [ L | alpha ],  K* generated on the fly
//     q[p]     = sum_n  C[p, n] * K*[p, n]           (columns of L;   Hadamard + row-sum epilogue)
//     mean[p,:] = C[p, N_pad:]                        (columns of alpha)
// With tri = 1, L is the lower-triangular packing of K^-1 (k^T K^-1 k == k^T L k), so column tile J
// only needs k >= 128 J: half the flops and half the bytes of the dense form.
//
// Kernel organisation (one persistent CTA per SM, 384 threads = 3 warpgroups, warp-specialised because
// on sm_100a the fp64 MMA runs on the tensor pipe while exp()/DFMA run on the separate fp64 pipe --
// ncu: sm__pipe_tensor_subpipe_dmma vs sm__pipe_fp64 -- so the K* prologue can overlap the MMAs):
//   * warpgroup 0 (4 warps, one per SM sub-partition) is the PRODUCER: it generates the A operand
//     (K* chunk [16 x 128], exp of an FMA chain over the particle/training records, + the linear kernel
//     for the dynamics GP) into a 3-deep shared ring -- the "GEMM prologue from latent coordinates" --
//     and its warp 0 also feeds the B ring: B tiles [16 x 128] of L / alpha arrive through TMA 1-D bulk
//     copies (cp.async.bulk, SASS UBLKCP) into a 6-stage ring, completion counted on mbarriers;
//   * warpgroups 1-2 (8 warps as 4 (rows) x 2 (columns)) are the CONSUMERS: fp64 tensor-core MMAs
//     (mma.sync m8n8k4, the fastest DMMA shape on sm_100a, see profiles/microbench) accumulate a
//     32 x 64 warp tile = 64 accumulators per lane, then run the fused epilogues;
//   * producers hand registers to the consumers with setmaxnreg (56 vs 224 per thread);
//   * rings are synchronised only by mbarriers (full/empty); the CTA-wide barrier is used once per
//     128-particle tile;
//   * shared tiles use a +4 double row padding, which makes both fragment loads bank-conflict free
//     for the m8n8k4 lane layout (address = (lane&3)*132 + lane>>2 (+const) covers 16 distinct 8-byte
//     banks per half warp);
//   * tiles are handed out through an atomic counter, so class-sorted dynamics tiles of different
//     block sizes balance across SMs.  Row results do not depend on tile assignment.
#include <math.h>

#include "common.cuh"
#include "fast_exp.cuh"

namespace gpmdm {

constexpr int TM = GPMDM_TILE;  // particles per tile
constexpr int TN = GPMDM_TILE;  // columns per column tile
constexpr int KC = 16;          // k rows per chunk
constexpr int BSTAGES = 6;      // B ring depth (TMA)
constexpr int ASTAGES = 3;      // A ring depth (generated K*)
constexpr int BAHEAD = BSTAGES - ASTAGES;  // B chunks in flight ahead of the producer's A chunk
constexpr int LDB = TN + 4;
constexpr int LDA = TM + 4;
constexpr int NPROD = 128;                 // producer threads (warpgroup 0)
constexpr int NCONS = 256;                 // consumer threads (warpgroups 1-2)
constexpr int NTHREADS = NPROD + NCONS;
constexpr int PROD_REGS = 72, CONS_REGS = 216;  // 128*72 + 256*216 == 384*168
constexpr int MAXD = GPMDM_MAX_LATENT;
constexpr int REC_MAX = 2 * MAXD + 1;

struct PredictParams {
    const gpmdm_gp_block* blocks;
    int n_blocks, d, dout, alpha_ld, tri;
    const double* ls;
    const double* lin_c2;
    const double* scale;  // obs: lambda_j^2 ; dyn: lambda_k^-2
    const double* x;      // [P, d]
    const int32_t* perm;
    const int32_t* tiles;
    const int32_t* n_tiles;
    long long P;
    int32_t* counter;
    // dynamics epilogue
    const double* eps;
    double* x_new;
    double* mean_out;
    double* var_out;
    // observation epilogue
    const double* z;
    double ll_const;  // 2 sum_j log lambda_j - c32
    double* ll;
    double* mu_out;
    double* v_out;
};

struct __align__(128) Smem {
    double B[BSTAGES][KC][LDB];
    double A[ASTAGES][KC][LDA];
    double Pb[MAXD + 1][TM];  // b_k = x_k / l_k ; row d holds -|b|^2
    double Px[MAXD][TM];      // raw x (linear kernel)
    double prior[TM];
    double vrow[TM];
    double red[2][TM];
    uint64_t b_full[BSTAGES], b_empty[BSTAGES], a_full[ASTAGES], a_empty[ASTAGES];
    int pidx[TM];
    int tile;
};

template <int KIND, int DL>
__device__ __forceinline__ double kstar_from_records(const double (&rec)[REC_MAX], const double (&pb)[MAXD + 1],
                                                     const double (&px)[MAXD], double c2last) {
    constexpr int d = DL;
    double arg = rec[d] + pb[d];
#pragma unroll
    for (int j = 0; j < d; j++) arg = fma(rec[j], pb[j], arg);
    double val = fast_exp(arg);
    if (KIND == 1) {
        double lin = c2last;
#pragma unroll
        for (int j = 0; j < d; j++) lin = fma(rec[d + 1 + j], px[j], lin);
        val += lin;
    }
    return val;
}

template <int KIND, int DL>
__device__ __forceinline__ void load_record(const double* __restrict__ src, double (&rec)[REC_MAX]) {
    constexpr int REC = KIND == 1 ? 2 * DL + 1 : DL + 1;
    if (REC % 2 == 0) {  // 16-byte aligned records: vector loads
#pragma unroll
        for (int q = 0; q < REC; q += 2) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(src + q));
            rec[q] = v.x;
            rec[q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < REC; q++) rec[q] = __ldg(src + q);
    }
}

// K* for NB consecutive training rows and one particle, evaluated stage by stage across the batch so that the
// NB exponentials form independent instruction chains (one producer warp per SM sub-partition has to hide the
// fp64 pipe latency by itself).
template <int KIND, int DL, int NB>
__device__ __forceinline__ void kstar_batch(const double* __restrict__ rbase, const double (&pb)[MAXD + 1],
                                            const double (&px)[MAXD], double c2last, double (&out)[NB]) {
    constexpr int d = DL;
    constexpr int REC = KIND == 1 ? 2 * d + 1 : d + 1;
    double r[NB], lin[NB];
    int n[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        double rec[REC_MAX];
        load_record<KIND, DL>(rbase + b * REC, rec);
        double arg = rec[d] + pb[d];
#pragma unroll
        for (int j = 0; j < d; j++) arg = fma(rec[j], pb[j], arg);
        if (KIND == 1) {
            double l = c2last;
#pragma unroll
            for (int j = 0; j < d; j++) l = fma(rec[d + 1 + j], px[j], l);
            lin[b] = l;
        }
        r[b] = exp_reduce(arg, n[b]);
    }
    constexpr double C[14] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                              1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0,
                              1.0 / 479001600.0, 1.0 / 6227020800.0};
    double p[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) p[b] = C[13];
#pragma unroll
    for (int k = 12; k >= 0; k--)
#pragma unroll
        for (int b = 0; b < NB; b++) p[b] = fma(p[b], r[b], C[k]);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const double v = exp_scale(p[b], n[b]);
        out[b] = KIND == 1 ? v + lin[b] : v;
    }
}

// Chunk schedule of one particle tile: column tile ct covers k-chunks [kbeg(ct), nkc).
struct ChunkCursor {
    int ct, k, nq, nct, nkc, tri;
    __device__ __forceinline__ int kbeg(int t) const { return (tri && t < nq) ? t * (TN / KC) : 0; }
    __device__ __forceinline__ void init(int nq_, int nct_, int nkc_, int tri_) {
        nq = nq_, nct = nct_, nkc = nkc_, tri = tri_, ct = 0, k = 0;
    }
    __device__ __forceinline__ bool done() const { return ct >= nct; }
    __device__ __forceinline__ void next() {
        if (++k == nkc) {
            ct++;
            k = kbeg(ct);
        }
    }
};

template <int KIND, int DL>
__global__ void __launch_bounds__(NTHREADS, 1) gp_predict_kernel(const PredictParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int d = DL;
    constexpr int REC = KIND == 1 ? 2 * d + 1 : d + 1;
    const double c2last = KIND == 1 ? prm.lin_c2[d] : 0.0;
    const bool is_producer = warp < NPROD / 32;

    if (tid == 0) {
        for (int i = 0; i < BSTAGES; i++) {
            mbar_init(&s.b_full[i], 1);           // one arrive.expect_tx + TMA bytes
            mbar_init(&s.b_empty[i], NCONS / 32);  // one arrival per consumer warp
        }
        for (int i = 0; i < ASTAGES; i++) {
            mbar_init(&s.a_full[i], NPROD);        // every producer thread
            mbar_init(&s.a_empty[i], NCONS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int total_tiles = prm.tiles ? *prm.n_tiles : (int)((prm.P + TM - 1) / TM);
    uint32_t g = 0;   // chunks processed by this role so far (ring positions and mbarrier parities)

    // Per-tile geometry, recomputed identically by both roles from s.tile.
#define GPMDM_TILE_GEOMETRY()                                                   \
    const int t = s.tile;                                                       \
    if (t >= total_tiles) break;                                                \
    int blk = 0, first = t * TM, count;                                         \
    if (prm.tiles) {                                                            \
        blk = prm.tiles[4 * t + 0];                                             \
        first = prm.tiles[4 * t + 1];                                           \
        count = prm.tiles[4 * t + 2];                                           \
    } else {                                                                    \
        long long rem = prm.P - (long long)first;                               \
        count = rem < TM ? (int)rem : TM;                                       \
    }                                                                           \
    const gpmdm_gp_block gbk = prm.blocks[blk];                                 \
    const int n_pad = (int)gbk.n_pad;                                           \
    const int nkc = n_pad / KC;                                                 \
    const int nq = n_pad / TN;              /* column tiles of L */             \
    const int nct = nq + prm.alpha_ld / TN; /* + column tiles of alpha */       \
    (void)first; (void)count;                                                   \
    ChunkCursor cur;                                                            \
    cur.init(nq, nct, nkc, prm.tri)

    // The two roles never re-converge after setmaxnreg (ptxas allocates registers per role only then).
    if (is_producer) {
        setmaxnreg_dec<PROD_REGS>();
        uint32_t gb = 0;  // warp 0: B chunks issued so far
        for (;;) {
            if (tid == 0) s.tile = atomicAdd(prm.counter, 1);
            named_bar_sync(2, NPROD);
            GPMDM_TILE_GEOMETRY();
            // ---- particle records ------------------------------------------------------------------
            {
                const int m = tid < count ? tid : count - 1;
                const int p = prm.perm ? prm.perm[first + m] : first + m;
                s.pidx[tid] = tid < count ? p : -1;
                double nb = 0.0, prior = 1.0;
#pragma unroll
                for (int j = 0; j < d; j++) {
                    const double xj = prm.x[(long long)p * d + j];
                    const double bj = xj / prm.ls[j];
                    s.Pb[j][tid] = bj;
                    nb = fma(bj, bj, nb);
                    if (KIND == 1) {
                        s.Px[j][tid] = xj;
                        prior = fma(prm.lin_c2[j] * xj, xj, prior);
                    }
                }
                s.Pb[d][tid] = -nb;
                if (KIND == 1) prior += c2last;
                s.prior[tid] = prior;
            }
            named_bar_sync(0, NTHREADS);  // tile + records visible to the consumers


            // ======================= PRODUCER: B ring (warp 0, TMA) + A ring (K* generation) =================
            ChunkCursor bcur = cur;  // B cursor runs BAHEAD chunks ahead of the A cursor
            auto issue_b = [&]() {
                const int st = (int)(gb % BSTAGES);
                mbar_wait(&s.b_empty[st], ((gb / BSTAGES) & 1) ^ 1);  // first pass: passes immediately
                const double* src;
                int ld;
                if (bcur.ct < nq) {
                    src = gbk.L + (long long)bcur.ct * TN;
                    ld = n_pad;
                } else {
                    src = gbk.alpha + (long long)(bcur.ct - nq) * TN;
                    ld = prm.alpha_ld;
                }
                if (lane == 0) mbar_expect_tx(&s.b_full[st], (uint32_t)(KC * TN * 8));
                __syncwarp();
                if (lane < KC)
                    bulk_g2s(&s.B[st][lane][0], src + (long long)(bcur.k * KC + lane) * ld, TN * 8, &s.b_full[st]);
                bcur.next();
                gb++;
            };
            if (warp == 0)
                for (int i = 0; i < BAHEAD && !bcur.done(); i++) issue_b();

            const int p = tid;  // this thread's particle row
            double pb[MAXD + 1], px[MAXD];
#pragma unroll
            for (int j = 0; j < MAXD; j++) {
                pb[j] = j < d ? s.Pb[j][p] : 0.0;
                px[j] = (KIND == 1 && j < d) ? s.Px[j][p] : 0.0;
            }
            pb[d] = s.Pb[d][p];

            for (; !cur.done(); cur.next(), g++) {
                if (warp == 0 && !bcur.done()) issue_b();
                const int buf = (int)(g % ASTAGES);
                mbar_wait(&s.a_empty[buf], ((g / ASTAGES) & 1) ^ 1);
                const double* rbase = gbk.coords + (long long)cur.k * KC * REC;
                constexpr int NB = 8;
#pragma unroll 1
                for (int k0 = 0; k0 < KC; k0 += NB) {
                    double kv[NB];
                    kstar_batch<KIND, DL, NB>(rbase + k0 * REC, pb, px, c2last, kv);
#pragma unroll
                    for (int b = 0; b < NB; b++) s.A[buf][k0 + b][p] = kv[b];
                }
                mbar_arrive(&s.a_full[buf]);
            }
            named_bar_sync(0, NTHREADS);  // consumers are done with s.Pb / s.pidx / s.tile
        }
        // the "no more tiles" value of s.tile must reach the consumers too
        named_bar_sync(0, NTHREADS);
    } else {
        setmaxnreg_inc<CONS_REGS>();
        for (;;) {
            named_bar_sync(0, NTHREADS);
            GPMDM_TILE_GEOMETRY();
            // ======================= CONSUMERS: DMMA main loop + fused epilogues ==============================
            const int cw = warp - NPROD / 32;
            const int wm = cw & 3, wn = cw >> 2;
            const int r = lane >> 2, c = lane & 3;
            const int ctid = tid - NPROD;
            double qacc[4] = {0.0, 0.0, 0.0, 0.0};
            double sacc[4] = {0.0, 0.0, 0.0, 0.0};

            for (int ct = 0; ct < nct; ct++) {
                double acc[4][8][2];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

                const int kbeg = cur.kbeg(ct);
                for (int k = kbeg; k < nkc; k++, g++) {
                    const int bst = (int)(g % BSTAGES), ast = (int)(g % ASTAGES);
                    mbar_wait(&s.a_full[ast], (g / ASTAGES) & 1);
                    mbar_wait(&s.b_full[bst], (g / BSTAGES) & 1);
#pragma unroll
                    for (int k4 = 0; k4 < KC / 4; k4++) {
                        double a[4], b[8];
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i] = s.A[ast][k4 * 4 + c][wm * 32 + i * 8 + r];
#pragma unroll
                        for (int j = 0; j < 8; j++) b[j] = s.B[bst][k4 * 4 + c][wn * 64 + j * 8 + r];
#pragma unroll
                        for (int i = 0; i < 4; i++)
#pragma unroll
                            for (int j = 0; j < 8; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    }
                    __syncwarp();
                    if (lane == 0) {  // this warp is done reading both ring slots
                        mbar_arrive(&s.a_empty[ast]);
                        mbar_arrive(&s.b_empty[bst]);
                    }
                }

                // ---- epilogues ---------------------------------------------------------------------------
                if (ct < nq) {
                    // q[p] += sum_n C[p,n] * K*[p,n] over this tile's columns (K* regenerated per element)
#pragma unroll
                    for (int j = 0; j < 8; j++) {  // fully unrolled: acc[][][] must stay in registers
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int n = ct * TN + wn * 64 + j * 8 + c * 2 + e;
                            double rec[REC_MAX];
                            load_record<KIND, DL>(gbk.coords + (long long)n * REC, rec);
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const int m = wm * 32 + i * 8 + r;
                                double pb[MAXD + 1], px[MAXD];
#pragma unroll
                                for (int q = 0; q < MAXD; q++) {
                                    pb[q] = q < d ? s.Pb[q][m] : 0.0;
                                    px[q] = (KIND == 1 && q < d) ? s.Px[q][m] : 0.0;
                                }
                                pb[d] = s.Pb[d][m];
                                const double kv = kstar_from_records<KIND, DL>(rec, pb, px, c2last);
                                qacc[i] = fma(acc[i][j][e], kv, qacc[i]);
                            }
                        }
                    }
                    if (ct == nq - 1) {
                        // quadratic form complete: v[p] = prior[p] - q[p]
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            double v = qacc[i];
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            if (c == 0) s.red[wn][wm * 32 + i * 8 + r] = v;
                        }
                        named_bar_sync(1, NCONS);
                        if (ctid < TM) s.vrow[ctid] = s.prior[ctid] - (s.red[0][ctid] + s.red[1][ctid]);
                        named_bar_sync(1, NCONS);
                    }
                } else {
                    const int cbase = (ct - nq) * TN + wn * 64;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int m = wm * 32 + i * 8 + r;
                        const int p = s.pidx[m];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const int col = cbase + j * 8 + c * 2 + e;
                                if (col >= prm.dout) continue;
                                const double mu = acc[i][j][e];
                                if (KIND == 0) {
                                    if (prm.z) {
                                        const double dz = __ldg(prm.z + col) - mu;
                                        sacc[i] = fma(__ldg(prm.scale + col) * dz, dz, sacc[i]);
                                    }
                                    if (prm.mu_out && p >= 0) prm.mu_out[(long long)p * prm.dout + col] = mu;
                                } else if (p >= 0) {
                                    const double var = s.vrow[m] * __ldg(prm.scale + col);
                                    const long long o = (long long)p * prm.dout + col;
                                    if (prm.x_new)  // torch.normal: randn * std + mean, two roundings (no FMA)
                                        prm.x_new[o] = __dadd_rn(__dmul_rn(__ldg(prm.eps + o), sqrt(var)), mu);
                                    if (prm.mean_out) prm.mean_out[o] = mu;
                                    if (prm.var_out) prm.var_out[o] = var;
                                }
                            }
                        }
                    }
                    if (KIND == 0 && ct == nct - 1) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            double v = sacc[i];
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            if (c == 0) s.red[wn][wm * 32 + i * 8 + r] = v;
                        }
                        named_bar_sync(1, NCONS);
                        if (ctid < TM && s.pidx[ctid] >= 0) {
                            const double S = s.red[0][ctid] + s.red[1][ctid];
                            const double v = s.vrow[ctid];
                            const int p = s.pidx[ctid];
                            if (prm.ll) prm.ll[p] = -0.5 * S / v - (double)prm.dout * log(v) + prm.ll_const;
                            if (prm.v_out) prm.v_out[p] = v;
                        }
                    }
                }
            }
            named_bar_sync(0, NTHREADS);  // s.tile, s.pidx, s.Pb ... are rewritten by the next tile
        }
    }
#undef GPMDM_TILE_GEOMETRY
}

// ---- host side -------------------------------------------------------------------------------------
template <int KIND, int DL>
static int launch_instance(const PredictParams& prm, int grid, cudaStream_t st) {
    static bool configured = false;  // per instantiation
    auto kern = gp_predict_kernel<KIND, DL>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gp_predict): %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    kern<<<grid, NTHREADS, sizeof(Smem), st>>>(prm);
    return check_launch("gp_predict_kernel");
}

template <int KIND>
static int dispatch_d(const PredictParams& prm, int grid, cudaStream_t st) {
    switch (prm.d) {
        case 1: return launch_instance<KIND, 1>(prm, grid, st);
        case 2: return launch_instance<KIND, 2>(prm, grid, st);
        case 3: return launch_instance<KIND, 3>(prm, grid, st);
        case 4: return launch_instance<KIND, 4>(prm, grid, st);
        case 5: return launch_instance<KIND, 5>(prm, grid, st);
        case 6: return launch_instance<KIND, 6>(prm, grid, st);
        case 7: return launch_instance<KIND, 7>(prm, grid, st);
        case 8: return launch_instance<KIND, 8>(prm, grid, st);
    }
    set_error("latent dimension %d outside [1, %d]", prm.d, MAXD);
    return GPMDM_E_UNSUPPORTED;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static int validate_model(const gpmdm_gp_model* m, int kind) {
    GPMDM_REQUIRE(m && m->blocks && m->lengthscales && m->lambdas, GPMDM_E_INVALID, "null model field");
    GPMDM_REQUIRE(m->kind == kind, GPMDM_E_INVALID, "model kind %d, expected %d", m->kind, kind);
    GPMDM_REQUIRE(m->d >= 1 && m->d <= MAXD, GPMDM_E_UNSUPPORTED, "latent dimension %d outside [1, %d]", m->d, MAXD);
    GPMDM_REQUIRE(m->dout >= 1 && m->alpha_ld >= m->dout && m->alpha_ld % TN == 0, GPMDM_E_INVALID,
                  "alpha_ld %d must be a multiple of %d and >= dout %d", m->alpha_ld, TN, m->dout);
    GPMDM_REQUIRE(m->n_blocks >= 1, GPMDM_E_INVALID, "model has no blocks");
    GPMDM_REQUIRE(kind == 0 || m->lin_c2, GPMDM_E_INVALID, "dynamics model needs lin_c2");
    return 0;
}

static void fill_common(PredictParams& prm, const gpmdm_gp_model* m) {
    prm = PredictParams{};
    prm.blocks = m->blocks;
    prm.n_blocks = m->n_blocks;
    prm.d = m->d;
    prm.dout = m->dout;
    prm.alpha_ld = m->alpha_ld;
    prm.tri = m->tri;
    prm.ls = m->lengthscales;
    prm.lin_c2 = m->lin_c2;
    prm.scale = m->lambdas;
}

}  // namespace gpmdm

using namespace gpmdm;

extern "C" int gpmdm_pf_propagate_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                                      const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                                      double* x_new, double* mean_out, double* var_out, int32_t* tile_counter,
                                      void* stream) {
    if (int rc = validate_model(dyn, 1)) return rc;
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P = %lld out of range", (long long)P);
    if (P == 0) return 0;
    GPMDM_REQUIRE(x_prev && perm && tiles && n_tiles && tile_counter, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(x_new == nullptr || eps != nullptr, GPMDM_E_INVALID, "x_new requested without eps");
    GPMDM_REQUIRE(dyn->dout == dyn->d, GPMDM_E_INVALID, "dynamics GP must have dout == d");
    cudaStream_t st = (cudaStream_t)stream;
    PredictParams prm;
    fill_common(prm, dyn);
    prm.x = x_prev;
    prm.perm = perm;
    prm.tiles = tiles;
    prm.n_tiles = n_tiles;
    prm.P = P;
    prm.counter = tile_counter;
    prm.eps = eps;
    prm.x_new = x_new;
    prm.mean_out = mean_out;
    prm.var_out = var_out;
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int32_t), st);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const long long max_tiles = (P + TM - 1) / TM + dyn->n_blocks;
    const int grid = (int)(max_tiles < num_sms() ? max_tiles : num_sms());
    return dispatch_d<1>(prm, grid, st);
}

extern "C" int gpmdm_pf_observe_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z,
                                    double ll_const, double* ll, double* mu_out, double* v_out, int32_t* tile_counter,
                                    void* stream) {
    if (int rc = validate_model(obs, 0)) return rc;
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P = %lld out of range", (long long)P);
    if (P == 0) return 0;
    GPMDM_REQUIRE(x && tile_counter, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(ll == nullptr || z != nullptr, GPMDM_E_INVALID, "ll requested without z");
    GPMDM_REQUIRE(obs->n_blocks == 1, GPMDM_E_INVALID, "observation GP has exactly one block");
    cudaStream_t st = (cudaStream_t)stream;
    PredictParams prm;
    fill_common(prm, obs);
    prm.x = x;
    prm.P = P;
    prm.counter = tile_counter;
    prm.z = z;
    prm.ll_const = ll_const;
    prm.ll = ll;
    prm.mu_out = mu_out;
    prm.v_out = v_out;
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int32_t), st);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const long long n_tiles = (P + TM - 1) / TM;
    const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
    return dispatch_d<0>(prm, grid, st);
}
