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
// Kernel organisation (one persistent CTA per SM, 256 threads = 8 warps; measured facts that shaped it, see
// profiles/: on sm_100a DMMA (sm__pipe_tensor_subpipe_dmma) and DFMA/exp (sm__pipe_fp64) contend for ONE
// fp64 math datapath of 64 FMA/clk/SM -- a warp-specialised producer/consumer split only moved the K*
// prologue into `stall_math` behind the consumers' DMMAs -- so the goal is zero idle time on that datapath
// and as few non-MMA fp64 instructions as possible):
//   * each warp owns 16 particle rows x all 128 columns of the column tile (64 accumulators per lane), so
//     every lane GENERATES EXACTLY ITS OWN A FRAGMENTS: the K* values it feeds to mma.sync (rows r, r+8;
//     k = 4 k4 + c) are computed in registers from the particle record (registers) and the training
//     records of the chunk (shared memory) -- the "GEMM prologue from latent coordinates" -- with no A tile
//     in shared memory, no redundancy between warps and no block-wide barrier in the main loop;
//   * the exponentials for chunk g+1 are software-pipelined into the DMMA stream of chunk g (custom
//     fast_exp, csrc/fast_exp.cuh), so the datapath always has an instruction to run;
//   * B tiles [16 x 128] of L / alpha and the 16 training records of the chunk arrive through TMA 1-D bulk
//     copies (cp.async.bulk, SASS UBLKCP) into an 8-stage shared ring; full/empty mbarriers are the only
//     synchronisation between warps (the issuing duty rotates over the warps, 4 chunks ahead);
//   * fp64 tensor-core MMAs are mma.sync m8n8k4, the fastest DMMA shape on sm_100a (profiles/microbench);
//   * the +4 double row padding of the B ring makes the fragment loads bank-conflict free for the m8n8k4
//     lane layout (address = (lane&3)*132 + lane>>2 (+const) covers 16 distinct 8-byte banks per half warp);
//   * epilogues (Hadamard row-sum for the quadratic form, log-likelihood / Gaussian draw) are per warp:
//     a row's 128 columns live in the 4 lanes of a quad, reductions are two shuffles;
//   * tiles are handed out through an atomic counter, so class-sorted dynamics tiles of different
//     block sizes balance across SMs.  Row results do not depend on tile assignment.
#include <math.h>

#include "common.cuh"
#include "fast_exp.cuh"

namespace gpmdm {

constexpr int TM = GPMDM_TILE;  // particles per tile
constexpr int TN = GPMDM_TILE;  // columns per column tile
constexpr int KC = 16;          // k rows per chunk
constexpr int STAGES = 8;       // B / record ring depth (TMA)
constexpr int AHEAD = 4;        // chunks in flight ahead of the consumers
constexpr int LDB = TN + 4;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int MAXD = GPMDM_MAX_LATENT;
constexpr int REC_MAX = 2 * MAXD;
// training record width: a_i[0..d) (+ c_k^2 x_i[k] for the dynamics GP), padded to an even number of doubles
__host__ __device__ constexpr int rec_width(int kind, int d) { return ((kind == 1 ? 2 * d : d) + 1) & ~1; }

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
    double B[STAGES][KC][LDB];
    double R[STAGES][KC * REC_MAX];
    uint64_t full[STAGES], empty[STAGES];
    int tile;
};

// Per-lane particle record: b_k = x_k / l_k, and raw x for the linear kernel.
template <int KIND, int DL>
struct ParticleRec {
    double b[DL];
    double x[KIND == 1 ? DL : 1];
};

template <int KIND, int DL>
__device__ __forceinline__ void load_record(const double* __restrict__ src, double (&rec)[REC_MAX]) {
    constexpr int REC = rec_width(KIND, DL);  // even: 16-byte vector loads (shared or global)
#pragma unroll
    for (int q = 0; q < REC; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + q);
        rec[q] = v.x;
        rec[q + 1] = v.y;
    }
}

// K* of one training record against two particle rows, evaluated stage by stage so that the two
// exponentials are independent chains the scheduler can weave into the surrounding DMMAs.
template <int KIND, int DL>
__device__ __forceinline__ void kstar_pair(const double (&rec)[REC_MAX], const ParticleRec<KIND, DL>& p0,
                                           const ParticleRec<KIND, DL>& p1, double c2last, double& k0, double& k1) {
    constexpr int d = DL;
    // -|a - b|^2 in difference form: the expansion |a|^2 + |b|^2 - 2ab of gpmdm.py:515 loses ~|a|^2 ulps, which the
    // ill-conditioned K^-1 amplifies; the difference form keeps K* at ~1 ulp (2 more fp64 ops per entry)
    double a0, a1;
    {
        const double t0 = rec[0] - p0.b[0], t1 = rec[0] - p1.b[0];
        a0 = -t0 * t0;
        a1 = -t1 * t1;
    }
#pragma unroll
    for (int j = 1; j < d; j++) {
        const double t0 = rec[j] - p0.b[j], t1 = rec[j] - p1.b[j];
        a0 = fma(-t0, t0, a0);
        a1 = fma(-t1, t1, a1);
    }
    int n0, n1;
    const double r0 = exp_reduce(a0, n0), r1 = exp_reduce(a1, n1);
    double e0, e1;
    exp_poly2(r0, r1, e0, e1);
    k0 = exp_scale(e0, n0);
    k1 = exp_scale(e1, n1);
    if (KIND == 1) {
        double l0 = c2last, l1 = c2last;
#pragma unroll
        for (int j = 0; j < d; j++) {
            l0 = fma(rec[d + j], p0.x[j], l0);
            l1 = fma(rec[d + j], p1.x[j], l1);
        }
        k0 += l0;
        k1 += l1;
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
    const int r = lane >> 2, c = lane & 3;
    constexpr int d = DL;
    constexpr int REC = rec_width(KIND, d);
    const double c2last = KIND == 1 ? prm.lin_c2[d] : 0.0;

    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&s.full[i], 1);        // one arrive.expect_tx + TMA bytes
            mbar_init(&s.empty[i], NWARPS);  // one arrival per warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int total_tiles = prm.tiles ? *prm.n_tiles : (int)((prm.P + TM - 1) / TM);
    uint32_t g = 0;  // chunks consumed so far by this CTA (ring position and mbarrier parity)

    for (;;) {
        if (tid == 0) s.tile = atomicAdd(prm.counter, 1);
        __syncthreads();
        const int t = s.tile;
        __syncthreads();
        if (t >= total_tiles) break;
        int blk = 0, first = t * TM, count;
        if (prm.tiles) {
            blk = prm.tiles[4 * t + 0];
            first = prm.tiles[4 * t + 1];
            count = prm.tiles[4 * t + 2];
        } else {
            long long rem = prm.P - (long long)first;
            count = rem < TM ? (int)rem : TM;
        }
        const gpmdm_gp_block gbk = prm.blocks[blk];
        const int n_pad = (int)gbk.n_pad;
        const int nkc = n_pad / KC;
        const int nq = n_pad / TN;               // column tiles of L
        const int nct = nq + prm.alpha_ld / TN;  // + column tiles of alpha

        // ---- this lane's two particle rows ------------------------------------------------------------
        ParticleRec<KIND, DL> pr[2];
        int pidx[2];
        double prior[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int row = warp * 16 + i * 8 + r;
            const int m = row < count ? row : count - 1;
            const int p = prm.perm ? prm.perm[first + m] : first + m;
            pidx[i] = row < count ? p : -1;
            double pri = 1.0;
#pragma unroll
            for (int j = 0; j < d; j++) {
                const double xj = prm.x[(long long)p * d + j];
                pr[i].b[j] = xj / prm.ls[j];
                if (KIND == 1) {
                    pr[i].x[j] = xj;
                    pri = fma(prm.lin_c2[j] * xj, xj, pri);
                }
            }
            if (KIND == 1) pri += c2last;
            prior[i] = pri;
        }

        // ---- TMA issue: all warps advance the same cursor, the duty warp issues ------------------------------
        ChunkCursor bcur;
        bcur.init(nq, nct, nkc, prm.tri);
        uint32_t gb = g;  // ring position of the next chunk to issue
        auto issue_b = [&]() {
            const int st = (int)(gb % STAGES);
            mbar_wait(&s.empty[st], ((gb / STAGES) & 1) ^ 1);  // first fill of a stage passes immediately
            const double* src;
            int ld;
            if (bcur.ct < nq) {
                src = gbk.L + (long long)bcur.ct * TN;
                ld = n_pad;
            } else {
                src = gbk.alpha + (long long)(bcur.ct - nq) * TN;
                ld = prm.alpha_ld;
            }
            if (lane == 0) mbar_expect_tx(&s.full[st], (uint32_t)(KC * TN * 8 + KC * REC * 8));
            __syncwarp();
            if (lane < KC)
                bulk_g2s(&s.B[st][lane][0], src + (long long)(bcur.k * KC + lane) * ld, TN * 8, &s.full[st]);
            else if (lane == KC)
                bulk_g2s(&s.R[st][0], gbk.coords + (long long)bcur.k * KC * REC, (uint32_t)(KC * REC * 8), &s.full[st]);
        };
        for (int i = 0; i < AHEAD && !bcur.done(); i++) {
            if (warp == 0) issue_b();
            bcur.next();
            gb++;
        }

        // ---- A fragments of the first chunk -------------------------------------------------------------
        double a[KC / 4][2];  // a[k4][i] = K*[row 16 w + 8 i + r][k = 4 k4 + c] of the current chunk
        mbar_wait(&s.full[g % STAGES], (g / STAGES) & 1);
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; k4++) {
            double rec[REC_MAX];
            load_record<KIND, DL>(&s.R[g % STAGES][(k4 * 4 + c) * REC], rec);
            kstar_pair<KIND, DL>(rec, pr[0], pr[1], c2last, a[k4][0], a[k4][1]);
        }

        ChunkCursor cur;
        cur.init(nq, nct, nkc, prm.tri);
        double qacc[2] = {0.0, 0.0};
        double sacc[2] = {0.0, 0.0};
        double vrow[2] = {0.0, 0.0};

        for (int ct = 0; ct < nct; ct++) {
            double acc[2][16][2];
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 16; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

            const int kbeg = cur.kbeg(ct);
            for (int k = kbeg; k < nkc; k++, g++) {
                const int st = (int)(g % STAGES);
                // keep the ring AHEAD chunks full; the duty rotates so that no warp is always the one waiting
                if (!bcur.done()) {
                    if (warp == (int)(g % NWARPS)) issue_b();
                    bcur.next();
                    gb++;
                }
                // the next chunk (of this particle tile) provides the records for the next A fragments
                const bool has_next = !(ct == nct - 1 && k == nkc - 1);
                const int stn = (int)((g + 1) % STAGES);
                if (has_next) mbar_wait(&s.full[stn], ((g + 1) / STAGES) & 1);
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; k4++) {
                    double b[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) b[j] = s.B[st][k4 * 4 + c][j * 8 + r];
                    const double a0 = a[k4][0], a1 = a[k4][1];
                    if (has_next) {  // next chunk's A fragments for this k4, woven into the MMAs below
                        double rec[REC_MAX];
                        load_record<KIND, DL>(&s.R[stn][(k4 * 4 + c) * REC], rec);
                        kstar_pair<KIND, DL>(rec, pr[0], pr[1], c2last, a[k4][0], a[k4][1]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        dmma_m8n8k4(acc[0][j][0], acc[0][j][1], a0, b[j]);
                        dmma_m8n8k4(acc[1][j][0], acc[1][j][1], a1, b[j]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.empty[st]);  // this warp is done with the ring slot
            }

            // ---- epilogues (per warp; a row's columns live in the 4 lanes of a quad) -------------------------
            if (ct < nq) {
                // q[p] += sum_n C[p,n] * K*[p,n] over this tile's columns (K* regenerated per element)
#pragma unroll
                for (int j = 0; j < 16; j++) {  // fully unrolled: acc[][][] must stay in registers
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int n = ct * TN + j * 8 + c * 2 + e;
                        double rec[REC_MAX];
                        load_record<KIND, DL>(gbk.coords + (long long)n * REC, rec);
                        double k0, k1;
                        kstar_pair<KIND, DL>(rec, pr[0], pr[1], c2last, k0, k1);
                        qacc[0] = fma(acc[0][j][e], k0, qacc[0]);
                        qacc[1] = fma(acc[1][j][e], k1, qacc[1]);
                    }
                }
                if (ct == nq - 1) {
                    // quadratic form complete: v[p] = prior[p] - q[p]
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        double v = qacc[i];
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        vrow[i] = prior[i] - v;
                    }
                }
            } else {
                const int cbase = (ct - nq) * TN;
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int p = pidx[i];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
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
                                const double var = vrow[i] * __ldg(prm.scale + col);
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
                    for (int i = 0; i < 2; i++) {
                        double S = sacc[i];
                        S += __shfl_xor_sync(0xffffffffu, S, 1);
                        S += __shfl_xor_sync(0xffffffffu, S, 2);
                        if (c == 0 && pidx[i] >= 0) {
                            const double v = vrow[i];
                            if (prm.ll) prm.ll[pidx[i]] = -0.5 * S / v - (double)prm.dout * log(v) + prm.ll_const;
                            if (prm.v_out) prm.v_out[pidx[i]] = v;
                        }
                    }
                }
            }
        }
    }
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
