// GP predictive mean / quadratic form for every particle without materialising the P x N
// cross-covariance -- the contraction kernel of the filter step.
//
// Replaces (reference paths): gpmdm/gpmdm.py:923-963 (map_x_to_y), :1032-1068
// (map_x_dynamics_for_class), and the per-particle Python loop of gpmdm/gpmdm_pf.py:188-192 plus
// the draw of :167-168, which are fused here as epilogues.
//
// Shape of the computation, for one 64-particle tile (rows p) of one GP block (N_pad training rows):
//     C[p, n]  = sum_k  K*[p, k] * B[k, n]          B = [ L | alpha ],  K* generated on the fly
//     q[p]     = sum_n  C[p, n] * K*[p, n]           (columns of L;   Hadamard + row-sum epilogue)
//     mean[p,:] = C[p, N_pad:]                        (columns of alpha)
// With tri = 1, L is the lower-triangular packing of K^-1 (k^T K^-1 k == k^T L k), so column tile J
// only needs k >= 256 J: half the flops and half the bytes of the dense form (the rows above are not even stored).
//
// Kernel organisation (one persistent CTA per SM, 256 threads = 8 warps; measured facts that shaped it, see
// profiles/: on sm_100a DMMA (sm__pipe_tensor_subpipe_dmma) and DFMA/exp (sm__pipe_fp64) contend for ONE
// fp64 math datapath of 64 FMA/clk/SM -- a warp-specialised producer/consumer split only moved the K*
// prologue into `stall_math` behind the consumers' DMMAs -- so the goal is zero idle time on that datapath
// and as few non-MMA fp64 instructions as possible):
//   * the CTA tile is 64 particles x 256 columns; each warp owns 8 particle rows x all 256 columns of the
//     column tile (64 accumulators per lane) -- a wide tile, because every generated K* entry then feeds 256
//     column MMAs -- and every lane GENERATES EXACTLY ITS OWN A FRAGMENTS: the K* values it feeds to mma.sync
//     (row r; k = 4 k4 + c) are computed in registers from the particle record (registers) and the training
//     records of the chunk (shared memory) -- the "GEMM prologue from latent coordinates" -- with no A tile
//     in shared memory, no redundancy between warps and no block-wide barrier in the main loop;
//   * CACHE instantiation (observation GP, fused mode): those fragments are evaluated once per particle tile into a
//     per-CTA global scratch in consumption order and re-read per column tile (see the comment on the kernel); in the
//     other instantiations the exponentials for chunk g+1 are software-pipelined into the DMMA stream of chunk g
//     (custom fast_exp, csrc/fast_exp.cuh), so the datapath always has an instruction to run;
//   * B tiles [16 x 256] of L / alpha arrive as ONE TMA 1-D bulk copy per chunk (cp.async.bulk, SASS UBLKCP; the
//     factors are packed as padded column panels, include/gpmdm_b200.h) -- plus one for the chunk's 16 training
//     records when K* is evaluated on the fly -- into a 6-stage shared ring; full/empty mbarriers are the only
//     synchronisation between warps (the issuing duty rotates over the warps, 3 chunks ahead);
//   * fp64 tensor-core MMAs are mma.sync m8n8k4, the fastest DMMA shape on sm_100a (profiles/microbench);
//   * the +4 double row padding of the B ring makes the fragment loads bank-conflict free for the m8n8k4
//     lane layout (address = (lane&3)*260 + lane>>2 (+const) covers 16 distinct 8-byte banks per half warp);
//   * epilogues (Hadamard row-sum for the quadratic form, log-likelihood / Gaussian draw) are per warp:
//     a row's 256 columns live in the 4 lanes of a quad, reductions are two shuffles;
//   * tiles are handed out through an atomic counter, so class-sorted dynamics tiles of different
//     block sizes balance across SMs.  Row results do not depend on tile assignment.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "fast_exp.cuh"

namespace gpmdm {

constexpr int TM = GPMDM_TILE_P;  // particles per tile (8 warps x 8 rows)
constexpr int TN = GPMDM_TILE_N;  // columns per column tile
constexpr int KC = 16;            // k rows per chunk
constexpr int STAGES = 6;         // B / record ring depth (TMA)
#ifndef GPMDM_AHEAD
#define GPMDM_AHEAD 3             /* tuning experiments: GPMDM_NVCC_EXTRA=-DGPMDM_AHEAD=.. python -m gpmdm_b200.build --force */
#endif
constexpr int AHEAD = GPMDM_AHEAD;  // chunks in flight ahead of the consumers
constexpr int LDB = GPMDM_PANEL_LD;  // = TN + 4
static_assert(LDB == TN + 4, "panel pitch");
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
    int32_t* counter;  // [0] tile hand-out, [1] round synchronisation
    int32_t* status;   // [0] dynamics, [1] observation: particles whose predictive variance was not a positive finite number
    int round_sync;    // re-align the CTAs after every particle tile (uniform tiles, several rounds)
    // low-latency (split) mode: one work item per (particle tile, column tile); partial results go to the workspace
    // The k range of every column tile is cut further into `nseg` segments of `seg_chunks` chunks, so that a handful of
    // particle tiles still yields several work items per SM.
    int split, max_nct, nseg, seg_chunks;
    double* qpart;  // [max_nq * nseg][P] per-(column tile, segment) contributions to k^T L k
    double* mu_ws;  // [nseg][P][dout] per-segment contributions to the means
    // K* cache (CACHE instantiation): per-CTA scratch of n_pad x 64 doubles in A-fragment order
    double* kcache;
    long long kcache_stride;  // doubles per CTA
    // dynamics epilogue
    const double* eps;
    double* x_new;
    double* mean_out;
    double* var_out;
    // observation epilogue
    const double* v_in;  // mean-only mode: variances supplied (tensor-core variants), skip the quadratic form
    // dynamics GP in mean-only mode: v_in holds 1 - |W_c k_rbf|^2 only; the low-rank (linear kernel) part of the variance
    // is finished here from the d + 1 extra alpha columns G_c = K_c^-1 [X,1] diag(c^2) and H_c = diag(c^2) [X,1]^T G_c
    const double* lr_h;  // [n_blocks][d + 1][d + 1]
    const double* z;
    double ll_const;  // 2 sum_j log lambda_j - c32
    double* ll;
    double* mu_out;
    double* v_out;
};

struct __align__(128) Smem {
    double B[STAGES][KC][LDB];
    double R[STAGES][KC * REC_MAX];
    double exptab[64];
    uint64_t full[STAGES], empty[STAGES];
    int tile;
};

__constant__ double c_exp_table[64] = {GPMDM_EXP_TABLE_VALUES};

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

// K* of two training records against one particle row, evaluated stage by stage so that the two
// exponentials are independent instruction chains.
//   -|a - b|^2 in difference form: the expansion |a|^2 + |b|^2 - 2ab of gpmdm.py:515 loses ~|a|^2 ulps, which
//   the ill-conditioned K^-1 amplifies; the difference form keeps K* at ~2 ulp.
template <int KIND, int DL>
__device__ __forceinline__ void kstar_pair(const double (&rec0)[REC_MAX], const double (&rec1)[REC_MAX],
                                           const ParticleRec<KIND, DL>& p, double c2last,
                                           const double* __restrict__ exptab, double& k0, double& k1) {
    constexpr int d = DL;
    double a0, a1;
    {
        const double t0 = rec0[0] - p.b[0], t1 = rec1[0] - p.b[0];
        a0 = -t0 * t0;
        a1 = -t1 * t1;
    }
#pragma unroll
    for (int j = 1; j < d; j++) {
        const double t0 = rec0[j] - p.b[j], t1 = rec1[j] - p.b[j];
        a0 = fma(-t0, t0, a0);
        a1 = fma(-t1, t1, a1);
    }
    int n0, n1;
    const double r0 = exp_reduce(a0, n0), r1 = exp_reduce(a1, n1);
    double e0 = 1.0 / 720.0, e1 = 1.0 / 720.0;
    constexpr double C[6] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0};
#pragma unroll
    for (int k = 5; k >= 0; k--) {
        e0 = fma(e0, r0, C[k]);
        e1 = fma(e1, r1, C[k]);
    }
    k0 = exp_scale(e0, n0, exptab);
    k1 = exp_scale(e1, n1, exptab);
    if (KIND == 1) {
        double l0 = c2last, l1 = c2last;
#pragma unroll
        for (int j = 0; j < d; j++) {
            l0 = fma(rec0[d + j], p.x[j], l0);
            l1 = fma(rec1[d + j], p.x[j], l1);
        }
        k0 += l0;
        k1 += l1;
    }
}

// K* of NB training records against one particle row: NB independent chains, evaluated stage by stage.
template <int KIND, int DL, int NB>
__device__ __forceinline__ void kstar_multi(const double* __restrict__ recs, int rec_stride,
                                            const ParticleRec<KIND, DL>& p, double c2last,
                                            const double* __restrict__ exptab, double (&out)[NB]) {
    constexpr int d = DL;
    double a[NB], lin[NB];
    int n[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        double rec[REC_MAX];
        load_record<KIND, DL>(recs + b * rec_stride, rec);
        const double t0 = rec[0] - p.b[0];
        double acc = -t0 * t0;
#pragma unroll
        for (int j = 1; j < d; j++) {
            const double t = rec[j] - p.b[j];
            acc = fma(-t, t, acc);
        }
        a[b] = acc;
        if (KIND == 1) {
            double l = c2last;
#pragma unroll
            for (int j = 0; j < d; j++) l = fma(rec[d + j], p.x[j], l);
            lin[b] = l;
        }
    }
    double r[NB], e[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) r[b] = exp_reduce(a[b], n[b]);
    constexpr double C[6] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0};
#pragma unroll
    for (int b = 0; b < NB; b++) e[b] = 1.0 / 720.0;
#pragma unroll
    for (int k = 5; k >= 0; k--)
#pragma unroll
        for (int b = 0; b < NB; b++) e[b] = fma(e[b], r[b], C[k]);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const double v = exp_scale(e[b], n[b], exptab);
        out[b] = KIND == 1 ? v + lin[b] : v;
    }
}

// Chunk schedule of one particle tile: column tile ct covers k-chunks [kbeg(ct), nkc).  `src` follows the chunk through
// the packed column panels incrementally (one add per chunk; the panel base is recomputed only when the tile changes).
struct ChunkCursor {
    int ct, k, nq, nct, nkc, tri;
    const double *L, *alpha, *src;
    long long n_pad;
    __device__ __forceinline__ int kbeg(int t) const { return (tri && t < nq) ? t * (TN / KC) : 0; }
    __device__ __forceinline__ const double* base(int t) const {
        return t < nq ? L + panel_row_offset(t, n_pad, tri) * LDB : alpha + (long long)(t - nq) * n_pad * LDB;
    }
    // kstart / kend >= 0 restrict the (single) column tile to the k-chunks [kstart, kend) (low-latency mode)
    __device__ __forceinline__ void init(int nq_, int nct_, int nkc_, int tri_, int ct0, const gpmdm_gp_block& b,
                                         int kstart = -1, int kend = -1) {
        nq = nq_, nct = nct_, nkc = kend >= 0 ? kend : nkc_, tri = tri_, ct = ct0;
        k = kstart >= 0 ? kstart : kbeg(ct0);
        L = b.L, alpha = b.alpha, n_pad = b.n_pad;
        src = ct < nct ? base(ct) + (long long)(k - kbeg(ct)) * (KC * LDB) : nullptr;
    }
    __device__ __forceinline__ bool done() const { return ct >= nct; }
    __device__ __forceinline__ void next() {
        src += KC * LDB;
        if (++k == nkc) {
            ct++;
            k = kbeg(ct);
            if (ct < nct) src = base(ct);
        }
    }
};

// CACHE: the A fragments (K* of this lane's particle row against every training row) are generated ONCE per particle
// tile into a per-CTA global scratch, in exactly the order the lane consumes them, and the k loop re-reads them (four
// 8-byte loads per lane and chunk, issued one chunk ahead) instead of re-evaluating the exponentials for every column
// tile: with the triangular packing a K* entry is used by nq/2 column tiles on average, so the fp64 datapath the
// DMMAs run on is relieved of ~97 % of the exp work.  Each lane reads back only what it wrote itself: no barrier,
// no fence.  Values are bit-identical to the on-the-fly instantiation (same function, same inputs).
#ifdef GPMDM_TIMELINE  /* diagnostic build (tools/lowlat_timeline.py): per-item time stamps of the work-item loop */
constexpr int TL_ITEMS = 64, TL_WORDS = 8;
__device__ unsigned long long g_timeline[160 * TL_ITEMS * TL_WORDS];
__device__ int g_timeline_n[160];
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GPMDM_TL(slot)                                                                                     \
    if (tid == 0 && tl_i < TL_ITEMS) g_timeline[((long long)blockIdx.x * TL_ITEMS + tl_i) * TL_WORDS + (slot)] = tl_now();
// first start / last end of the item, fill and finalize kernels of the LAST launch of each kind (node timeline of a frame)
__device__ unsigned long long g_node_stamps[16];  // [2 id] start (min), [2 id + 1] end (max); id = kernel (0 items, 1 finalize, 2 fill) * 2 + KIND
#define GPMDM_NODE_BEGIN(id) if (threadIdx.x == 0) atomicMin(&g_node_stamps[2 * (id)], tl_now());
#define GPMDM_NODE_END(id) if (threadIdx.x == 0) atomicMax(&g_node_stamps[2 * (id) + 1], tl_now());
#else
#define GPMDM_TL(slot)
#define GPMDM_NODE_BEGIN(id)
#define GPMDM_NODE_END(id)
#endif

// SPLIT (low-latency launches, CACHE only): the K* slices come from kstar_fill_kernel (one slice per particle TILE, shared by
// all the tile's work items), and warps whose 8 particle rows lie beyond the tile's count skip the fragment loads, the MMAs
// and the epilogues -- they only keep the ring's barriers moving -- so a ragged tile costs its rows rounded up to 8, not 64.
template <int KIND, int DL, bool CACHE, bool SPLIT = false>
__global__ void __launch_bounds__(NTHREADS, 1) gp_predict_kernel(const PredictParams prm) {
    static_assert(!SPLIT || CACHE, "the low-latency instantiation reads the shared K* slices");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, c = lane & 3;
    constexpr int d = DL;
    constexpr int REC = rec_width(KIND, d);
    constexpr int NJ = TN / 8;  // 8-column blocks per column tile
    const double c2last = KIND == 1 ? prm.lin_c2[d] : 0.0;

    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&s.full[i], 1);        // one arrive.expect_tx + TMA bytes
            mbar_init(&s.empty[i], NWARPS);  // one arrival per warp
        }
        mbar_fence_init();
    }
    if (tid < 64) s.exptab[tid] = c_exp_table[tid];
    __syncthreads();
    const double* exptab = s.exptab;

    const int total_tiles = prm.tiles ? *prm.n_tiles : (int)((prm.P + TM - 1) / TM);
    // ring positions as (stage, parity) pairs advanced incrementally (STAGES is not a power of two)
    int cst = 0, cph = 0;   // consumer: stage / parity of the chunk being consumed
    int duty = 0;           // warp whose turn it is to issue the next TMA chunk

    const int total_items = prm.split ? total_tiles * prm.max_nct * prm.nseg : total_tiles;
    int rounds_done = 0;
    GPMDM_NODE_BEGIN(0 * 2 + KIND)
#ifdef GPMDM_TIMELINE
    int tl_i = -1;
    const unsigned long long tl_enter = tl_now();
#endif
    for (;;) {
#ifdef GPMDM_TIMELINE
        tl_i++;
        GPMDM_TL(0)
#endif
        if (tid == 0) s.tile = atomicAdd(prm.counter, 1);
        __syncthreads();
        const int item = s.tile;
        __syncthreads();
#ifdef GPMDM_TIMELINE
        GPMDM_TL(1)
        if (tid == 0 && tl_i < TL_ITEMS) {
            g_timeline[((long long)blockIdx.x * TL_ITEMS + tl_i) * TL_WORDS + 6] = (unsigned long long)(unsigned)item;
            g_timeline[((long long)blockIdx.x * TL_ITEMS + tl_i) * TL_WORDS + 7] = tl_enter;
            g_timeline[((long long)blockIdx.x * TL_ITEMS + tl_i) * TL_WORDS + 2] = 0;
            g_timeline_n[blockIdx.x] = tl_i + 1;
        }
#endif
        if (item >= total_items) {
            // leaving: satisfy every later round so that nobody waits for this CTA
            if (prm.round_sync && tid == 0) atomicAdd(prm.counter + 1, 1 << 20);  // > any target (n_tiles < 2^20)
            GPMDM_NODE_END(0 * 2 + KIND)
            break;
        }
        // split mode: items are ordered column tile first, so the longest (ct = 0 of every particle tile) start first
        const int t = prm.split ? item % total_tiles : item;
        int item_ct = -1, item_s = 0;
        if (prm.split) {
            const int rest = item / total_tiles;
            item_ct = rest / prm.nseg;
            item_s = rest - item_ct * prm.nseg;
        }
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
        // k-chunks that hold real training rows: the rows padding N up to a multiple of 256 are zero in L and alpha, so
        // the chunks made of them only add exact zeros -- 14 of 1264 chunks of EVERY column tile at N = 20 000 (2.2 % of
        // the MMAs).  The panel layout (and every offset into it) still follows n_pad.
        const int nkc = (int)((gbk.n + KC - 1) / KC);
        const int nq = n_pad / TN;               // column tiles of L
        const int nct = nq + prm.alpha_ld / TN;  // + column tiles of alpha
        const int ct0 = prm.v_in ? nq : 0;  // mean-only mode (variances supplied) starts at the alpha tiles
        int ct_begin = ct0, ct_end = nct;
        int kfirst = -1, kend = nkc;  // k-chunk range of the work item (split mode: one segment of one column tile)
        if (prm.split) {
            // Every (tile, column tile, segment) item owns one slot of the partial-sum workspace for its tile's particles:
            // qpart[ct][s] when ct indexes a column tile of the LARGEST block, mu_ws[s] when ct is one of this block's alpha
            // tiles.  An item with nothing to compute (the block is smaller, or the segment starts beyond its rows) writes
            // the zeros itself, so the workspace needs no memset between launches.
            kfirst = ((prm.tri && item_ct < nq) ? item_ct * (TN / KC) : 0) + item_s * prm.seg_chunks;
            const bool is_l = item_ct >= ct0 && item_ct < nq, is_a = item_ct >= nq && item_ct < nct;
            const int max_nq = prm.max_nct - prm.alpha_ld / TN;
            if (item_ct < max_nq && !(is_l && kfirst < nkc))
                for (int m = tid; m < count; m += NTHREADS) {
                    const int p = prm.perm ? prm.perm[first + m] : first + m;
                    prm.qpart[((long long)item_ct * prm.nseg + item_s) * prm.P + p] = 0.0;
                }
            if (is_a && kfirst >= nkc) {
                const int c0 = (item_ct - nq) * TN, c1 = min(prm.dout, c0 + TN);
                for (int i = tid; i < count * (c1 - c0); i += NTHREADS) {
                    const int m = i / (c1 - c0), col = c0 + i % (c1 - c0);
                    const int p = prm.perm ? prm.perm[first + m] : first + m;
                    prm.mu_ws[((long long)item_s * prm.P + p) * prm.dout + col] = 0.0;
                }
            }
            if (!(is_l || is_a) || kfirst >= nkc) continue;  // uniform for the CTA
            ct_begin = item_ct;
            ct_end = item_ct + 1;
            kend = min(kfirst + prm.seg_chunks, nkc);
        }

        const bool active = !SPLIT || warp * 8 < count;  // warp-uniform
        // ---- this lane's particle row --------------------------------------------------------------------
        ParticleRec<KIND, DL> pr;
        int pidx;
        double prior = 1.0;
        {
            const int row = warp * 8 + r;
            const int m = row < count ? row : count - 1;
            const int p = prm.perm ? prm.perm[first + m] : first + m;
            pidx = row < count ? p : -1;
#pragma unroll
            for (int j = 0; j < d; j++) {
                const double xj = prm.x[(long long)p * d + j];
                pr.b[j] = xj / prm.ls[j];
                if (KIND == 1) {
                    pr.x[j] = xj;
                    prior = fma(prm.lin_c2[j] * xj, xj, prior);
                }
            }
            if (KIND == 1) prior += c2last;
        }

        // ---- TMA issue: all warps advance the same cursor, the duty warp issues ------------------------------
        ChunkCursor bcur;
        bcur.init(nq, ct_end, nkc, prm.tri, ct_begin, gbk, kfirst, prm.split ? kend : -1);
        int pst = cst, pph = cph;  // producer: stage / parity of the next chunk to issue
        auto issue_b = [&]() {
            const int st = pst;
            mbar_wait(&s.empty[st], pph ^ 1);  // first fill of a stage passes immediately
            // 16 consecutive rows of a column panel are contiguous in the packed factors: ONE bulk copy per chunk
            const double* src = bcur.src;
            if (lane == 0) {
                // the training records feed the on-the-fly K* prologue only: not needed with the K* cache
                mbar_expect_tx(&s.full[st], (uint32_t)(KC * LDB * 8 + (CACHE ? 0 : KC * REC * 8)));
                bulk_g2s(&s.B[st][0][0], src, KC * LDB * 8, &s.full[st]);
                if (!CACHE)
                    bulk_g2s(&s.R[st][0], gbk.coords + (long long)bcur.k * KC * REC, (uint32_t)(KC * REC * 8), &s.full[st]);
            }
        };
#define GPMDM_ADVANCE(st_, ph_)   \
    if (++(st_) == STAGES) {      \
        (st_) = 0;                \
        (ph_) ^= 1;               \
    }
        for (int i = 0; i < AHEAD && !bcur.done(); i++) {
            if (warp == 0) issue_b();
            bcur.next();
            GPMDM_ADVANCE(pst, pph)
        }

        // ---- K* cache: this lane's fragments for every chunk of the block, written in consumption order ----------
        double* kc = nullptr;
        constexpr int KCHUNK = TM * KC;  // doubles per chunk in the cache: [warp][k4][lane]
        if (CACHE) {
            if ((long long)n_pad * TM > prm.kcache_stride) __trap();  // workspace sized for a smaller block
            kc = prm.kcache + (long long)(SPLIT ? t : (int)blockIdx.x) * prm.kcache_stride + warp * (KC / 4 * 32) + lane;
            for (int k = 0; !SPLIT && k < nkc; k++) {
                double g[KC / 4];
                kstar_multi<KIND, DL, KC / 4>(gbk.coords + (long long)(k * KC + c) * REC, 4 * REC, pr, c2last, exptab, g);
#pragma unroll
                for (int i = 0; i < KC / 4; i++) __stcg(kc + (long long)k * KCHUNK + i * 32, g[i]);
            }
            __syncwarp();  // the Hadamard epilogue reads slots written by the other lanes of this warp
        }

        // ---- A fragments of the first chunk: a[k4] = K*[row 8 w + r][k = 4 k4 + c] -----------------------------
        double a[KC / 4];
        mbar_wait(&s.full[cst], cph);
        GPMDM_TL(2)
        if (CACHE) {
            const int k0 = kfirst >= 0 ? kfirst : ((prm.tri && ct_begin < nq) ? ct_begin * (TN / KC) : 0);
#pragma unroll
            for (int i = 0; i < KC / 4; i++) a[i] = active ? __ldcg(kc + (long long)k0 * KCHUNK + i * 32) : 0.0;
        } else {
            kstar_multi<KIND, DL, KC / 4>(&s.R[cst][c * REC], 4 * REC, pr, c2last, exptab, a);
        }

        ChunkCursor cur;
        cur.init(nq, ct_end, nkc, prm.tri, ct_begin, gbk, kfirst, prm.split ? kend : -1);
        double qacc = 0.0, sacc = 0.0, vrow = 0.0;
        if (prm.v_in) vrow = pidx >= 0 ? prm.v_in[pidx] : 1.0;

        for (int ct = ct_begin; ct < ct_end; ct++) {
            double acc[NJ][2];
#pragma unroll
            for (int j = 0; j < NJ; j++) acc[j][0] = acc[j][1] = 0.0;

            const int kbeg = kfirst >= 0 ? kfirst : cur.kbeg(ct);
            // The Hadamard epilogue of a tile of L re-reads the tile's 256 training records from global memory, 64
            // dependent iterations per lane: request them into L1 now (it is otherwise idle: B arrives by TMA, K* by
            // ld.cg), so that the epilogue does not pay 64 L2 round trips (~30 us per column tile).
            if (ct < nq && !CACHE) {
                const char* recs = reinterpret_cast<const char*>(gbk.coords + (long long)ct * TN * REC);
                for (int off = tid * 128; off < TN * REC * 8; off += NTHREADS * 128)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(recs + off));
            }
            // The k loop of one column tile.  GROUP_ON(jg) says whether the 8 column blocks starting at jg are
            // needed: always for tiles of L; for alpha tiles only the blocks that hold real output columns.
#ifdef GPMDM_DIAG_NO_EXP  /* timing diagnostic only: results are wrong */
#define GPMDM_MAIN_LOOP_KSTAR a[0] += 1e-300; a[1] += 1e-300; a[2] += 1e-300; a[3] += 1e-300;
#else
#define GPMDM_MAIN_LOOP_KSTAR                                                                        \
    if (CACHE) {                                                                                     \
        _Pragma("unroll") for (int i = 0; i < KC / 4; i++) a[i] = an[i];                             \
    } else {                                                                                         \
        kstar_multi<KIND, DL, KC / 4>(&s.R[stn][c * REC], 4 * REC, pr, c2last, exptab, a);           \
    }
#endif
#define constexpr_next_group(jg, k4)                                                                           \
    {                                                                                                            \
        const int njg = (jg) + 8 < jlim8 ? (jg) + 8 : 0;                                                         \
        const int nk4 = (jg) + 8 < jlim8 ? (k4) : (k4) + 1;                                                      \
        if (nk4 < KC / 4) {                                                                                      \
            _Pragma("unroll") for (int j = 0; j < 8; j++) bq[j] = s.B[st][nk4 * 4 + c][(njg + j) * 8 + r];       \
        }                                                                                                        \
    }
#define GPMDM_K_LOOP(GROUP_ON)                                                                                  \
    for (int k = kbeg; k < kend; k++) {                                                                          \
        const int st = cst;                                                                                      \
        /* keep the ring AHEAD chunks full; the duty rotates so that no warp is always the one waiting */        \
        if (!bcur.done()) {                                                                                      \
            if (warp == duty) issue_b();                                                                         \
            duty = (duty + 1) & (NWARPS - 1);                                                                    \
            bcur.next();                                                                                         \
            GPMDM_ADVANCE(pst, pph)                                                                              \
        }                                                                                                        \
        /* The next chunk provides the records for the next A fragments.  After the last chunk of the          \
           particle tile the fragments are recomputed from the current stage (values unused). */                 \
        const bool has_next = !(ct == ct_end - 1 && k == kend - 1);                                              \
        int stn = cst, phn = cph;                                                                                \
        if (has_next) { GPMDM_ADVANCE(stn, phn) }                                                                \
        /* K* cache: the next chunk's fragments are requested now and consumed after the MMA blocks */           \
        double an[KC / 4];                                                                                       \
        if (CACHE && active) {                                                                                   \
            const int kn = !has_next ? k : (k + 1 < kend ? k + 1 : cur.kbeg(ct + 1));                            \
            _Pragma("unroll") for (int i = 0; i < KC / 4; i++) an[i] = __ldcg(kc + (long long)kn * KCHUNK + i * 32); \
        }                                                                                                        \
        /* probe the next chunk's barrier now, look at the answer after the MMA blocks (hides the probe latency) */ \
        const uint32_t ready = has_next ? mbar_test(&s.full[stn], phn) : 1u;                                     \
        /* B fragments are software-pipelined one group of 8 column blocks ahead of the MMAs that use them */      \
        double bq[8];                                                                                            \
        if (active && GROUP_ON(0)) {                                                                             \
            _Pragma("unroll") for (int j = 0; j < 8; j++) bq[j] = s.B[st][c][j * 8 + r];                         \
        }                                                                                                        \
        if (active) _Pragma("unroll") for (int k4 = 0; k4 < KC / 4; k4++) {                                      \
            const double ak = a[k4];                                                                             \
            _Pragma("unroll") for (int jg = 0; jg < NJ; jg += 8) {                                               \
                if (GROUP_ON(jg)) {                                                                              \
                    double b[8];                                                                                 \
                    _Pragma("unroll") for (int j = 0; j < 8; j++) b[j] = bq[j];                                  \
                    /* next group: same k4 block, or the first group of the next block */                        \
                    constexpr_next_group(jg, k4);                                                                \
                    _Pragma("unroll") for (int j = 0; j < 8; j++)                                                \
                        dmma_m8n8k4(acc[jg + j][0], acc[jg + j][1], ak, b[j]);                                   \
                }                                                                                                \
            }                                                                                                    \
        }                                                                                                        \
        /* the next chunk's four A fragments: four independent exp chains in one block */                        \
        if (!ready) mbar_wait(&s.full[stn], phn);                                                                \
        if (active) { GPMDM_MAIN_LOOP_KSTAR }                                                                    \
        __syncwarp();                                                                                            \
        if (lane == 0) mbar_arrive(&s.empty[st]); /* this warp is done with the ring slot */                     \
        GPMDM_ADVANCE(cst, cph)                                                                                  \
    }
#define GPMDM_ALL_GROUPS(jg) true
#define GPMDM_SOME_GROUPS(jg) ((jg) < jlim)
            // 8-column blocks in use (mean-only dynamics: + the d + 1 columns of G_c)
            const int dcols = (KIND == 1 && prm.lr_h) ? 2 * DL + 1 : prm.dout;
            const int jlim = ct < nq ? NJ : (min(TN, dcols - (ct - nq) * TN) + 7) / 8;
            if (jlim > NJ - 8) {
                constexpr int jlim8 = NJ;
                GPMDM_K_LOOP(GPMDM_ALL_GROUPS)
            } else {
                const int jlim8 = (jlim + 7) & ~7;
                GPMDM_K_LOOP(GPMDM_SOME_GROUPS)
            }
#undef GPMDM_K_LOOP
#undef constexpr_next_group
            GPMDM_TL(3)
#undef GPMDM_ALL_GROUPS
#undef GPMDM_SOME_GROUPS

            // ---- epilogues (per warp; a row's columns live in the 4 lanes of a quad) -------------------------
            if (!active) {
                // nothing to finish for this warp
            } else if (ct < nq) {
                // q[p] += sum_n C[p,n] * K*[p,n] over this tile's columns
                if (CACHE) {
                    // K*[row][n] sits in the slices: column n = 16 kk + 4 k4 + c' was written (as an A fragment) by lane
                    // 4 r + c' of this warp -- this lane's columns 8 j + 2 c + {0, 1} are two adjacent doubles.  Same
                    // function, same inputs as kstar_pair below: the bits are the same, the 64 exponentials per lane and
                    // column tile are not spent.  Columns in k-chunks beyond the block's rows (never filled) multiply
                    // accumulators that are exactly zero: skipped.
                    const double* kb = kc - lane + (long long)(ct * (TN / KC)) * KCHUNK + r * 4 + 2 * (c & 1) + (c >> 1) * 32;
                    const int jmax = 2 * (nkc - ct * (TN / KC));
#pragma unroll
                    for (int j = 0; j < NJ; j++) {  // fully unrolled: acc[][] must stay in registers
                        if (j < jmax) {
                            const double* q = kb + (long long)(j >> 1) * KCHUNK + (j & 1) * 64;
                            const double k0 = __ldcg(q), k1 = __ldcg(q + 1);
                            qacc = fma(acc[j][0], k0, qacc);
                            qacc = fma(acc[j][1], k1, qacc);
                        }
                    }
                } else
#pragma unroll
                for (int j = 0; j < NJ; j++) {  // fully unrolled: acc[][] must stay in registers (K* regenerated per element)
                    const int n = ct * TN + j * 8 + c * 2;
                    double rec0[REC_MAX], rec1[REC_MAX];
                    load_record<KIND, DL>(gbk.coords + (long long)n * REC, rec0);
                    load_record<KIND, DL>(gbk.coords + (long long)(n + 1) * REC, rec1);
                    double k0, k1;
                    kstar_pair<KIND, DL>(rec0, rec1, pr, c2last, exptab, k0, k1);
                    qacc = fma(acc[j][0], k0, qacc);
                    qacc = fma(acc[j][1], k1, qacc);
                }
                if (prm.split) {
                    double v = qacc;
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (c == 0 && pidx >= 0) prm.qpart[((long long)ct * prm.nseg + item_s) * prm.P + pidx] = v;
                } else if (ct == nq - 1) {
                    // quadratic form complete: v[p] = prior[p] - q[p]
                    double v = qacc;
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    vrow = prior - v;
                    // the reference lets sqrt / log of such a variance produce NaN (gpmdm_pf.py:168, :189); so does this
                    // kernel, but it also counts the particles it happened to
                    if (c == 0 && pidx >= 0 && !(vrow > 0.0 && vrow < INFINITY)) atomicAdd(prm.status + (KIND == 0 ? 1 : 0), 1);
                }
            } else {
                const int cbase = (ct - nq) * TN;
                if (KIND == 1 && prm.lr_h && ct == nq) {
                    // var = u + x~^T S x~ - 2 t x~ + x~^T H_c x~,  t = k^T G_c = accumulator columns d .. 2d of this row
                    // (include/gpmdm_b200.h, "dynamics GP variance on the tensor cores"); a row's columns live in its quad
                    double corr = 0.0, xhx = 0.0;
                    const double* H = prm.lr_h + (long long)blk * (DL + 1) * (DL + 1);
#pragma unroll
                    for (int i = 0; i <= DL; i++) {
                        const int col = DL + i;  // compile-time after unrolling
                        const double xi = i < DL ? pr.x[(KIND == 1 && i < DL) ? i : 0] : 1.0;
                        if (c == ((col & 7) >> 1)) corr = fma(acc[col >> 3][col & 1], xi, corr);
                        double hrow = __ldg(H + i * (DL + 1) + DL);
#pragma unroll
                        for (int j = 0; j < DL; j++) hrow = fma(__ldg(H + i * (DL + 1) + j), pr.x[KIND == 1 ? j : 0], hrow);
                        xhx = fma(hrow, xi, xhx);
                    }
                    corr += __shfl_xor_sync(0xffffffffu, corr, 1);
                    corr += __shfl_xor_sync(0xffffffffu, corr, 2);
                    vrow = vrow + (prior - 1.0) - 2.0 * corr + xhx;
                    if (c == 0 && pidx >= 0 && !(vrow > 0.0 && vrow < INFINITY)) atomicAdd(prm.status, 1);
                }
#pragma unroll
                for (int j = 0; j < NJ; j++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int col = cbase + j * 8 + c * 2 + e;
                        if (col >= prm.dout) continue;
                        const double mu = acc[j][e];
                        if (prm.split) {  // finished by predict_finalize_kernel
                            if (pidx >= 0) prm.mu_ws[((long long)item_s * prm.P + pidx) * prm.dout + col] = mu;
                        } else if (KIND == 0) {
                            if (prm.z) {
                                const double dz = __ldg(prm.z + col) - mu;
                                sacc = fma(__ldg(prm.scale + col) * dz, dz, sacc);
                            }
                            if (prm.mu_out && pidx >= 0) prm.mu_out[(long long)pidx * prm.dout + col] = mu;
                        } else if (pidx >= 0) {
                            const double var = vrow * __ldg(prm.scale + col);
                            const long long o = (long long)pidx * prm.dout + col;
                            if (prm.x_new)  // torch.normal: randn * std + mean, two roundings (no FMA)
                                prm.x_new[o] = __dadd_rn(__dmul_rn(__ldg(prm.eps + o), sqrt(var)), mu);
                            if (prm.mean_out) prm.mean_out[o] = mu;
                            if (prm.var_out) prm.var_out[o] = var;
                        }
                    }
                }
                if (KIND == 0 && !prm.split && ct == nct - 1) {
                    double S = sacc;
                    S += __shfl_xor_sync(0xffffffffu, S, 1);
                    S += __shfl_xor_sync(0xffffffffu, S, 2);
                    if (c == 0 && pidx >= 0) {
                        if (prm.ll) prm.ll[pidx] = -0.5 * S / vrow - (double)prm.dout * log(vrow) + prm.ll_const;
                        if (prm.v_out) prm.v_out[pidx] = vrow;
                    }
                }
            }
        }
        GPMDM_TL(4)
        // Round synchronisation (observation kernel, several uniform tiles per CTA): every CTA waits, with a bounded
        // number of polls, until all CTAs have finished the same number of tiles.  Within one round the CTAs walk L in
        // lockstep and it is read from HBM once (L2 hit 97 %); without re-alignment they drift beyond the reach of
        // the 126 MB L2 over the rounds and L is re-streamed ~9x (profiles/launches_r01.txt).
        if (prm.round_sync) {
            rounds_done++;
            if (tid == 0) {
                atomicAdd(prm.counter + 1, 1);
                const int target = rounds_done * (int)gridDim.x;
                for (int spin = 0; spin < 200000; spin++) {  // bounded: a missing CTA costs time, never a hang
                    if (*reinterpret_cast<volatile int*>(prm.counter + 1) >= target) break;
                    __nanosleep(100);
                }
            }
            __syncthreads();
        }
    }
}

// Low-latency launches: the K* slice of every particle TILE ([chunk][warp][k4][lane] = the order gp_predict_kernel's lanes
// consume A fragments in), written once by a wide grid and read by all the tile's (column tile, k segment) work items.
// Same function, same inputs as the per-CTA fill of the fused CACHE instantiation: the values are bit-identical.
constexpr int FILL_CHUNKS = 8;  // k-chunks per CTA
template <int KIND, int DL>
__global__ void __launch_bounds__(NTHREADS) kstar_fill_kernel(const PredictParams prm) {
    __shared__ double exptab[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, c = lane & 3;
    constexpr int REC = rec_width(KIND, DL);
    if (tid < 64) exptab[tid] = c_exp_table[tid];
    __syncthreads();
    const int total_tiles = prm.tiles ? *prm.n_tiles : (int)((prm.P + TM - 1) / TM);
    const int t = blockIdx.x;
    if (t >= total_tiles) return;
    int blk = 0, first = t * TM, count;
    if (prm.tiles) {
        blk = prm.tiles[4 * t + 0];
        first = prm.tiles[4 * t + 1];
        count = prm.tiles[4 * t + 2];
    } else {
        const long long rem = prm.P - (long long)first;
        count = rem < TM ? (int)rem : TM;
    }
    if (warp * 8 >= count) return;  // rows beyond the tile's count: their slots are never read
    const gpmdm_gp_block gbk = prm.blocks[blk];
    if ((long long)gbk.n_pad * TM > prm.kcache_stride) __trap();
    const int nkc = (int)((gbk.n + KC - 1) / KC);
    const double c2last = KIND == 1 ? prm.lin_c2[DL] : 0.0;
    ParticleRec<KIND, DL> pr;
    {
        const int row = warp * 8 + r;
        const int m = row < count ? row : count - 1;
        const int p = prm.perm ? prm.perm[first + m] : first + m;
#pragma unroll
        for (int j = 0; j < DL; j++) {
            const double xj = prm.x[(long long)p * DL + j];
            pr.b[j] = xj / prm.ls[j];
            if (KIND == 1) pr.x[j] = xj;
        }
    }
    constexpr int KCHUNK = TM * KC;
    double* kc = prm.kcache + (long long)t * prm.kcache_stride + warp * (KC / 4 * 32) + lane;
    const int k1 = min(nkc, ((int)blockIdx.y + 1) * FILL_CHUNKS);
    for (int k = (int)blockIdx.y * FILL_CHUNKS; k < k1; k++) {
        double g[KC / 4];
        kstar_multi<KIND, DL, KC / 4>(gbk.coords + (long long)(k * KC + c) * REC, 4 * REC, pr, c2last, exptab, g);
#pragma unroll
        for (int i = 0; i < KC / 4; i++) __stcg(kc + (long long)k * KCHUNK + i * 32, g[i]);
    }
}

// Low-latency mode, second kernel: one thread per particle adds the per-column-tile contributions in a fixed order and
// runs the epilogue the fused kernel would have run (log-likelihood, or the Gaussian draw).
// One warp per particle: lanes stride over the (column tile, segment) partials and over the output columns.
template <int KIND>
__global__ void __launch_bounds__(128) predict_finalize_kernel(const PredictParams prm, int max_nq) {
    const long long p = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    GPMDM_NODE_BEGIN(1 * 2 + KIND)
    // the item kernel is done with the hand-out counter: leave it zeroed for the next low-latency launch of a fixed sequence
    // (csrc/pf_small.cu issues the dynamics and the observation launch back to back without memset nodes in between)
    if (blockIdx.x == 0 && threadIdx.x == 0) prm.counter[0] = 0, prm.counter[1] = 0;
    if (p >= prm.P) return;
    const int d = prm.d;
    // fixed order: lane l adds entries l, l + 32, ...; then a shuffle tree
    double q = 0.0;
    for (int i = lane; i < max_nq * prm.nseg; i += 32) q += prm.qpart[(long long)i * prm.P + p];
    q = warp_sum(q);
    auto mean = [&](int j) {
        double m = 0.0;
        for (int sg = 0; sg < prm.nseg; sg++) m += prm.mu_ws[((long long)sg * prm.P + p) * prm.dout + j];
        return m;
    };
    if (KIND == 0) {
        const double v = prm.v_in ? prm.v_in[p] : 1.0 - q;
        if (lane == 0 && !prm.v_in && !(v > 0.0 && v < INFINITY)) atomicAdd(prm.status + 1, 1);
        double S = 0.0;
        for (int j = lane; j < prm.dout; j += 32) {
            const double mu = mean(j);
            if (prm.z) {
                const double dz = prm.z[j] - mu;
                S = fma(prm.scale[j] * dz, dz, S);
            }
            if (prm.mu_out) prm.mu_out[p * prm.dout + j] = mu;
        }
        S = warp_sum(S);
        if (lane == 0) {
            if (prm.ll) prm.ll[p] = -0.5 * S / v - (double)prm.dout * log(v) + prm.ll_const;
            if (prm.v_out) prm.v_out[p] = v;
        }
    } else {
        double prior = 1.0;
        for (int j = 0; j < d; j++) {
            const double xj = prm.x[p * d + j];
            prior = fma(prm.lin_c2[j] * xj, xj, prior);
        }
        prior += prm.lin_c2[d];
        const double v = prior - q;
        if (lane == 0 && !(v > 0.0 && v < INFINITY)) atomicAdd(prm.status, 1);
        for (int k = lane; k < prm.dout; k += 32) {
            const double var = v * prm.scale[k], mu = mean(k);
            const long long o = p * prm.dout + k;
            if (prm.x_new) prm.x_new[o] = __dadd_rn(__dmul_rn(prm.eps[o], sqrt(var)), mu);
            if (prm.mean_out) prm.mean_out[o] = mu;
            if (prm.var_out) prm.var_out[o] = var;
        }
    }
    GPMDM_NODE_END(1 * 2 + KIND)  // (by the first warp of every block; the other warps finish within a few hundred ns of it)
}

// ---- host side -------------------------------------------------------------------------------------
template <int KIND, int DL, bool CACHE = false, bool SPLIT = false>
static int launch_instance(const PredictParams& prm, int grid, cudaStream_t st) {
    // the opt-in to > 48 KB of dynamic shared memory is per function AND per device
    static bool configured[64] = {};  // per instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    auto kern = gp_predict_kernel<KIND, DL, CACHE, SPLIT>;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gp_predict): %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured[dev] = true;
    }
    kern<<<grid, NTHREADS, sizeof(Smem), st>>>(prm);
    return check_launch("gp_predict_kernel");
}

template <int KIND>
static int dispatch_d_cached(const PredictParams& prm, int grid, cudaStream_t st) {
    switch (prm.d) {
        case 1: return launch_instance<KIND, 1, true>(prm, grid, st);
        case 2: return launch_instance<KIND, 2, true>(prm, grid, st);
        case 3: return launch_instance<KIND, 3, true>(prm, grid, st);
        case 4: return launch_instance<KIND, 4, true>(prm, grid, st);
        case 5: return launch_instance<KIND, 5, true>(prm, grid, st);
        case 6: return launch_instance<KIND, 6, true>(prm, grid, st);
        case 7: return launch_instance<KIND, 7, true>(prm, grid, st);
        case 8: return launch_instance<KIND, 8, true>(prm, grid, st);
    }
    set_error("latent dimension %d outside [1, %d]", prm.d, MAXD);
    return GPMDM_E_UNSUPPORTED;
}

// low-latency launches: K* slices per particle tile (kstar_fill_kernel), then the (tile, column tile, k segment) items
template <int KIND, int DL>
static int launch_split(const PredictParams& prm, int grid, int tiles_bound, int max_nkc, cudaStream_t st) {
    kstar_fill_kernel<KIND, DL><<<dim3((unsigned)tiles_bound, (unsigned)((max_nkc + FILL_CHUNKS - 1) / FILL_CHUNKS)), NTHREADS, 0,
                                  st>>>(prm);
    if (int rc = check_launch("kstar_fill_kernel")) return rc;
    return launch_instance<KIND, DL, true, true>(prm, grid, st);
}

template <int KIND>
static int dispatch_d_split(const PredictParams& prm, int grid, int tiles_bound, int max_nkc, cudaStream_t st) {
    switch (prm.d) {
        case 1: return launch_split<KIND, 1>(prm, grid, tiles_bound, max_nkc, st);
        case 2: return launch_split<KIND, 2>(prm, grid, tiles_bound, max_nkc, st);
        case 3: return launch_split<KIND, 3>(prm, grid, tiles_bound, max_nkc, st);
        case 4: return launch_split<KIND, 4>(prm, grid, tiles_bound, max_nkc, st);
        case 5: return launch_split<KIND, 5>(prm, grid, tiles_bound, max_nkc, st);
        case 6: return launch_split<KIND, 6>(prm, grid, tiles_bound, max_nkc, st);
        case 7: return launch_split<KIND, 7>(prm, grid, tiles_bound, max_nkc, st);
        case 8: return launch_split<KIND, 8>(prm, grid, tiles_bound, max_nkc, st);
    }
    set_error("latent dimension %d outside [1, %d]", prm.d, MAXD);
    return GPMDM_E_UNSUPPORTED;
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

static int propagate_impl(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm, const int32_t* tiles,
                          const int32_t* n_tiles, int64_t P, const double* eps, double* x_new, double* mean_out,
                          double* var_out, int32_t* tile_counter, void* stream, void* kstar_ws, int64_t kstar_ws_bytes,
                          int64_t max_n_pad, const double* v_in = nullptr, const double* lr_h = nullptr) {
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
    prm.status = tile_counter + 2;
    prm.eps = eps;
    prm.x_new = x_new;
    prm.mean_out = mean_out;
    prm.var_out = var_out;
    prm.v_in = v_in;
    prm.lr_h = lr_h;
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, 2 * sizeof(int32_t), st);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const long long max_tiles = (P + TM - 1) / TM + dyn->n_blocks;
    const int grid = (int)(max_tiles < num_sms() ? max_tiles : num_sms());
    if (kstar_ws) {
        prm.kcache = static_cast<double*>(kstar_ws);
        prm.kcache_stride = (long long)max_n_pad * TM;  // the kernel traps if a block on the device is larger
        GPMDM_REQUIRE((int64_t)grid * prm.kcache_stride * 8 <= kstar_ws_bytes, GPMDM_E_INVALID,
                      "K* workspace too small: %lld bytes for %d CTAs x n_pad %lld", (long long)kstar_ws_bytes, grid,
                      (long long)max_n_pad);
        return dispatch_d_cached<1>(prm, grid, st);
    }
    return dispatch_d<1>(prm, grid, st);
}

extern "C" int gpmdm_pf_propagate_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                                      const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                                      double* x_new, double* mean_out, double* var_out, int32_t* tile_counter,
                                      void* stream) {
    return propagate_impl(dyn, x_prev, perm, tiles, n_tiles, P, eps, x_new, mean_out, var_out, tile_counter, stream, nullptr,
                          0, 0);
}

// The same call with the per-CTA K* cache of gpmdm_pf_observe_cached_f64 (one slice of max_n_pad x 64 doubles per CTA; the
// two calls of a step may share the scratch): each class block's cross-kernel is evaluated once per particle tile.
// Bit-identical results; pays from ~4 column panels per block (N_c >= 1024).
extern "C" int gpmdm_pf_propagate_cached_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                                             const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                                             double* x_new, double* mean_out, double* var_out, int64_t max_n_pad,
                                             int32_t* tile_counter, void* kstar_ws, int64_t kstar_ws_bytes, void* stream) {
    GPMDM_REQUIRE(kstar_ws != nullptr && kstar_ws_bytes > 0, GPMDM_E_INVALID, "K* workspace is required");
    GPMDM_REQUIRE(max_n_pad > 0 && max_n_pad % TN == 0, GPMDM_E_INVALID,
                  "max_n_pad must be the largest padded block size (multiple of %d)", TN);
    return propagate_impl(dyn, x_prev, perm, tiles, n_tiles, P, eps, x_new, mean_out, var_out, tile_counter, stream, kstar_ws,
                          kstar_ws_bytes, max_n_pad);
}

static int observe_impl(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                        const double* v_in, double* ll, double* mu_out, double* v_out, int32_t* tile_counter,
                        void* stream, void* kstar_ws = nullptr, int64_t kstar_ws_bytes = 0, int64_t n_pad = 0);

extern "C" int gpmdm_pf_observe_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z,
                                    double ll_const, double* ll, double* mu_out, double* v_out, int32_t* tile_counter,
                                    void* stream) {
    return observe_impl(obs, x, P, z, ll_const, nullptr, ll, mu_out, v_out, tile_counter, stream);
}

// Same call with a scratch for the K* cache (see gp_predict_kernel<.., CACHE>): one slice of n_pad x 64 doubles per CTA.
extern "C" int64_t gpmdm_pf_observe_kstar_workspace_bytes(int64_t n_pad) {
    return (int64_t)num_sms() * n_pad * TM * 8;
}

extern "C" int gpmdm_pf_observe_cached_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z,
                                           double ll_const, double* ll, double* mu_out, double* v_out,
                                           int64_t n_pad, int32_t* tile_counter, void* kstar_ws,
                                           int64_t kstar_ws_bytes, void* stream) {
    GPMDM_REQUIRE(kstar_ws != nullptr && kstar_ws_bytes > 0, GPMDM_E_INVALID, "K* workspace is required");
    GPMDM_REQUIRE(n_pad > 0 && n_pad % TN == 0, GPMDM_E_INVALID, "n_pad must be the block's padded size (multiple of %d)", TN);
    return observe_impl(obs, x, P, z, ll_const, nullptr, ll, mu_out, v_out, tile_counter, stream, kstar_ws,
                        kstar_ws_bytes, n_pad);
}

extern "C" int gpmdm_pf_propagate_meanonly_f64(const gpmdm_gp_model* dyn, const double* lowrank_h, const double* x_prev,
                                               const int32_t* perm, const int32_t* tiles, const int32_t* n_tiles,
                                               int64_t P, const double* eps, const double* u_in, double* x_new,
                                               double* mean_out, double* var_out, int32_t* tile_counter, void* stream) {
    GPMDM_REQUIRE(u_in != nullptr && lowrank_h != nullptr, GPMDM_E_INVALID, "u_in and lowrank_h are required");
    GPMDM_REQUIRE(dyn && 2 * dyn->d + 1 <= dyn->alpha_ld, GPMDM_E_INVALID, "alpha tile too narrow for [alpha | G]");
    return propagate_impl(dyn, x_prev, perm, tiles, n_tiles, P, eps, x_new, mean_out, var_out, tile_counter, stream, nullptr,
                          0, 0, u_in, lowrank_h);
}

extern "C" int gpmdm_pf_loglik_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z,
                                   double ll_const, const double* v_in, double* ll, double* mu_out,
                                   int32_t* tile_counter, void* stream) {
    GPMDM_REQUIRE(v_in != nullptr, GPMDM_E_INVALID, "v_in is required");
    return observe_impl(obs, x, P, z, ll_const, v_in, ll, mu_out, nullptr, tile_counter, stream);
}

static int observe_impl(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                        const double* v_in, double* ll, double* mu_out, double* v_out, int32_t* tile_counter,
                        void* stream, void* kstar_ws, int64_t kstar_ws_bytes, int64_t ws_n_pad) {
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
    prm.status = tile_counter + 2;
    prm.z = z;
    prm.v_in = v_in;
    prm.ll_const = ll_const;
    prm.ll = ll;
    prm.mu_out = mu_out;
    prm.v_out = v_out;
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, 2 * sizeof(int32_t), st);
    GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const long long n_tiles = (P + TM - 1) / TM;
    const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
    prm.round_sync = (n_tiles >= 2ll * grid && n_tiles < (1ll << 20)) ? 1 : 0;
    if (kstar_ws) {
        prm.kcache = static_cast<double*>(kstar_ws);
        prm.kcache_stride = (long long)ws_n_pad * TM;  // the kernel traps if the block on the device is larger
        GPMDM_REQUIRE((int64_t)grid * prm.kcache_stride * 8 <= kstar_ws_bytes, GPMDM_E_INVALID,
                      "K* workspace too small: %lld bytes for %d CTAs x n_pad %lld", (long long)kstar_ws_bytes, grid,
                      (long long)ws_n_pad);
        return dispatch_d_cached<0>(prm, grid, st);
    }
    return dispatch_d<0>(prm, grid, st);
}

// ---- low-latency variants: when there are fewer particle tiles than SMs, split every tile's column tiles over CTAs ----
// k-segment length (in chunks) of the low-latency work items.  Default rule: a sixteenth of the k range, at least 16 chunks
// -- a function of the model size only (not of P or of the device), so that results do not depend on how the particles are
// batched or sharded (GPMDM.map_x_* rely on that).  A caller that knows the whole cloud (the particle filter) may pass
// an explicit length instead, e.g. from gpmdm_predict_lowlat_pick_segment: results then depend on that choice only through
// the summation order over k.
static void choose_segments(int64_t max_n_pad, int32_t seg_chunks, int& seg, int& nseg) {
    const long long nkc = max_n_pad / KC;
    long long sg = seg_chunks > 0 ? seg_chunks : (nkc + 15) / 16;
    if (seg_chunks <= 0) sg = sg < 16 ? 16 : sg;
    sg = sg > nkc ? nkc : sg;
    seg = (int)sg;
    nseg = (int)((nkc + sg - 1) / sg);
}

// workspace = per-(column tile, segment) partial sums [max_nq * nseg][P], per-segment means [nseg][P][dout], and the K*
// slices of the particle tiles [(P / 64 rounded up) + n_blocks][max_n_pad * 64]
extern "C" int64_t gpmdm_predict_lowlat_workspace_bytes(int64_t P, int64_t max_n_pad, int32_t dout, int32_t seg_chunks,
                                                        int32_t n_blocks) {
    int seg, nseg;
    choose_segments(max_n_pad, seg_chunks, seg, nseg);
    const int64_t tiles_bound = (P + TM - 1) / TM + (n_blocks > 0 ? n_blocks : 1);
    return (((max_n_pad / TN) * P + P * (int64_t)dout) * nseg + tiles_bound * max_n_pad * TM) * 8;
}

// Segment length that minimises the modelled makespan of the item grid of ONE launch on this device: `n_tiles` particle
// tiles against a block of n_pad rows (alpha_ld / 256 mean tiles) -- waves of #SMs items, each item costing its chunks
// (~2.25 us per [16 x 256] chunk at the DMMA rate with the shared K* slices) plus a fixed ~8 us (item fetch, first TMA round trip, Hadamard
// epilogue, partial-sum writes).  With 100 particles and N = 2000 it turns 88 items of 16 chunks (148 SMs: 60 % busy, 57 us)
// into 148 items of 10.
extern "C" int32_t gpmdm_predict_lowlat_pick_segment(int64_t n_tiles, int64_t n_pad, int32_t alpha_ld, int32_t tri) {
    if (n_tiles <= 0 || n_pad <= 0 || n_pad % TN != 0) return 0;
    const long long nkc = n_pad / KC, nq = n_pad / TN, na = alpha_ld / TN;
    int dflt, nseg_d;
    choose_segments(n_pad, 0, dflt, nseg_d);
    const double t_chunk = 2.25, t_item = 8.0;
    double best = 1e300;
    int best_seg = dflt;
    for (long long sg = dflt; sg >= 4; sg--) {
        long long items = na * ((nkc + sg - 1) / sg);
        for (long long J = 0; J < nq; J++) {
            const long long len = tri ? (nq - J) * (TN / KC) : nkc;
            items += (len + sg - 1) / sg;
        }
        items *= n_tiles;
        const long long waves = (items + num_sms() - 1) / num_sms();
        const double t = waves * (sg * t_chunk + t_item);
        if (t < best - 1e-9) best = t, best_seg = (int)sg;
    }
    return best_seg;
}

// counters_zero: the caller guarantees tile_counter[0..1] == 0 on the stream (a previous finalize kernel, or its own kernel).
template <int KIND>
static int run_split(PredictParams& prm, int64_t max_n_pad, int32_t seg_chunks, void* workspace, cudaStream_t st,
                     bool counters_zero) {
    GPMDM_REQUIRE(workspace != nullptr && max_n_pad > 0 && max_n_pad % TN == 0, GPMDM_E_INVALID,
                  "low-latency mode needs a workspace and max_n_pad (multiple of %d)", TN);
    const int max_nq = (int)(max_n_pad / TN);
    prm.split = 1;
    prm.max_nct = max_nq + prm.alpha_ld / TN;
    choose_segments(max_n_pad, seg_chunks, prm.seg_chunks, prm.nseg);
    // every slot of the partial-sum workspace is written by the item that owns it (zeros included): no memset
    prm.qpart = static_cast<double*>(workspace);
    prm.mu_ws = prm.qpart + (long long)max_nq * prm.nseg * prm.P;
    if (!counters_zero) {
        cudaError_t e = cudaMemsetAsync(prm.counter, 0, 2 * sizeof(int32_t), st);
        GPMDM_REQUIRE(e == cudaSuccess, (int)e, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    const long long tiles_bound = (prm.P + TM - 1) / TM + prm.n_blocks;
    const long long items = tiles_bound * prm.max_nct * prm.nseg;
    const int grid = (int)(items < num_sms() ? items : num_sms());
    // Shared K* slices (one more launch, ~10 us) pay once the launch has enough chunks per SM: a chunk costs ~2.38 us with
    // them, ~2.55 us with the exponentials evaluated inside the k loop.  The A fragments are bit-identical either way.
    const long long chunks_per_tile = (long long)max_nq * (max_n_pad / KC) / (prm.tri ? 2 : 1);
    bool shared_kstar = ((prm.P + TM - 1) / TM) * chunks_per_tile >= 128ll * num_sms();
    if (const char* force = getenv("GPMDM_LOWLAT_KSTAR"))  // tests: "shared" / "inline" force either instantiation
        shared_kstar = force[0] == 's' ? true : (force[0] == 'i' ? false : shared_kstar);
    if (shared_kstar) {
        prm.kcache = prm.mu_ws + (long long)prm.nseg * prm.P * prm.dout;
        prm.kcache_stride = (long long)max_n_pad * TM;  // the kernels trap if a block on the device is larger
        if (int rc = dispatch_d_split<KIND>(prm, grid, (int)tiles_bound, (int)(max_n_pad / KC), st)) return rc;
    } else {
        if (int rc = dispatch_d<KIND>(prm, grid, st)) return rc;
    }
    predict_finalize_kernel<KIND><<<(unsigned)((prm.P + 3) / 4), 128, 0, st>>>(prm, max_nq);
    return check_launch("predict_finalize_kernel");
}

int gpmdm::observe_lowlat_impl(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z, double ll_const,
                               const double* v_in, double* ll, double* mu_out, double* v_out, int64_t max_n_pad,
                               int32_t seg_chunks, int32_t* tile_counter, void* workspace, void* stream, bool counters_zero) {
    if (int rc = validate_model(obs, 0)) return rc;
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P = %lld out of range", (long long)P);
    if (P == 0) return 0;
    GPMDM_REQUIRE(x && tile_counter, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(ll == nullptr || z != nullptr, GPMDM_E_INVALID, "ll requested without z");
    GPMDM_REQUIRE(obs->n_blocks == 1, GPMDM_E_INVALID, "observation GP has exactly one block");
    PredictParams prm;
    fill_common(prm, obs);
    prm.x = x;
    prm.P = P;
    prm.counter = tile_counter;
    prm.status = tile_counter + 2;
    prm.z = z;
    prm.v_in = v_in;
    prm.ll_const = ll_const;
    prm.ll = ll;
    prm.mu_out = mu_out;
    prm.v_out = v_out;
    return run_split<0>(prm, max_n_pad, seg_chunks, workspace, (cudaStream_t)stream, counters_zero);
}

extern "C" int gpmdm_pf_observe_lowlat_f64(const gpmdm_gp_model* obs, const double* x, int64_t P, const double* z,
                                           double ll_const, const double* v_in, double* ll, double* mu_out,
                                           double* v_out, int64_t max_n_pad, int32_t seg_chunks, int32_t* tile_counter,
                                           void* workspace, void* stream) {
    return observe_lowlat_impl(obs, x, P, z, ll_const, v_in, ll, mu_out, v_out, max_n_pad, seg_chunks, tile_counter, workspace,
                               stream, false);
}

int gpmdm::propagate_lowlat_impl(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm, const int32_t* tiles,
                                 const int32_t* n_tiles, int64_t P, const double* eps, double* x_new, double* mean_out,
                                 double* var_out, int64_t max_n_pad, int32_t seg_chunks, int32_t* tile_counter, void* workspace,
                                 void* stream, bool counters_zero) {
    if (int rc = validate_model(dyn, 1)) return rc;
    GPMDM_REQUIRE(P >= 0 && P < (1ll << 31), GPMDM_E_INVALID, "P = %lld out of range", (long long)P);
    if (P == 0) return 0;
    GPMDM_REQUIRE(x_prev && perm && tiles && n_tiles && tile_counter, GPMDM_E_INVALID, "null argument");
    GPMDM_REQUIRE(x_new == nullptr || eps != nullptr, GPMDM_E_INVALID, "x_new requested without eps");
    GPMDM_REQUIRE(dyn->dout == dyn->d, GPMDM_E_INVALID, "dynamics GP must have dout == d");
    PredictParams prm;
    fill_common(prm, dyn);
    prm.x = x_prev;
    prm.perm = perm;
    prm.tiles = tiles;
    prm.n_tiles = n_tiles;
    prm.P = P;
    prm.counter = tile_counter;
    prm.status = tile_counter + 2;
    prm.eps = eps;
    prm.x_new = x_new;
    prm.mean_out = mean_out;
    prm.var_out = var_out;
    return run_split<1>(prm, max_n_pad, seg_chunks, workspace, (cudaStream_t)stream, counters_zero);
}

extern "C" int gpmdm_pf_propagate_lowlat_f64(const gpmdm_gp_model* dyn, const double* x_prev, const int32_t* perm,
                                             const int32_t* tiles, const int32_t* n_tiles, int64_t P, const double* eps,
                                             double* x_new, double* mean_out, double* var_out, int64_t max_n_pad,
                                             int32_t seg_chunks, int32_t* tile_counter, void* workspace, void* stream) {
    return propagate_lowlat_impl(dyn, x_prev, perm, tiles, n_tiles, P, eps, x_new, mean_out, var_out, max_n_pad, seg_chunks,
                                 tile_counter, workspace, stream, false);
}

#ifdef GPMDM_TIMELINE
// diagnostic build only: copies out and clears the per-item time stamps ([160][64][8] words, [160] counts)
// node stamps ([16] words) of the last item / finalize launches; reset to (max, 0) pairs afterwards
extern "C" int gpmdm_debug_node_stamps(unsigned long long* host16) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host16, gpmdm::g_node_stamps, sizeof(unsigned long long) * 16);
    unsigned long long init[16];
    for (int i = 0; i < 16; i++) init[i] = (i & 1) ? 0ull : ~0ull;
    cudaMemcpyToSymbol(gpmdm::g_node_stamps, init, sizeof(init));
    return 0;
}

extern "C" int gpmdm_debug_timeline(unsigned long long* host_words, int* host_counts) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_words, gpmdm::g_timeline, sizeof(unsigned long long) * 160 * gpmdm::TL_ITEMS * gpmdm::TL_WORDS);
    cudaMemcpyFromSymbol(host_counts, gpmdm::g_timeline_n, sizeof(int) * 160);
    static int zeros[160];
    cudaMemcpyToSymbol(gpmdm::g_timeline_n, zeros, sizeof(zeros));
    return 0;
}
#endif
