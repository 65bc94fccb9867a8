// K3 / K4 on the 5th-generation tensor cores: the two-phase contrastive-similarity engine with
// tcgen05.mma (bf16 operands, fp32 accumulators in TMEM), operands staged by TMA (128B swizzle) and
// a warp-specialised mbarrier pipeline.  One CTA owns a stripe of BM = 128 anchor rows (resident in
// shared memory for the whole kernel) and walks key tiles of BN keys.
//
//   warp 0      TMA producer: A stripe once, then one B tile (BN keys x Dp, as Dp/64 swizzled chunks)
//               per pipeline stage
//   warp 1      MMA issuer (one thread):
//                 MMA1  S[buf]  = A . B_tile^T        (K = Dp; both operands K-major)      -> s_full
//                 MMA2  dA     += H[hb] . B_tile      (K = BN; H K-major, B MN-major: the SAME
//                                                       shared-memory tile read transposed)  phase 2 only
//               issued software-pipelined: MMA1(t+1) goes out before MMA2(t), so the tensor pipe
//               computes the next similarity tile while the epilogue warps turn tile t into H
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, half of the tile columns each): thread owns anchor
//               row r (TMEM lane r): tcgen05.ld the S row segment, apply the
//               per-pair functor of sim_common.cuh (masks / exp / weights),
//                 phase 1: accumulate the row statistics in registers
//                 phase 2: write H (bf16, swizzled K-major) to shared memory for MMA2
//
// TMEM: S double-buffered (2 x BN columns) + dA (<= 256 columns).  Output dims beyond 256 are covered
// by blockIdx.z halves (S is recomputed per half).  The N x N matrix never leaves the SM.
//
// Reference math: models/infonce.py:70-98, models/contrastive.py:3-169 via sim_common.cuh.
#define SIM_EXP(x) __expf(x)
#include "sim_common.cuh"
#include "tc_common.cuh"

#include <algorithm>

int sim_check_aux(const SimAux& a, int phase);
SimAux sim_make_aux(int mode, int N, int row_offset, float temperature, const float* y, const float* yhat, float w_thr,
                    float e_push, const int64_t* key, int C, float coef_multi, const float* wrow, const float* wcol,
                    const float* rs_row, const float* rs_col);

namespace {

constexpr int BM = 128;
constexpr int CHUNK_K = 64;                 // bf16 elements per 128-byte swizzled row
constexpr int A_CHUNK_BYTES = BM * 128;     // one K-chunk of the A stripe
constexpr int H_ATOM_BYTES = BM * 128;      // one [128][64] bf16 swizzle atom of H
constexpr int MAX_STAGES = 8;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_DA_COL = 256;

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

using namespace tc;      // PTX wrappers (mbarrier, TMA, tcgen05) and descriptor builders: tc_common.cuh

// instruction descriptor: bf16 x bf16 -> f32, M = 128, N = n; b_mn = 1: B operand is MN-major
__host__ __device__ constexpr uint32_t instr_desc(int n, int b_mn) { return instr_desc_mn(BM, n, 0, b_mn); }

struct TcParams {
    SimAux aux;
    float* out;          // phase 1: stats (M, 8); phase 2: dA (M, ldout)
    long long ldout;
    int M, D, nkc, nstage, tiles_per_split, use_atomics;
    __nv_bfloat16* hout; // phase 3: H tiles (M, ldh) bf16
    long long ldh;
};

constexpr int NUM_EPI_WARPS = 8;            // two warps per TMEM lane quadrant, each takes half of the tile's columns
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + NUM_EPI_THREADS;
constexpr int META_F = 6;                   // staged per-column floats: y, yhat, wrow, wcol, rs0, rs1
constexpr float LOG2E_F = 1.4426950408889634f;

// H (the bf16 gradient-coefficient tile, A operand of MMA2): [128 rows][BN] K-major.  BN >= 64: 128-byte rows,
// 128B swizzle, one 16 KB atom per 64 columns.  BN == 32: 64-byte rows, 64B swizzle (8 KB).
__host__ __device__ constexpr uint32_t h_buf_bytes(int bn) { return bn == 32 ? 128u * 64u : (uint32_t)((bn + 63) / 64) * H_ATOM_BYTES; }

struct SmemLayout {
    uint32_t a_off, b_off, h_off, meta_off, bar_off, total;
    uint32_t b_stage_bytes, meta_buf_bytes;
};
__host__ __device__ inline SmemLayout smem_layout(int nkc, int bn, int nstage, int phase) {
    SmemLayout s;
    s.a_off = 0;
    s.b_off = nkc * A_CHUNK_BYTES;
    s.b_stage_bytes = (phase != 2 ? 1 : nkc) * bn * 128;      // phases 1 / 3 stream single 64-wide K chunks, phase 2 whole key tiles
    s.h_off = s.b_off + nstage * s.b_stage_bytes;
    const uint32_t h_bytes = phase == 2 ? 2u * h_buf_bytes(bn) : 0u;
    s.meta_off = s.h_off + h_bytes;
    s.meta_buf_bytes = bn * (META_F * 4 + 8);
    s.bar_off = s.meta_off + 2 * s.meta_buf_bytes;
    s.total = s.bar_off + 256;
    return s;
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(NUM_EPI_THREADS) : "memory"); }
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

// per-row constants of an epilogue thread (registers) and per-column constants of a tile (shared memory)
struct RowCtx {
    float y, yh, wr, wc, ri0, ri1;
    long long key;
    int gi;
};
struct ColMeta {
    const float *y, *yh, *wr, *wc, *rs0, *rs1;
    const long long* key;
};
struct TcConst {
    float inv_t, k2, k2t, w_thr, e_push;    // k2 = log2(e)/t, k2t = k2 (exponent shift of InfoNCE: exp(z - 1/t))
};

// ---- phase 1, one similarity value (column c of the tile, global key j).  Branch-free: every update is a
// select + accumulate, so a warp whose lanes see different masks does not serialise (the epilogue warps are the
// pacing stage for SupCon / ConR).  Counts are integers, the positive logits are summed un-scaled.
struct P1Acc {
    float e_pos, e_neg, sdot;
    int n_pos, n_neg;
};
template <int MODE>
__device__ __forceinline__ void p1_elem(const TcConst& k, const RowCtx& r, const ColMeta& m, P1Acc& a, float dot, int j, int c) {
    if (MODE == SIM_INFONCE) {
        a.e_pos += ex2f(fmaf(dot, k.k2, -k.k2t));
        a.sdot += (r.gi == j) ? dot : 0.f;
    } else {
        bool pos, neg;
        float w = r.wr * m.wc[c];                 // r.wr carries e_push in the regression mode
        if (MODE == SIM_REGRESS) {
            const float l = fabsf(__fsub_rn(r.y, m.y[c])), pd = fabsf(__fsub_rn(r.yh, m.yh[c]));
            const bool close = l <= k.w_thr;
            pos = close && (r.gi != j);
            neg = (!close) && (pd <= k.w_thr);
            w *= l;
        } else {
            const bool same = r.key == m.key[c];
            pos = same && (r.gi != j);
            neg = !same;
        }
        const float e = ex2f(dot * k.k2);
        a.e_pos += pos ? e : 0.f;
        a.sdot += pos ? dot : 0.f;
        a.n_pos += pos ? 1 : 0;
        a.e_neg = fmaf(neg ? w : 0.f, e, a.e_neg);
        a.n_neg += neg ? 1 : 0;
    }
}
// ---- phases 2 / 3, gradient coefficient H_ij (branch-free).  Row constants are pre-combined by the caller:
//   r.ri0 = c_i, r.ri1 = alpha_i, r.wr = alpha_i * wrow_i (* e_push), r.wc = wcol_i (* e_push);
//   per column: m.rs0 = c_j, m.rs1 = alpha_j, m.wr = alpha_j * wrow_j, m.wc = wcol_j.
template <int MODE>
__device__ __forceinline__ float p2_elem(const TcConst& k, const RowCtx& r, const ColMeta& m, float dot, int j, int c) {
    if (MODE == SIM_INFONCE) {       // ri0 / rs0 hold lse * log2(e)
        const float h = ex2f(fmaf(dot, k.k2, -r.ri0)) + ex2f(fmaf(dot, k.k2, -m.rs0[c]));
        return h - ((r.gi == j) ? 2.f : 0.f);
    }
    bool pos, neg;
    float e = ex2f(dot * k.k2);
    const float hp = fmaf(r.ri1 + m.rs1[c], e, -(r.ri0 + m.rs0[c]));
    if (MODE == SIM_REGRESS) {
        const float l = fabsf(__fsub_rn(r.y, m.y[c])), pd = fabsf(__fsub_rn(r.yh, m.yh[c]));
        const bool close = l <= k.w_thr;
        pos = close && (r.gi != j);
        neg = (!close) && (pd <= k.w_thr);
        e *= l;
    } else {
        const bool same = r.key == m.key[c];
        pos = same && (r.gi != j);
        neg = !same;
    }
    const float hn = fmaf(r.wr, m.wc[c], r.wc * m.wr[c]) * e;
    return pos ? hp : (neg ? hn : 0.f);
}

template <int PHASE, int BN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 128-byte swizzle atoms need 1024-byte aligned tiles: align by hand (the launch adds 1 KB of slack)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const SimAux& aux = p.aux;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkc = p.nkc, nstage = p.nstage;
    const SmemLayout lay = smem_layout(nkc, BN, nstage, PHASE);
    unsigned char* sA = smem + lay.a_off;
    unsigned char* sB = smem + lay.b_off;
    unsigned char* sH = smem + lay.h_off;
    unsigned char* sMeta = smem + lay.meta_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    uint64_t* a_full = bars + 0;
    uint64_t* d_full = bars + 1;
    uint64_t* s_full = bars + 2;      // [2]
    uint64_t* s_empty = bars + 4;     // [2]
    uint64_t* h_full = bars + 6;      // [2]
    uint64_t* h_empty = bars + 8;     // [2]
    uint64_t* full = bars + 10;       // [MAX_STAGES]
    uint64_t* empty = bars + 18;      // [MAX_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int i0 = blockIdx.x * BM;
    const int ntiles_all = (aux.N + BN - 1) / BN;
    const int t_begin = blockIdx.y * p.tiles_per_split;
    const int T = min(ntiles_all, t_begin + p.tiles_per_split) - t_begin;      // tiles of this CTA (>= 1)
    const int dcol0 = PHASE == 2 ? blockIdx.z * 256 : 0;
    const int ND = PHASE == 2 ? min(nkc * CHUNK_K - dcol0, 256) : 0;          // output columns of this CTA
    constexpr uint32_t da_col = TMEM_DA_COL;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(d_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], NUM_EPI_THREADS);
            mbar_init(&h_full[i], NUM_EPI_THREADS);
            mbar_init(&h_empty[i], 1);
        }
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================= TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(a_full, (uint32_t)(nkc * A_CHUNK_BYTES));
            for (int kc = 0; kc < nkc; ++kc) tma_load_2d(sA + kc * A_CHUNK_BYTES, &tmA, kc * CHUNK_K, i0, a_full);
            // ring positions are carried as (slot, parity) counters: no runtime divisions in the single-thread loops
            int st = 0, par = 1;                    // par = parity of the PREVIOUS use of the slot (first pass: nothing to wait for)
            bool wrapped = false;
            if (PHASE != 2) {
                // K-chunk ring: every stage holds one (BN keys x 64) chunk; many small loads in flight
                for (int t = 0; t < T; ++t) {
                    const int j0 = (t_begin + t) * BN;
                    for (int kc = 0; kc < nkc; ++kc) {
                        if (wrapped) mbar_wait_g(&empty[st], par);
                        mbar_arrive_expect_tx(&full[st], lay.b_stage_bytes);
                        tma_load_2d(sB + (size_t)st * lay.b_stage_bytes, &tmB, kc * CHUNK_K, j0, &full[st]);
                        if (++st == nstage) { st = 0; par ^= 1; wrapped = true; }
                    }
                }
            } else {
                for (int t = 0; t < T; ++t) {
                    if (wrapped) mbar_wait_g(&empty[st], par);
                    mbar_arrive_expect_tx(&full[st], lay.b_stage_bytes);
                    unsigned char* dst = sB + (size_t)st * lay.b_stage_bytes;
                    const int j0 = (t_begin + t) * BN;
                    for (int kc = 0; kc < nkc; ++kc) tma_load_2d(dst + kc * BN * 128, &tmB, kc * CHUNK_K, j0, &full[st]);
                    if (++st == nstage) { st = 0; par ^= 1; wrapped = true; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = instr_desc(BN, 0);
            const uint32_t idesc_d = instr_desc(ND > 0 ? ND : 16, 1);
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB), h_addr = smem_u32(sH);
            // descriptors are built once; the loops only add (byte offset >> 4) to the 14-bit start-address field
            // (the issuing thread is a single dependent instruction stream: every ALU op here delays the tensor pipe)
            const uint64_t adesc0 = smem_desc(a_addr, 16, 1024);
            const uint64_t bdesc0 = smem_desc(b_addr, 16, 1024);                                               // K-major keys (MMA1)
            const uint64_t bdesc2 = smem_desc(b_addr + (dcol0 / CHUNK_K) * BN * 128, BN * 128, 1024);          // MN-major keys (MMA2)
            const uint64_t hdesc0 = BN == 32 ? smem_desc(h_addr, 16, 512, 4) : smem_desc(h_addr, 16, 1024);
            const uint32_t stage16 = lay.b_stage_bytes >> 4;
            int st2 = 0;                            // ring slot of the tile whose MMA2 is issued next
            auto mma2 = [&](int u) {
                const int hb = u & 1;
                mbar_wait_g(&h_full[hb], (u >> 1) & 1);
                tc_fence_after();
                const uint64_t bst = bdesc2 + (uint64_t)(st2 * stage16);
                const uint64_t hst = hdesc0 + (uint64_t)(hb * (h_buf_bytes(BN) >> 4));
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks) {
                    const uint64_t ad = BN == 32 ? hst + (uint64_t)(ks * 2) : hst + (uint64_t)((ks >> 2) * (H_ATOM_BYTES >> 4) + (ks & 3) * 2);
                    const uint64_t bd = bst + (uint64_t)(ks * 16 * 128 >> 4);
                    tc_mma(tmem_base + da_col, ad, bd, idesc_d, (u > 0 || ks > 0) ? 1u : 0u);
                }
                tc_commit(&h_empty[hb]);
                tc_commit(&empty[st2]);
                if (++st2 == nstage) st2 = 0;
            };
            mbar_wait_g(a_full, 0);
            int st = 0;
            uint32_t par = 0;
            for (int t = 0; t < T; ++t) {
                const int buf = t & 1;
                const uint32_t d_s = tmem_base + buf * BN;
                if (t >= 2) mbar_wait_g(&s_empty[buf], ((t >> 1) - 1) & 1);
                if (PHASE != 2) {
                    uint64_t ad = adesc0;
                    for (int kc = 0; kc < nkc; ++kc) {
                        mbar_wait_g(&full[st], par);
                        tc_fence_after();
                        const uint64_t bd = bdesc0 + (uint64_t)(st * stage16);
                        tc_mma(d_s, ad, bd, idesc_s, kc > 0 ? 1u : 0u);
                        tc_mma(d_s, ad + 2, bd + 2, idesc_s, 1u);
                        tc_mma(d_s, ad + 4, bd + 4, idesc_s, 1u);
                        tc_mma(d_s, ad + 6, bd + 6, idesc_s, 1u);
                        tc_commit(&empty[st]);
                        ad += A_CHUNK_BYTES >> 4;
                        if (++st == nstage) { st = 0; par ^= 1u; }
                    }
                    tc_commit(&s_full[buf]);
                } else {
                    mbar_wait_g(&full[st], par);
                    tc_fence_after();
                    uint64_t ad = adesc0, bd = bdesc0 + (uint64_t)(st * stage16);
                    for (int kc = 0; kc < nkc; ++kc) {
                        tc_mma(d_s, ad, bd, idesc_s, kc > 0 ? 1u : 0u);
                        tc_mma(d_s, ad + 2, bd + 2, idesc_s, 1u);
                        tc_mma(d_s, ad + 4, bd + 4, idesc_s, 1u);
                        tc_mma(d_s, ad + 6, bd + 6, idesc_s, 1u);
                        ad += A_CHUNK_BYTES >> 4;
                        bd += (uint64_t)(BN * 128 >> 4);
                    }
                    tc_commit(&s_full[buf]);
                    if (++st == nstage) { st = 0; par ^= 1u; }
                    if (t > 0) mma2(t - 1);
                }
            }
            if (PHASE == 2) {
                mma2(T - 1);
                tc_commit(d_full);
            }
        }
    } else {
        // ================================================= epilogue warps
        // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant split the tile's columns
        constexpr int HC = BN / 2;                       // columns per epilogue thread and tile
        const int ew = warp - 2;
        const int quad = warp & 3, half = ew >> 2;
        const int et = ew * 32 + lane;                   // 0..255
        const int r = quad * 32 + lane;                  // anchor row of this thread inside the stripe
        const bool row_ok = (i0 + r) < p.M;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        TcConst kc;
        kc.inv_t = aux.inv_t;
        kc.k2 = aux.inv_t * LOG2E_F;
        kc.k2t = kc.k2;
        kc.w_thr = aux.w_thr;
        kc.e_push = aux.e_push;
        RowCtx rc;
        rc.gi = aux.row_offset + i0 + r;
        rc.y = rc.yh = 0.f;
        rc.wr = rc.wc = 1.f;
        rc.ri0 = rc.ri1 = 0.f;
        rc.key = 0;
        if (row_ok) {
            const int gi = rc.gi;
            if (MODE == SIM_REGRESS) { rc.y = aux.y[gi]; rc.yh = aux.yhat[gi]; }
            if (MODE == SIM_SINGLE) rc.key = aux.key[gi];
            if (MODE != SIM_INFONCE) {
                if (aux.wrow) rc.wr = aux.wrow[gi];
                if (aux.wcol) rc.wc = aux.wcol[gi];
            }
            if (PHASE >= 2) {
                if (MODE == SIM_INFONCE) rc.ri0 = aux.rs_row[i0 + r] * LOG2E_F;
                else { rc.ri0 = aux.rs_row[2 * (i0 + r)]; rc.ri1 = aux.rs_row[2 * (i0 + r) + 1]; }
            }
        }
        RowCtx rc2 = rc;                                 // phase-2 view: pre-combined factors of p2_elem
        if (PHASE >= 2 && MODE != SIM_INFONCE) {
            const float ep = MODE == SIM_REGRESS ? aux.e_push : 1.f;
            rc2.wr = rc.ri1 * rc.wr * ep;
            rc2.wc = rc.wc * ep;
        }
        RowAcc racc;
        racc.clear();
        P1Acc pacc;
        pacc.e_pos = pacc.e_neg = pacc.sdot = 0.f;
        pacc.n_pos = pacc.n_neg = 0;
        RowCtx rc1 = rc;                                 // phase-1 view of the row constants
        if (MODE == SIM_REGRESS) rc1.wr = rc.wr * aux.e_push;
        float gri0 = 0.f, gri1 = 0.f;                    // un-scaled row statistics for the generic (tail / multi) path
        if (PHASE >= 2 && row_ok) {
            if (MODE == SIM_INFONCE) gri0 = aux.rs_row[i0 + r];
            else { gri0 = rc.ri0; gri1 = rc.ri1; }
        }
        for (int t = 0; t < T; ++t) {
            const int buf = t & 1;
            const int j0 = (t_begin + t) * BN;
            const bool fast = (MODE != SIM_MULTI) && (j0 + BN <= aux.N);
            // ---- stage the per-column constants of this tile (overlaps the MMA of the tile)
            ColMeta cm;
            {
                float* mf = reinterpret_cast<float*>(sMeta + buf * lay.meta_buf_bytes);
                long long* mk = reinterpret_cast<long long*>(mf + META_F * BN);
                cm.y = mf; cm.yh = mf + BN; cm.wr = mf + 2 * BN; cm.wc = mf + 3 * BN; cm.rs0 = mf + 4 * BN; cm.rs1 = mf + 5 * BN;
                cm.key = mk;
                if (fast && et < BN) {
                    const int j = j0 + et;
                    if (MODE == SIM_REGRESS) { mf[et] = aux.y[j]; mf[BN + et] = aux.yhat[j]; }
                    if (MODE == SIM_SINGLE) mk[et] = aux.key[j];
                    if (MODE != SIM_INFONCE) {
                        mf[2 * BN + et] = aux.wrow ? aux.wrow[j] : 1.f;
                        mf[3 * BN + et] = aux.wcol ? aux.wcol[j] : 1.f;
                    }
                    if (PHASE >= 2) {
                        if (MODE == SIM_INFONCE) mf[4 * BN + et] = aux.rs_col[j] * LOG2E_F;
                        else {
                            const float aj = aux.rs_col[2 * j + 1];
                            mf[4 * BN + et] = aux.rs_col[2 * j];
                            mf[5 * BN + et] = aj;
                            mf[2 * BN + et] = aj * (aux.wrow ? aux.wrow[j] : 1.f);       // alpha_j * wrow_j
                        }
                    }
                }
                epi_bar();
            }
            mbar_wait_g(&s_full[buf], (t >> 1) & 1);
            tc_fence_after();
            if (PHASE == 2 && t >= 2) mbar_wait_g(&h_empty[buf], ((t >> 1) - 1) & 1);
            unsigned char* hrow = sH + buf * h_buf_bytes(BN) + r * (BN == 32 ? 64 : 128);
#pragma unroll 1
            for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                uint32_t v[32];
                constexpr int NC = HC < 32 ? HC : 32;    // columns per TMEM load
                if (NC == 16) tc_ld16(lane_addr + buf * BN + c0, v);
                else tc_ld32(lane_addr + buf * BN + c0, v);
                if (PHASE == 1) {
                    if (row_ok) {
                        if (fast) {
#pragma unroll
                            for (int c = 0; c < NC; ++c) p1_elem<MODE>(kc, rc1, cm, pacc, __uint_as_float(v[c]), j0 + c0 + c, c0 + c);
                        } else {
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                const int j = j0 + c0 + c;
                                if (j < aux.N) sim_stats_accum(aux, racc, rc.gi, j, __uint_as_float(v[c]));
                            }
                        }
                    }
                } else {
                    const bool lean = fast && row_ok;
                    __nv_bfloat16* grow = PHASE == 3 ? p.hout + (long long)(i0 + r) * p.ldh + j0 + c0 : nullptr;
#pragma unroll
                    for (int g8 = 0; g8 < NC / 8; ++g8) {
                        uint32_t w[4];
                        if (lean) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = g8 * 8 + e * 2;
                                const float h0 = p2_elem<MODE>(kc, rc2, cm, __uint_as_float(v[c]), j0 + c0 + c, c0 + c);
                                const float h1 = p2_elem<MODE>(kc, rc2, cm, __uint_as_float(v[c + 1]), j0 + c0 + c + 1, c0 + c + 1);
                                w[e] = pack_bf16(h0, h1);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float h2[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int c = g8 * 8 + e * 2 + u;
                                    const int j = j0 + c0 + c;
                                    h2[u] = (row_ok && j < aux.N) ? sim_grad_coeff(aux, rc.gi, j, __uint_as_float(v[c]), gri0, gri1) : 0.f;
                                }
                                w[e] = pack_bf16(h2[0], h2[1]);
                            }
                        }
                        if (PHASE == 3) {
                            // 16 bytes = 8 consecutive keys of this thread's row (ldh is a multiple of 8, tiles start at
                            // multiples of 32): columns at or beyond ldh are never written
                            if (row_ok && j0 + c0 + g8 * 8 < p.ldh) *reinterpret_cast<uint4*>(grow + g8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                        } else {
                            const int col8 = (c0 >> 3) + g8;                   // 16-byte chunk index along K
                            unsigned char* dst = BN == 32 ? hrow + ((col8 ^ ((r >> 1) & 3)) << 4)
                                                          : hrow + (col8 >> 3) * H_ATOM_BYTES + (((col8 & 7) ^ (r & 7)) << 4);
                            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&s_empty[buf]);
            if (PHASE == 2) {
                fence_proxy_async();
                mbar_arrive(&h_full[buf]);
            }
        }
        if (PHASE == 1) {
            if (MODE == SIM_INFONCE) {
                racc.v[0] += pacc.e_pos;
                racc.v[1] = fmaf(pacc.sdot, kc.inv_t, racc.v[1]);
            } else {
                racc.v[0] += pacc.e_pos;
                racc.v[1] += pacc.e_neg;
                racc.v[2] += (float)pacc.n_pos;
                racc.v[3] += (float)pacc.n_neg;
                racc.v[4] = fmaf(pacc.sdot, kc.inv_t, racc.v[4]);
            }
            if (row_ok) {
                float* dst = p.out + (long long)(i0 + r) * SIM_NSTAT;
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    if (racc.v[k] != 0.f) atomicAdd(dst + k, racc.v[k]);      // two threads per row (+ key splits)
            }
        } else if (PHASE == 2) {
            mbar_wait_g(d_full, 0);
            tc_fence_after();
            const int ndh = ND / 2;                      // ND is a multiple of 64
            for (int c0 = half * ndh; c0 < (half + 1) * ndh; c0 += 32) {
                uint32_t v[32];
                tc_ld32(lane_addr + da_col + c0, v);
                if (row_ok) {
                    float* dst = p.out + (long long)(i0 + r) * p.ldout + dcol0 + c0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (dcol0 + c0 + c < p.D) {
                            if (p.use_atomics) atomicAdd(dst + c, __uint_as_float(v[c]));
                            else dst[c] = __uint_as_float(v[c]);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
// (rows, Dp) bf16 row-major -> 2-D map with a (64 x box_rows) box, 128-byte swizzle, zero OOB fill
int make_map(CUtensorMap* map, const void* base, int rows, int Dp, int box_rows) {
    return make_map_bf16(map, base, rows, Dp, Dp, box_rows);
}

template <int PHASE, int BN>
int launch_bn(const void* A, const void* B, int M, int N, int Dp, int D, const SimAux& aux, float* out, long long ldout,
              cudaStream_t st) {
    const int nkc = Dp / CHUNK_K;
    int nstage = MAX_STAGES;
    if (const char* e = getenv("MMDTI_SIM_STAGES")) nstage = std::max(2, std::min(MAX_STAGES, atoi(e)));      // tuning knob
    SmemLayout lay = smem_layout(nkc, BN, nstage, PHASE);
    while (nstage > 2 && lay.total + 1024 > 227 * 1024) lay = smem_layout(nkc, BN, --nstage, PHASE);
    if (lay.total + 1024 > 227 * 1024) { mmdti_set_error("sim_tc: shared memory budget exceeded (Dp=%d BN=%d)", Dp, BN); return MMDTI_ERR_ARG; }
    CUtensorMap tmA, tmB;
    if (int rc = make_map(&tmA, A, M, Dp, BM)) return rc;
    if (int rc = make_map(&tmB, B, N, Dp, BN)) return rc;
    const int stripes = (M + BM - 1) / BM;
    const int ntiles = (N + BN - 1) / BN;
    const int halves = PHASE == 2 ? (Dp + 255) / 256 : 1;
    int jsplit = 1;
    if (stripes * halves < 2 * num_sms()) jsplit = std::max(1, std::min(ntiles, (2 * num_sms()) / (stripes * halves)));
    const int per = (ntiles + jsplit - 1) / jsplit;
    jsplit = (ntiles + per - 1) / per;
    TcParams p;
    p.aux = aux; p.out = out; p.ldout = ldout; p.M = M; p.D = D; p.nkc = nkc; p.nstage = nstage;
    p.tiles_per_split = per; p.use_atomics = jsplit > 1;
    p.hout = PHASE == 3 ? reinterpret_cast<__nv_bfloat16*>(out) : nullptr;
    p.ldh = PHASE == 3 ? ldout : 0;
    const size_t outbytes = PHASE == 1 ? (size_t)M * SIM_NSTAT * sizeof(float) : (size_t)M * ldout * sizeof(float);
    if (PHASE != 3 && (jsplit > 1 || PHASE == 1)) MMDTI_CUDA_OK(cudaMemsetAsync(out, 0, outbytes, st));      // phase 1 always accumulates with atomics
    const int smem_bytes = (int)lay.total + 1024;
    dim3 grid(stripes, jsplit, halves);
#define SIM_GO(MODE)                                                                                                    \
    do {                                                                                                                \
        auto kern = sim_tc_kernel<PHASE, BN, MODE>;                                                                     \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));             \
        kern<<<grid, NUM_THREADS, smem_bytes, st>>>(tmA, tmB, p);                                                       \
    } while (0)
    switch (aux.mode) {
        case SIM_INFONCE: SIM_GO(SIM_INFONCE); break;
        case SIM_REGRESS: SIM_GO(SIM_REGRESS); break;
        case SIM_SINGLE: SIM_GO(SIM_SINGLE); break;
        default: SIM_GO(SIM_MULTI); break;
    }
#undef SIM_GO
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

template <int PHASE>
int launch(const void* A, const void* B, int M, int N, int Dp, int D, const SimAux& aux, float* out, long long ldout, cudaStream_t st) {
    MMDTI_REQUIRE(Dp >= 64 && Dp <= 512 && Dp % 64 == 0, "sim_tc: Dp must be a multiple of 64 in [64, 512] (got %d)", Dp);
    MMDTI_REQUIRE(mmdti_aligned(A, 16) && mmdti_aligned(B, 16), "sim_tc: operands must be 16-byte aligned");
    if constexpr (PHASE != 2) {
        // 256-key tiles halve the MMA instructions per flop (the single issuing thread, not the tensor pipe, paces
        // 128-key tiles); small N keeps 128-key tiles for parallelism
        if (N >= 4096 && !getenv("MMDTI_SIM_BN128")) return launch_bn<PHASE, 256>(A, B, M, N, Dp, D, aux, out, ldout, st);
        return launch_bn<PHASE, 128>(A, B, M, N, Dp, D, aux, out, ldout, st);
    }
    if (Dp <= 128) return launch_bn<PHASE, 128>(A, B, M, N, Dp, D, aux, out, ldout, st);
    if (Dp <= 256) return launch_bn<PHASE, 64>(A, B, M, N, Dp, D, aux, out, ldout, st);
    return launch_bn<PHASE, 32>(A, B, M, N, Dp, D, aux, out, ldout, st);
}

}  // namespace

extern "C" int mmdti_sim_stats_tc(const void* A, const void* B, int M, int N, int Dp, int row_offset, int mode, float temperature,
                                  const float* y, const float* yhat, float w_thr, float e_push, const int64_t* key, int C,
                                  float coef_multi, const float* wrow, const float* wcol, float* stats, void* stream) {
    MMDTI_REQUIRE(A && B && stats && M > 0, "sim_stats_tc: bad arguments");
    const SimAux aux = sim_make_aux(mode, N, row_offset, temperature, y, yhat, w_thr, e_push, key, C, coef_multi, wrow, wcol, nullptr, nullptr);
    if (int rc = sim_check_aux(aux, 1)) return rc;
    return launch<1>(A, B, M, N, Dp, Dp, aux, stats, SIM_NSTAT, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_sim_grad_tc(const void* A, const void* B, int M, int N, int Dp, int row_offset, int mode, float temperature,
                                 const float* y, const float* yhat, float w_thr, float e_push, const int64_t* key, int C,
                                 float coef_multi, const float* wrow, const float* wcol, const float* rs_row, const float* rs_col,
                                 float* dA, int64_t lddA, void* stream) {
    MMDTI_REQUIRE(A && B && dA && M > 0 && lddA >= 1 && lddA <= Dp, "sim_grad_tc: bad arguments (need 1 <= lddA <= Dp)");
    const SimAux aux = sim_make_aux(mode, N, row_offset, temperature, y, yhat, w_thr, e_push, key, C, coef_multi, wrow, wcol, rs_row, rs_col);
    if (int rc = sim_check_aux(aux, 2)) return rc;
    return launch<2>(A, B, M, N, Dp, (int)lddA, aux, dA, lddA, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_sim_coef_tc(const void* A, const void* B, int M, int N, int Dp, int row_offset, int mode, float temperature,
                                 const float* y, const float* yhat, float w_thr, float e_push, const int64_t* key, int C,
                                 float coef_multi, const float* wrow, const float* wcol, const float* rs_row, const float* rs_col,
                                 void* H, int64_t ldh, void* stream) {
    MMDTI_REQUIRE(A && B && H && M > 0 && ldh >= N && ldh % 8 == 0 && mmdti_aligned(H, 16),
                  "sim_coef_tc: bad arguments (need ldh >= N, ldh %% 8 == 0, 16-byte aligned H)");
    const SimAux aux = sim_make_aux(mode, N, row_offset, temperature, y, yhat, w_thr, e_push, key, C, coef_multi, wrow, wcol, rs_row, rs_col);
    if (int rc = sim_check_aux(aux, 2)) return rc;
    return launch<3>(A, B, M, N, Dp, Dp, aux, static_cast<float*>(H), ldh, static_cast<cudaStream_t>(stream));
}
