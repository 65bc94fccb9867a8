// Dense projections of the Uni-Core encoder layer on the 5th-generation tensor cores, with the layer's elementwise
// work fused into the GEMM epilogues (reference call sites models/transformers.py:82-91,136-139,
// layer semantics SURVEY.md Appendix A):
//
//   forward    in_proj    qkv = h1 W_in^T + b                                             (EPI_BIAS)
//              out_proj   x1  = x + dropout(o W_out^T + b);  h2 = LayerNorm2(x1)          (EPI_DROPRES_LN)
//              fc1        z   = h2 W_fc1^T + b;  u = gelu(z)                              (EPI_BIAS_GELU)
//              fc2        x2  = x1 + dropout(u W_fc2^T + b); h' = LayerNorm1_next(x2)     (EPI_DROPRES_LN)
//   backward   dgrad fc2  dz  = (df W_fc2) * gelu'(z);  db_fc1 += colsum(dz)              (EPI_GELU_BWD)
//              dgrad fc1  dx1 = dx2 + LN2'(dz W_fc1); da = dropout'(dx1); dLN2.w/b, db_out (EPI_LNBWD_DROP)
//              dgrad out  d_o = da W_out                                                   (EPI_STORE)
//              dgrad in   dx  = dx1 + LN1'(dqkv W_in); df' = dropout'(dx) ...              (EPI_LNBWD_DROP)
//              wgrad      dW  = dY^T X  (fp32, split over the token dimension)             (EPI_WGRAD)
//
// One kernel template.  CTA tile 128 (M) x 256 (N), K chunks of 64 bf16 (128-byte swizzled rows):
//   warp 0      TMA producer: A and B chunk per ring stage (4 stages x 48 KB)
//   warp 1      MMA issuer (one thread): tcgen05.mma M = 128, N = 256, K = 16, fp32 accumulators in TMEM,
//               double-buffered over tiles (2 x 256 columns) so that the epilogue of tile t overlaps the MMAs of t + 1
//   warps 2..9  epilogue: thread = one accumulator row (TMEM lane), two warps per lane quadrant splitting the
//               256 columns; tcgen05.ld 32 columns at a time, epilogue math in registers, 16-byte global stores
// Operand majors: forward GEMMs read both operands K-major; dgrad reads W MN-major (the SAME weight tensor, no
// transposed copy); wgrad reads both operands MN-major (dY^T and X^T straight from the row-major activations).
// LayerNorm epilogues need whole rows (N <= 512): a 2-CTA cluster splits the 512 columns, per-row partial sums are
// exchanged through distributed shared memory (st.shared::cluster + remote mbarrier arrive), intermediates are parked
// in TMEM (tcgen05.st) between the two passes.  Column sums (bias / LayerNorm-parameter gradients) are reduced across
// the 32 rows of a warp with a shuffle butterfly, across warps with shared-memory atomics, across CTAs with one global
// atomic per column and tile.
#include "tc_common.cuh"

#include <algorithm>

using namespace tc;

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int NSTAGE = 4;
constexpr int A_BYTES = BM * BK * 2;             // 16 KB
constexpr int B_BYTES = BN * BK * 2;             // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int MN_CHUNK_BYTES = BK * 128;         // one 64-element MN chunk of an MN-major operand tile: 64 K rows x 128 B
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + NUM_EPI_THREADS;
constexpr uint32_t TMEM_COLS = 512;

enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_DROPRES_LN = 3, EPI_GELU_BWD = 4, EPI_LNBWD_DROP = 5, EPI_WGRAD = 6 };

struct GemmParams {
    int M, N, K;                 // D is M x N, reduction length K (all in GEMM terms)
    int tiles_m, tiles_n, ksplit, kchunks, kchunks_per_split;
    // epilogue operands (meaning per EPI, see the entry points)
    void* out0; long long ld0;
    void* out1; long long ld1;
    const void* aux0; long long ldaux0;      // bf16 z (GELU backward)
    const bf16* bias;                        // (N) bf16
    const float* res; float* xo;             // fp32 residual in / out, (M, N) dense (LayerNorm epilogues: ld = N)
    const float *ln_w, *ln_b;
    float *mean, *rstd;                      // per row: written (forward) or read (backward)
    float eps;
    float *colsum0, *colsum1, *colsum2;      // (N) fp32, accumulated with atomics
    uint32_t key, thresh16; float keep_scale;
    const unsigned long long* seed_off;
    int use_atomics;                         // EPI_WGRAD with ksplit > 1
};

// keep decision of flat element idx: identical to elementwise.cu (pairs of consecutive elements share one hash)
__device__ __forceinline__ uint32_t ew_bits(uint32_t key, unsigned long long idx) {
    const unsigned long long pr = idx >> 1;
    return mix32(key ^ (uint32_t)pr ^ mix32((uint32_t)(pr >> 32) + 0x27d4eb2fU));
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(NUM_EPI_THREADS) : "memory"); }

// ---- cluster helpers (2-CTA LayerNorm epilogues)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t caddr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};\n" ::"r"(caddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t done;
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// Sum v[j] over the 32 lanes of the warp for every j: on return lane l holds the total of column l.
// Butterfly: at stride s a lane keeps the half of its values selected by bit s of its lane id and adds the partner's copy.
__device__ __forceinline__ float warp_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float keep = up ? v[i + s] : v[i];
            const float send = up ? v[i] : v[i + s];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

__device__ __forceinline__ void load_bf16x8(const bf16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store_bf16x8(bf16* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ float round_bf16(float x) { return __uint_as_float(pack_bf16(x, 0.f) << 16); }

struct Smem {
    static constexpr uint32_t stages = 0;
    static constexpr uint32_t bars = NSTAGE * STAGE_BYTES;            // full[4] empty[4] tfull[2] tempty[2] xbar[2], tmem slot
    static constexpr uint32_t cs = bars + 128;                        // column-sum scratch: 3 x 256 floats
    static constexpr uint32_t xchg = cs + 3 * BN * 4;                 // [2 parities][4 slots][128 rows] float2
    static constexpr uint32_t total = xchg + 2 * 4 * BM * 8;
};

template <int AMN, int BMN, int EPI, int NCTA>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint64_t* full = bars;              // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;    // [NSTAGE]
    uint64_t* tfull = bars + 2 * NSTAGE;      // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* xbar = tempty + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2);
    float* s_cs = reinterpret_cast<float*>(smem + Smem::cs);
    float2* s_x = reinterpret_cast<float2*>(smem + Smem::xchg);
    // the LayerNorm-backward epilogue parks two intermediates per element in TMEM: accumulator not double-buffered there
    constexpr bool DOUBLE_ACC = EPI != EPI_LNBWD_DROP;

    const uint32_t crank = NCTA > 1 ? cluster_ctarank() : 0u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], NUM_EPI_THREADS);
            mbar_init(&xbar[i], NCTA * NUM_EPI_THREADS);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 3 * BN; i += NUM_THREADS) s_cs[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (NCTA > 1) cluster_sync_all();      // the peer's barriers exist before anybody arrives on them remotely
    const uint32_t tmem_base = *tmem_slot;

    // ---- work items of this CTA.  NCTA == 1: item = (tile, k split), tile = m_blk * tiles_n + n_blk, strided over the
    // grid.  NCTA == 2: the cluster walks row stripes, CTA rank = column half.
    const int n_items = NCTA == 1 ? p.tiles_m * p.tiles_n * p.ksplit : p.tiles_m;
    const int item0 = NCTA == 1 ? (int)blockIdx.x : (int)(blockIdx.x / NCTA);
    const int item_step = NCTA == 1 ? (int)gridDim.x : (int)(gridDim.x / NCTA);
    auto decode = [&](int item, int& m_blk, int& n_blk, int& kc0, int& kc1) {
        if (NCTA == 1) {
            const int tile = item / p.ksplit, ks = item - tile * p.ksplit;
            m_blk = tile / p.tiles_n;
            n_blk = tile - m_blk * p.tiles_n;
            kc0 = ks * p.kchunks_per_split;
            kc1 = min(p.kchunks, kc0 + p.kchunks_per_split);
        } else {
            m_blk = item;
            n_blk = (int)crank;
            kc0 = 0;
            kc1 = p.kchunks;
        }
    };

    if (warp == 0) {
        // ================================================= TMA producer
        if (lane == 0) {
            int st = 0, par = 1;
            bool wrapped = false;
            for (int item = item0; item < n_items; item += item_step) {
                int m_blk, n_blk, kc0, kc1;
                decode(item, m_blk, n_blk, kc0, kc1);
                for (int kc = kc0; kc < kc1; ++kc) {
                    if (wrapped) mbar_wait_g(&empty[st], par);
                    unsigned char* sA = smem + Smem::stages + (size_t)st * STAGE_BYTES;
                    unsigned char* sB = sA + A_BYTES;
                    mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
                    if (AMN) {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * MN_CHUNK_BYTES, &tmA, m_blk * BM + j * 64, kc * BK, &full[st]);
                    } else {
                        tma_load_2d(sA, &tmA, kc * BK, m_blk * BM, &full[st]);
                    }
                    if (BMN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * MN_CHUNK_BYTES, &tmB, n_blk * BN + j * 64, kc * BK, &full[st]);
                    } else {
                        tma_load_2d(sB, &tmB, kc * BK, n_blk * BN, &full[st]);
                    }
                    if (++st == NSTAGE) { st = 0; par ^= 1; wrapped = true; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_mn(BM, BN, AMN, BMN);
            const uint32_t s_addr = smem_u32(smem + Smem::stages);
            const uint64_t adesc0 = AMN ? smem_desc(s_addr, MN_CHUNK_BYTES, 1024) : smem_desc(s_addr, 16, 1024);
            const uint64_t bdesc0 = BMN ? smem_desc(s_addr + A_BYTES, MN_CHUNK_BYTES, 1024) : smem_desc(s_addr + A_BYTES, 16, 1024);
            constexpr uint32_t a_kstep = AMN ? (16 * 128) >> 4 : 32 >> 4;      // descriptor advance per K = 16
            constexpr uint32_t b_kstep = BMN ? (16 * 128) >> 4 : 32 >> 4;
            int st = 0;
            uint32_t par = 0;
            int it = 0;
            for (int item = item0; item < n_items; item += item_step, ++it) {
                int m_blk, n_blk, kc0, kc1;
                decode(item, m_blk, n_blk, kc0, kc1);
                const int buf = DOUBLE_ACC ? (it & 1) : 0;
                const int use = DOUBLE_ACC ? (it >> 1) : it;                    // how often this buffer was used before
                if (use > 0) mbar_wait_g(&tempty[buf], (use - 1) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int kc = kc0; kc < kc1; ++kc) {
                    mbar_wait_g(&full[st], par);
                    tc_fence_after();
                    const uint64_t ad = adesc0 + (uint64_t)(st * (STAGE_BYTES >> 4));
                    const uint64_t bd = bdesc0 + (uint64_t)(st * (STAGE_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma(d_tmem, ad + (uint64_t)(k * a_kstep), bd + (uint64_t)(k * b_kstep), idesc, (kc > kc0 || k > 0) ? 1u : 0u);
                    tc_commit(&empty[st]);
                    if (++st == NSTAGE) { st = 0; par ^= 1u; }
                }
                tc_commit(&tfull[buf]);
            }
        }
    } else {
        // ================================================= epilogue warps
        const int ew = warp - 2;
        const int quad = warp & 3, half = ew >> 2;
        const int et = ew * 32 + lane;                    // 0..255
        const int r = quad * 32 + lane;                   // accumulator row of this thread inside the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        constexpr int HC = BN / 2;                        // columns per epilogue thread
        uint32_t key = 0;
        if (EPI == EPI_DROPRES_LN || EPI == EPI_LNBWD_DROP) key = p.thresh16 ? rng_effective_key(p.key, p.seed_off) : 0u;
        int it = 0;
        for (int item = item0; item < n_items; item += item_step, ++it) {
            int m_blk, n_blk, kc0, kc1;
            decode(item, m_blk, n_blk, kc0, kc1);
            const int buf = DOUBLE_ACC ? (it & 1) : 0;
            const int use = DOUBLE_ACC ? (it >> 1) : it;
            const long long row = (long long)m_blk * BM + r;
            const bool row_ok = row < p.M;
            const int n0 = n_blk * BN;
            const uint32_t acc_addr = lane_addr + buf * BN;
            mbar_wait_g(&tfull[buf], use & 1);
            tc_fence_after();

            if constexpr (EPI == EPI_STORE || EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
#pragma unroll 1
                for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
                    if (row_ok) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const int col = n0 + c0 + g8 * 8;
                            if (col < p.N) {
                                float f[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g8 * 8 + e]);
                                if (EPI != EPI_STORE) {
                                    float b8[8];
                                    load_bf16x8(p.bias + col, b8);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] += b8[e];
                                }
                                store_bf16x8(static_cast<bf16*>(p.out0) + row * p.ld0 + col, f);
                                if (EPI == EPI_BIAS_GELU) {
                                    // u = gelu(z) of the ROUNDED z, the value the backward reads back
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] = gelu_fast_val(round_bf16(f[e]));
                                    store_bf16x8(static_cast<bf16*>(p.out1) + row * p.ld1 + col, f);
                                }
                            }
                        }
                    }
                }
            } else if constexpr (EPI == EPI_WGRAD) {
#pragma unroll 1
                for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
                    if (row_ok) {
                        float* dst = static_cast<float*>(p.out0) + row * p.ld0 + n0 + c0;
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if (n0 + c0 + g4 * 4 < p.N) {
                                const float a = __uint_as_float(v[g4 * 4]), b = __uint_as_float(v[g4 * 4 + 1]);
                                const float c = __uint_as_float(v[g4 * 4 + 2]), d = __uint_as_float(v[g4 * 4 + 3]);
                                if (p.use_atomics) red_add_v4(dst + g4 * 4, a, b, c, d);
                                else *reinterpret_cast<float4*>(dst + g4 * 4) = make_float4(a, b, c, d);
                            }
                        }
                    }
                }
            } else if constexpr (EPI == EPI_GELU_BWD) {
#pragma unroll 1
                for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
                    float cs[32];
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const int col = n0 + c0 + g8 * 8;
                        float f[8];
                        if (row_ok && col < p.N) {
                            float z8[8];
                            load_bf16x8(static_cast<const bf16*>(p.aux0) + row * p.ldaux0 + col, z8);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float val, grad;
                                gelu_fast_both(z8[e], val, grad);
                                f[e] = __uint_as_float(v[g8 * 8 + e]) * grad;
                            }
                            store_bf16x8(static_cast<bf16*>(p.out0) + row * p.ld0 + col, f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) cs[g8 * 8 + e] = round_bf16(f[e]);      // sum what was stored
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) cs[g8 * 8 + e] = 0.f;
                        }
                    }
                    const float tot = warp_col_reduce32(cs, lane);
                    atomicAdd(&s_cs[c0 + lane], tot);
                }
                epi_bar();
                if (n0 + et < p.N && s_cs[et] != 0.f) atomicAdd(p.colsum0 + n0 + et, s_cs[et]);
                s_cs[et] = 0.f;
                epi_bar();
            } else if constexpr (EPI == EPI_DROPRES_LN) {
                const int xpar = it & 1;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
                for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const int col = n0 + c0 + g4 * 4;
                        float o[4] = {0.f, 0.f, 0.f, 0.f};
                        if (row_ok && col < p.N) {
                            const uint2 bu = __ldg(reinterpret_cast<const uint2*>(p.bias + col));
                            const float2 b01 = unpack_bf16(bu.x), b23 = unpack_bf16(bu.y);
                            float a[4] = {__uint_as_float(v[g4 * 4]) + b01.x, __uint_as_float(v[g4 * 4 + 1]) + b01.y,
                                          __uint_as_float(v[g4 * 4 + 2]) + b23.x, __uint_as_float(v[g4 * 4 + 3]) + b23.y};
                            const long long idx = row * p.N + col;
                            if (p.thresh16) {
                                const uint32_t h0 = ew_bits(key, (unsigned long long)idx), h1 = ew_bits(key, (unsigned long long)idx + 2);
                                a[0] = rng_keep(h0, 0, p.thresh16) ? a[0] * p.keep_scale : 0.f;
                                a[1] = rng_keep(h0, 1, p.thresh16) ? a[1] * p.keep_scale : 0.f;
                                a[2] = rng_keep(h1, 0, p.thresh16) ? a[2] * p.keep_scale : 0.f;
                                a[3] = rng_keep(h1, 1, p.thresh16) ? a[3] * p.keep_scale : 0.f;
                            }
                            const float4 x4 = *reinterpret_cast<const float4*>(p.res + idx);
                            o[0] = x4.x + a[0]; o[1] = x4.y + a[1]; o[2] = x4.z + a[2]; o[3] = x4.w + a[3];
                            *reinterpret_cast<float4*>(p.xo + idx) = make_float4(o[0], o[1], o[2], o[3]);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s1 += o[e];
                            s2 = fmaf(o[e], o[e], s2);
                            v[g4 * 4 + e] = __float_as_uint(o[e]);
                        }
                    }
                    if (p.ln_w) tc_st32(acc_addr + c0, v);            // parked for the normalisation pass
                }
                if (p.ln_w) {
                    // per-row sums: 2 column halves x NCTA column blocks -> every CTA of the cluster gets all partials
                    const int slot = (int)crank * 2 + half;
                    float2* mine = s_x + ((size_t)xpar * 4 + slot) * BM + r;
                    *mine = make_float2(s1, s2);
                    if (NCTA > 1) {
                        const uint32_t peer = crank ^ 1u;
                        st_cluster_f2(mapa_u32(smem_u32(mine), peer), s1, s2);
                        mbar_arrive_remote(mapa_u32(smem_u32(&xbar[xpar]), peer));
                    }
                    mbar_arrive(&xbar[xpar]);
                    mbar_wait_cluster(&xbar[xpar], (it >> 1) & 1);
                    float S1 = 0.f, S2 = 0.f;
#pragma unroll
                    for (int sl = 0; sl < 2 * NCTA; ++sl) {
                        const float2 t = s_x[((size_t)xpar * 4 + sl) * BM + r];
                        S1 += t.x;
                        S2 += t.y;
                    }
                    const float inv_n = 1.f / (float)p.N;
                    const float mu = S1 * inv_n;
                    const float rs = rsqrtf(fmaxf(fmaf(-mu, mu, S2 * inv_n), 0.f) + p.eps);
                    if (row_ok && slot == 0) { p.mean[row] = mu; p.rstd[row] = rs; }
#pragma unroll 1
                    for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                        uint32_t v[32];
                        tc_ld32(acc_addr + c0, v);
                        if (row_ok) {
#pragma unroll
                            for (int g8 = 0; g8 < 4; ++g8) {
                                const int col = n0 + c0 + g8 * 8;
                                if (col < p.N) {
                                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ln_w + col)), w1 = __ldg(reinterpret_cast<const float4*>(p.ln_w + col + 4));
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + 4));
                                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                                    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                                    float f[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] = fmaf((__uint_as_float(v[g8 * 8 + e]) - mu) * rs, wv[e], bv[e]);
                                    store_bf16x8(static_cast<bf16*>(p.out0) + row * p.ld0 + col, f);
                                }
                            }
                        }
                    }
                }
            } else if constexpr (EPI == EPI_LNBWD_DROP) {
                const int xpar = it & 1;
                const uint32_t scr_addr = lane_addr + BN;             // scratch columns [256, 512)
                float mu = 0.f, rs = 0.f;
                if (row_ok) { mu = p.mean[row]; rs = p.rstd[row]; }
                float c1 = 0.f, c2 = 0.f;
#pragma unroll 1
                for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
                    float xh[32], aw[32], ab[32];
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const int col = n0 + c0 + g4 * 4;
                        if (row_ok && col < p.N) {
                            const float4 x4 = *reinterpret_cast<const float4*>(p.res + row * p.N + col);
                            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ln_w + col));
                            const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = g4 * 4 + e;
                                const float dy = __uint_as_float(v[i]);
                                const float h = (xv[e] - mu) * rs;
                                const float g = dy * wv[e];
                                xh[i] = h;
                                aw[i] = dy * h;                       // -> d ln_w
                                ab[i] = dy;                           // -> d ln_b
                                c1 += g;
                                c2 = fmaf(g, h, c2);
                                v[i] = __float_as_uint(g);            // the accumulator slot now holds g = dy * w
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = g4 * 4 + e;
                                xh[i] = 0.f; aw[i] = 0.f; ab[i] = 0.f;
                                v[i] = 0u;
                            }
                        }
                    }
                    tc_st32(acc_addr + c0, v);
                    {
                        uint32_t xu[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) xu[i] = __float_as_uint(xh[i]);
                        tc_st32(scr_addr + c0, xu);
                    }
                    const float t_w = warp_col_reduce32(aw, lane);
                    const float t_b = warp_col_reduce32(ab, lane);
                    atomicAdd(&s_cs[c0 + lane], t_w);
                    atomicAdd(&s_cs[BN + c0 + lane], t_b);
                }
                // per-row c1, c2 across the column halves and the cluster
                const int slot = (int)crank * 2 + half;
                float2* mine = s_x + ((size_t)xpar * 4 + slot) * BM + r;
                *mine = make_float2(c1, c2);
                if (NCTA > 1) {
                    const uint32_t peer = crank ^ 1u;
                    st_cluster_f2(mapa_u32(smem_u32(mine), peer), c1, c2);
                    mbar_arrive_remote(mapa_u32(smem_u32(&xbar[xpar]), peer));
                }
                mbar_arrive(&xbar[xpar]);
                mbar_wait_cluster(&xbar[xpar], (it >> 1) & 1);
                float C1 = 0.f, C2 = 0.f;
#pragma unroll
                for (int sl = 0; sl < 2 * NCTA; ++sl) {
                    const float2 t = s_x[((size_t)xpar * 4 + sl) * BM + r];
                    C1 += t.x;
                    C2 += t.y;
                }
                const float inv_n = 1.f / (float)p.N;
                C1 *= inv_n;
                C2 *= inv_n;
#pragma unroll 1
                for (int c0 = half * HC; c0 < (half + 1) * HC; c0 += 32) {
                    uint32_t gv[32], hv[32];
                    tc_ld32(acc_addr + c0, gv);
                    tc_ld32(scr_addr + c0, hv);
                    float ad[32];
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const int col = n0 + c0 + g4 * 4;
                        if (row_ok && col < p.N) {
                            const long long idx = row * p.N + col;
                            float o[4];
                            float4 add4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.xo) add4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.xo) + idx);      // dx_add
                            const float addv[4] = {add4.x, add4.y, add4.z, add4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = g4 * 4 + e;
                                o[e] = addv[e] + rs * (__uint_as_float(gv[i]) - C1 - __uint_as_float(hv[i]) * C2);
                            }
                            *reinterpret_cast<float4*>(static_cast<float*>(p.out1) + idx) = make_float4(o[0], o[1], o[2], o[3]);      // dx
                            if (p.thresh16) {
                                const uint32_t h0 = ew_bits(key, (unsigned long long)idx), h1 = ew_bits(key, (unsigned long long)idx + 2);
                                o[0] = rng_keep(h0, 0, p.thresh16) ? o[0] * p.keep_scale : 0.f;
                                o[1] = rng_keep(h0, 1, p.thresh16) ? o[1] * p.keep_scale : 0.f;
                                o[2] = rng_keep(h1, 0, p.thresh16) ? o[2] * p.keep_scale : 0.f;
                                o[3] = rng_keep(h1, 1, p.thresh16) ? o[3] * p.keep_scale : 0.f;
                            }
                            const uint32_t u0 = pack_bf16(o[0], o[1]), u1 = pack_bf16(o[2], o[3]);
                            *reinterpret_cast<uint2*>(static_cast<bf16*>(p.out0) + row * p.ld0 + col) = make_uint2(u0, u1);      // da
                            const float2 f0 = unpack_bf16(u0), f1 = unpack_bf16(u1);
                            ad[g4 * 4] = f0.x; ad[g4 * 4 + 1] = f0.y; ad[g4 * 4 + 2] = f1.x; ad[g4 * 4 + 3] = f1.y;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) ad[g4 * 4 + e] = 0.f;
                        }
                    }
                    const float t_d = warp_col_reduce32(ad, lane);
                    atomicAdd(&s_cs[2 * BN + c0 + lane], t_d);
                }
                epi_bar();
                if (n0 + et < p.N) {
                    atomicAdd(p.colsum0 + n0 + et, s_cs[et]);                   // d ln_w
                    atomicAdd(p.colsum1 + n0 + et, s_cs[BN + et]);              // d ln_b
                    atomicAdd(p.colsum2 + n0 + et, s_cs[2 * BN + et]);          // d bias of the linear layer under the dropout
                }
                s_cs[et] = 0.f; s_cs[BN + et] = 0.f; s_cs[2 * BN + et] = 0.f;
                epi_bar();
            }
            tc_fence_before();
            mbar_arrive(&tempty[buf]);
        }
    }
    __syncthreads();
    __syncwarp();
    if (NCTA > 1) cluster_sync_all();          // nobody leaves while the peer may still address this CTA's shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

int num_sms() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

inline void drop_params(float p, uint32_t& thresh16, float& keep_scale) {
    double t = floor((double)p * 65536.0 + 0.5);
    if (t < 0) t = 0;
    if (t > 65535) t = 65535;
    thresh16 = (uint32_t)t;
    keep_scale = (float)(65536.0 / (65536.0 - t));
}

// rng key of a flat-tensor dropout stream: the same derivation as elementwise.cu (mmdti_dropres_layernorm_fwd etc.), so
// that fused and unfused kernels draw the same mask for a given seed
inline uint32_t ew_key(uint64_t seed) { return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x165667B1U)); }

template <int AMN, int BMN, int EPI, int NCTA>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t st) {
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    p.kchunks = (p.K + BK - 1) / BK;
    if (p.ksplit < 1) p.ksplit = 1;
    p.ksplit = std::min(p.ksplit, p.kchunks);
    p.kchunks_per_split = (p.kchunks + p.ksplit - 1) / p.ksplit;
    p.ksplit = (p.kchunks + p.kchunks_per_split - 1) / p.kchunks_per_split;
    if (p.ksplit > 1) p.use_atomics = 1;
    auto kern = gemm_tc_kernel<AMN, BMN, EPI, NCTA>;
    const int smem_bytes = (int)Smem::total + 1024;
    MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int sms = num_sms();
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    if (NCTA == 1) {
        const long long items = (long long)p.tiles_m * p.tiles_n * p.ksplit;
        cfg.gridDim = dim3((unsigned)std::min<long long>(items, sms));
        cfg.numAttrs = 0;
    } else {
        MMDTI_REQUIRE(p.tiles_n == NCTA, "gemm_tc: the LayerNorm epilogue needs N = %d..%d columns for a %d-CTA cluster (N = %d)",
                      (NCTA - 1) * BN + 1, NCTA * BN, NCTA, p.N);
        const int clusters = std::min(p.tiles_m, sms / NCTA);
        cfg.gridDim = dim3((unsigned)(clusters * NCTA));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NCTA;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    MMDTI_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
    return MMDTI_OK;
}

int check_ld(const void* ptr, long long ld, const char* what) {
    MMDTI_REQUIRE(ptr && mmdti_aligned(ptr, 16) && ld > 0 && (ld * 2) % 16 == 0, "gemm_tc: %s must be non-null, 16-byte aligned, with a row stride that is a multiple of 8 elements", what);
    return MMDTI_OK;
}

GemmParams base_params(int M, int N, int K) {
    GemmParams p = {};
    p.M = M; p.N = N; p.K = K;
    p.ksplit = 1;
    p.eps = 1e-5f;
    return p;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int mmdti_gemm_bias(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* Y, int64_t ldy, int M,
                               int N, int K, void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "gemm_bias: need M, N, K > 0 and N, K multiples of 8");
    if (int rc = check_ld(X, ldx, "X")) return rc;
    if (int rc = check_ld(W, ldw, "W")) return rc;
    if (int rc = check_ld(Y, ldy, "Y")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, X, M, K, ldx, BM)) return rc;
    if (int rc = make_map_bf16(&tmB, W, N, K, ldw, BN)) return rc;
    GemmParams p = base_params(M, N, K);
    p.out0 = Y; p.ld0 = ldy; p.bias = static_cast<const bf16*>(bias);
    if (bias) {
        MMDTI_REQUIRE(mmdti_aligned(bias, 16), "gemm_bias: bias must be 16-byte aligned");
        return launch<0, 0, EPI_BIAS, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
    }
    return launch<0, 0, EPI_STORE, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_gemm_bias_gelu(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* Z, int64_t ldz,
                                    void* U, int64_t ldu, int M, int N, int K, void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && bias && mmdti_aligned(bias, 16),
                  "gemm_bias_gelu: need M, N, K > 0, N, K multiples of 8 and a 16-byte aligned bias");
    if (int rc = check_ld(X, ldx, "X")) return rc;
    if (int rc = check_ld(W, ldw, "W")) return rc;
    if (int rc = check_ld(Z, ldz, "Z")) return rc;
    if (int rc = check_ld(U, ldu, "U")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, X, M, K, ldx, BM)) return rc;
    if (int rc = make_map_bf16(&tmB, W, N, K, ldw, BN)) return rc;
    GemmParams p = base_params(M, N, K);
    p.out0 = Z; p.ld0 = ldz; p.out1 = U; p.ld1 = ldu; p.bias = static_cast<const bf16*>(bias);
    return launch<0, 0, EPI_BIAS_GELU, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_gemm_dropres_ln(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, const float* res,
                                     float* xo, const float* ln_w, const float* ln_b, void* Y, float* mean, float* rstd, int M, int N,
                                     int K, float eps, float dropout_p, uint64_t seed, void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && N <= 2 * BN, "gemm_dropres_ln: need N %% 8 == 0, K %% 8 == 0 and N <= 512 (N = %d)", N);
    MMDTI_REQUIRE(bias && res && xo && mmdti_aligned(bias, 16) && mmdti_aligned(res, 16) && mmdti_aligned(xo, 16), "gemm_dropres_ln: bias / res / xo must be 16-byte aligned");
    MMDTI_REQUIRE(!ln_w || (ln_b && Y && mean && rstd && mmdti_aligned(ln_w, 16) && mmdti_aligned(ln_b, 16) && mmdti_aligned(Y, 16)),
                  "gemm_dropres_ln: the LayerNorm outputs need ln_b, Y, mean and rstd");
    MMDTI_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gemm_dropres_ln: dropout_p out of range");
    if (int rc = check_ld(X, ldx, "X")) return rc;
    if (int rc = check_ld(W, ldw, "W")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, X, M, K, ldx, BM)) return rc;
    if (int rc = make_map_bf16(&tmB, W, N, K, ldw, BN)) return rc;
    GemmParams p = base_params(M, N, K);
    p.bias = static_cast<const bf16*>(bias); p.res = res; p.xo = xo; p.ln_w = ln_w; p.ln_b = ln_b; p.out0 = Y; p.ld0 = N;
    p.mean = mean; p.rstd = rstd; p.eps = eps;
    drop_params(dropout_p, p.thresh16, p.keep_scale);
    p.key = ew_key(seed);
    p.seed_off = mmdti_seed_offset_ptr();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N > BN) return launch<0, 0, EPI_DROPRES_LN, 2>(tmA, tmB, p, st);
    return launch<0, 0, EPI_DROPRES_LN, 1>(tmA, tmB, p, st);
}

extern "C" int mmdti_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dX, int64_t lddx, int M, int N, int K,
                                void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0, "gemm_dgrad: need M, N, K > 0 and N, K multiples of 8");
    if (int rc = check_ld(dY, lddy, "dY")) return rc;
    if (int rc = check_ld(W, ldw, "W")) return rc;
    if (int rc = check_ld(dX, lddx, "dX")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, dY, M, N, lddy, BM)) return rc;         // A = dY: (M x N), reduction N, K-major
    if (int rc = make_map_bf16(&tmB, W, N, K, ldw, BK)) return rc;           // B = W (N x K): rows = reduction, MN-major
    GemmParams p = base_params(M, K, N);
    p.out0 = dX; p.ld0 = lddx;
    return launch<0, 1, EPI_STORE, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_gemm_dgrad_gelu(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* Z, int64_t ldz, void* dZ,
                                     int64_t lddz, float* dbias, int M, int N, int K, void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && dbias, "gemm_dgrad_gelu: need M, N, K > 0, N, K multiples of 8, dbias");
    if (int rc = check_ld(dY, lddy, "dY")) return rc;
    if (int rc = check_ld(W, ldw, "W")) return rc;
    if (int rc = check_ld(Z, ldz, "Z")) return rc;
    if (int rc = check_ld(dZ, lddz, "dZ")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, dY, M, N, lddy, BM)) return rc;
    if (int rc = make_map_bf16(&tmB, W, N, K, ldw, BK)) return rc;
    GemmParams p = base_params(M, K, N);
    p.out0 = dZ; p.ld0 = lddz; p.aux0 = Z; p.ldaux0 = ldz; p.colsum0 = dbias;
    return launch<0, 1, EPI_GELU_BWD, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_gemm_dgrad_lnbwd(const void* dY, int64_t lddy, const void* W, int64_t ldw, const float* x, const float* mean,
                                      const float* rstd, const float* ln_w, const float* dx_add, float* dx, float* dw, float* db,
                                      void* da, float* dbias, int M, int N, int K, float dropout_p, uint64_t seed, void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && K <= 2 * BN, "gemm_dgrad_lnbwd: need N %% 8 == 0, K %% 8 == 0 and K <= 512 (K = %d)", K);
    MMDTI_REQUIRE(x && mean && rstd && ln_w && dx && dw && db && da && dbias && mmdti_aligned(x, 16) && mmdti_aligned(ln_w, 16) &&
                      mmdti_aligned(dx, 16) && mmdti_aligned(da, 16) && mmdti_aligned(dx_add, 16),
                  "gemm_dgrad_lnbwd: null or misaligned buffer");
    MMDTI_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "gemm_dgrad_lnbwd: dropout_p out of range");
    if (int rc = check_ld(dY, lddy, "dY")) return rc;
    if (int rc = check_ld(W, ldw, "W")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, dY, M, N, lddy, BM)) return rc;
    if (int rc = make_map_bf16(&tmB, W, N, K, ldw, BK)) return rc;
    GemmParams p = base_params(M, K, N);
    p.res = x; p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd); p.ln_w = ln_w;
    p.xo = const_cast<float*>(dx_add);           // read-only here
    p.out1 = dx; p.out0 = da; p.ld0 = K;
    p.colsum0 = dw; p.colsum1 = db; p.colsum2 = dbias;
    drop_params(dropout_p, p.thresh16, p.keep_scale);
    p.key = ew_key(seed);
    p.seed_off = mmdti_seed_offset_ptr();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (K > BN) return launch<0, 1, EPI_LNBWD_DROP, 2>(tmA, tmB, p, st);
    return launch<0, 1, EPI_LNBWD_DROP, 1>(tmA, tmB, p, st);
}

extern "C" int mmdti_gemm_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw, int M, int N, int K,
                                int accumulate, void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && dW && mmdti_aligned(dW, 16) && lddw % 4 == 0,
                  "gemm_wgrad: need M, N, K > 0, N, K multiples of 8, a 16-byte aligned dW with lddw %% 4 == 0");
    if (int rc = check_ld(dY, lddy, "dY")) return rc;
    if (int rc = check_ld(X, ldx, "X")) return rc;
    CUtensorMap tmA, tmB;
    if (int rc = make_map_bf16(&tmA, dY, M, N, lddy, BK)) return rc;         // A = dY^T: rows of dY = reduction, MN-major
    if (int rc = make_map_bf16(&tmB, X, M, K, ldx, BK)) return rc;           // B = X^T : rows of X  = reduction, MN-major
    GemmParams p = base_params(N, K, M);
    p.out0 = dW; p.ld0 = lddw;
    // split the token dimension so that tiles x splits fills the SMs; partial products meet in dW through 16-byte
    // fp32 reductions (dW zeroed first unless the caller accumulates)
    const int tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
    const int kchunks = (M + BK - 1) / BK;
    int ks = std::max(1, std::min(kchunks, num_sms() / std::max(tiles, 1)));
    const int per = (kchunks + ks - 1) / ks;
    ks = (kchunks + per - 1) / per;                 // what launch() will settle on
    p.ksplit = ks;
    p.use_atomics = (ks > 1 || accumulate) ? 1 : 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ks > 1 && !accumulate) {
        if (lddw == K) MMDTI_CUDA_OK(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st));
        else MMDTI_CUDA_OK(cudaMemset2DAsync(dW, (size_t)lddw * sizeof(float), 0, (size_t)K * sizeof(float), (size_t)N, st));
    }
    return launch<1, 1, EPI_WGRAD, 1>(tmA, tmB, p, st);
}
