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
constexpr int MAX_STAGE = 4;      // ring stages: 4 (plain epilogues) or 3 (epilogues that stream row operands through input slabs)
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
    int gelu_grad;                           // EPI_BIAS_GELU: out0 = gelu'(z) instead of z;  EPI_GELU_BWD: aux0 holds gelu'(z)
    int* sched;                              // NCTA == 1: {next item, CTAs done} of the dynamic tile scheduler (self-resetting)
};

// keep decision of flat element idx: identical to elementwise.cu (pairs of consecutive elements share one hash)
//   bits(idx) = mix32(key ^ lo32(idx >> 1) ^ mix32(hi32(idx >> 1) + 0x27d4eb2f))
// Tensors of fewer than 2^33 elements (every case of this model) have hi32 == 0: the inner hash is a constant that is
// folded into the key once per thread (ew_fold_key), one avalanche hash per element pair instead of two.
__device__ __forceinline__ uint32_t ew_fold_key(uint32_t key) { return key ^ mix32(0x27d4eb2fU); }
__device__ __forceinline__ uint32_t ew_bits_folded(uint32_t fkey, unsigned long long idx) { return mix32(fkey ^ (uint32_t)(idx >> 1)); }
__device__ __forceinline__ uint32_t ew_bits_full(uint32_t key, unsigned long long idx) {
    const unsigned long long pr = idx >> 1;
    return mix32(key ^ (uint32_t)pr ^ mix32((uint32_t)(pr >> 32) + 0x27d4eb2fU));
}

template <int NTHR> __device__ __forceinline__ void epi_bar_n() { asm volatile("bar.sync 1, %0;\n" ::"n"(NTHR) : "memory"); }

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

__device__ __forceinline__ float round_bf16(float x) { return __uint_as_float(pack_bf16(x, 0.f) << 16); }

constexpr int SLAB_BYTES = BM * 64;              // staging slab: 128 rows x 64 bytes (32 bf16 / 16 fp32 columns), 64-byte swizzle

// epilogues that read a per-row operand (residual, z, dx_add) stream it through TMA input slabs; they run a 3-stage ring
__host__ __device__ constexpr bool epi_has_input(int epi) { return epi == EPI_GELU_BWD || epi == EPI_DROPRES_LN || epi == EPI_LNBWD_DROP; }
__host__ __device__ constexpr int epi_stages(int epi) { return epi_has_input(epi) ? 3 : 4; }

// NG = column groups of the epilogue (4 warps each): 2 with 8 epilogue warps, 4 with 16
template <int NST, int NG = 2, bool HAS_IN = (NST == 3)>
struct SmemT {
    static constexpr uint32_t stages = 0;
    static constexpr int n_out = (NST == 3 && NG == 2) ? 2 : 1;        // output slabs per column group (double-buffered when there is room)
    static constexpr int n_in = HAS_IN ? 4 / NG : 0;                   // input slabs per column group: 2 (two groups) or 1 (four groups)
    static constexpr uint32_t ostage = NST * STAGE_BYTES;
    static constexpr uint32_t istage = ostage + NG * n_out * SLAB_BYTES;     // input slabs (epilogues with a per-row operand only)
    static constexpr uint32_t bars = istage + NG * n_in * SLAB_BYTES;  // mbarriers + tmem slot (256 bytes)
    static constexpr uint32_t cs = bars + 256;                        // column-sum scratch: 3 x 256 floats
    static constexpr uint32_t xchg = cs + 3 * BN * 4;                 // [2 parities][4 slots][128 rows] float2
    static constexpr uint32_t vec = xchg + 2 * 4 * BM * 8;            // per-tile column vectors: [2 tiles][bias | ln_w | ln_b][256]
    static constexpr uint32_t total = vec + 2 * 3 * BN * 4;
};

// NEW = epilogue warps: 8 (10 warps per CTA: one scheduler holds 3 of them => 168 registers per thread) or 16 (18 warps, 112
// registers: for the ALU-heavy epilogues that fit, 4 warps per scheduler keep the issue slots busy where 2 could not)
template <int AMN, int BMN, int EPI, int NCTA, int NEW = NUM_EPI_WARPS>
__global__ void __launch_bounds__(64 + NEW * 32, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO0,
               const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmI0, const __grid_constant__ CUtensorMap tmI1,
               const GemmParams p) {
    constexpr int NET = NEW * 32, NTHR = 64 + NET, NG = NEW / 4;      // epilogue threads, CTA threads, column groups
    static_assert(NEW == 8 || (NEW == 16 && NCTA == 1 && (!epi_has_input(EPI) || EPI == EPI_GELU_BWD)), "16 epilogue warps: bias / GELU epilogues only");
    constexpr int NSTAGE = NEW == 16 ? 3 : epi_stages(EPI);
    using Smem = SmemT<NSTAGE, NG, epi_has_input(EPI)>;
    auto epi_bar = [] { epi_bar_n<NET>(); };
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint64_t* full = bars;              // [MAX_STAGE]
    uint64_t* empty = bars + MAX_STAGE; // [MAX_STAGE]
    uint64_t* tfull = bars + 2 * MAX_STAGE;   // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* xbar = tempty + 2;              // [2]
    uint64_t* inbar = xbar + 2;               // [2 halves][2 buffers]
    uint64_t* sfull = inbar + 4;              // [4] tile-scheduler ring: item published
    uint64_t* sempty = sfull + 4;             // [4] item consumed by the MMA thread and every epilogue thread
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + 4);
    volatile int* s_item = reinterpret_cast<volatile int*>(tmem_slot + 1);      // [4]
    float* s_cs = reinterpret_cast<float*>(smem + Smem::cs);
    float2* s_x = reinterpret_cast<float2*>(smem + Smem::xchg);
    float* s_vec = reinterpret_cast<float*>(smem + Smem::vec);
    // the LayerNorm-backward epilogue parks two intermediates per element in TMEM: accumulator not double-buffered there
    constexpr bool DOUBLE_ACC = EPI != EPI_LNBWD_DROP;

    const uint32_t crank = NCTA > 1 ? cluster_ctarank() : 0u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], NET);
            mbar_init(&xbar[i], NCTA * NET);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&inbar[i], 1);
            mbar_init(&sfull[i], 1);
            mbar_init(&sempty[i], 1 + NET);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 3 * BN; i += NTHR) s_cs[i] = 0.f;
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

    // ---- work items.  NCTA == 1: item = (tile, k split), tile = m_blk * tiles_n + n_blk, handed out DYNAMICALLY: the
    // producer thread takes the next item from a global counter and publishes it to the MMA and epilogue warps through a
    // 4-slot ring, so CTAs that start late (SMs held by a kernel of another stream: the weight-gradient GEMMs, NCCL) simply
    // take fewer items instead of delaying the whole kernel.  NCTA == 2: the cluster walks row stripes statically, CTA rank
    // = column half.
    const int n_items = NCTA == 1 ? p.tiles_m * p.tiles_n * p.ksplit : p.tiles_m;
    const int item0 = (int)(blockIdx.x / NCTA);
    const int item_step = (int)(gridDim.x / NCTA);
    // consumer side of the scheduler ring: item of iteration `it` (or >= n_items: no more work)
    auto next_item = [&](int it) -> int {
        if (NCTA != 1) return item0 + it * item_step;
        const int sl = it & 3;
        mbar_wait_g(&sfull[sl], (it >> 2) & 1);
        const int item = s_item[sl];
        mbar_arrive(&sempty[sl]);
        return item;
    };
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
            for (int it = 0;; ++it) {
                int item;
                if (NCTA == 1) {
                    const int sl = it & 3;
                    if (it >= 4) mbar_wait_g(&sempty[sl], ((it >> 2) - 1) & 1);
                    item = atomicAdd(p.sched, 1);
                    s_item[sl] = item;
                    mbar_arrive(&sfull[sl]);                  // release: the item is visible to whoever acquires the barrier
                } else {
                    item = item0 + it * item_step;
                }
                if (item >= n_items) break;
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
            for (int it = 0;; ++it) {
                const int item = next_item(it);
                if (item >= n_items) break;
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
        // Latency discipline (two warps per scheduler, nothing else to hide behind): per-column vectors (bias, LayerNorm
        // weight / bias) are fetched BEFORE the wait for the accumulator and staged in shared memory (broadcast reads);
        // per-row operands (residual, z, dx_add) are register-prefetched one 32-column chunk ahead.
        const int ew = warp - 2;
        const int quad = warp & 3, half = ew >> 2;        // `half` = column group of this warp: 0..NG-1
        const int et = ew * 32 + lane;                    // 0..NET-1
        const int r = quad * 32 + lane;                   // accumulator row of this thread inside the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        constexpr int HC = BN / NG;                       // columns per epilogue thread
        constexpr int NCH = HC / 32;                      // 32-column chunks per thread
        constexpr bool HAS_VEC = EPI == EPI_BIAS || EPI == EPI_BIAS_GELU || EPI == EPI_DROPRES_LN || EPI == EPI_LNBWD_DROP;
        uint32_t key = 0, fkey = 0;
        if (EPI == EPI_DROPRES_LN || EPI == EPI_LNBWD_DROP) {
            key = p.thresh16 ? rng_effective_key(p.key, p.seed_off) : 0u;
            fkey = ew_fold_key(key);
        }
        const bool small_idx = (unsigned long long)p.tiles_m * BM * (unsigned long long)p.N < (1ull << 33);      // incl. the clipped rows of the last tile
        auto ew_bits = [&](uint32_t, unsigned long long idx) -> uint32_t {
            return small_idx ? ew_bits_folded(fkey, idx) : ew_bits_full(key, idx);
        };
        // Coalesced output: the 128 threads of a column half write their 64-byte row segments into a swizzled shared-memory
        // slab (conflict-free), one elected thread hands the slab to the TMA (tile store, or fp32 reduce-add for the
        // split wgrad); rows / columns beyond the matrix are clipped by the TMA.
        constexpr int NOUT = Smem::n_out;
        unsigned char* slab0 = smem + Smem::ostage + half * NOUT * SLAB_BYTES;
        const bool slab_leader = (et & 127) == 0;
        uint32_t out_cnt = 0;                             // stores issued by this column half (uniform over its threads)
        auto half_bar = [&]() { asm volatile("bar.sync %0, 128;\n" ::"r"(2 + half) : "memory"); };
        // leader: the slab the NEXT store will use (last used NOUT stores ago) has been drained by the TMA
        auto out_acquire = [&]() { if (slab_leader) bulk_wait_read<NOUT - 1>(); };
        // pre_synced: the caller ran out_acquire() and a half_bar() since the previous store
        auto slab_store = [&](const CUtensorMap* tm, int col, int row0, const uint4 (&d)[4], bool reduce, bool pre_synced = false) {
            if (!pre_synced) {
                out_acquire();
                half_bar();
            }
            unsigned char* slab = slab0 + (NOUT > 1 ? (out_cnt & 1) * SLAB_BYTES : 0);
            ++out_cnt;
            unsigned char* rowp = slab + r * 64;
            const int sw = (r >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = d[j];
            fence_proxy_async();
            half_bar();
            if (slab_leader) {
                if (reduce) tma_reduce_add_2d(tm, slab, col, row0);
                else tma_store_2d(tm, slab, col, row0);
                bulk_commit();
            }
        };
        // input slabs: the leader of a column half issues TMA loads (two buffers, one mbarrier each); every thread waits
        // for the slab and reads its own 64-byte row segment (swizzled: conflict-free).  Counters are uniform per half.
        constexpr int NIB = Smem::n_in > 0 ? Smem::n_in : 1;              // input buffers per column group
        unsigned char* islab = smem + Smem::istage + half * NIB * SLAB_BYTES;
        uint64_t* ibar = inbar + half * NIB;
        uint32_t in_issued = 0, in_read = 0;              // in_issued is only meaningful in the leader
        auto in_issue = [&](const CUtensorMap* tm, int col, int row0) {
            const int b = in_issued % NIB;
            mbar_arrive_expect_tx(&ibar[b], SLAB_BYTES);
            tma_load_2d(islab + b * SLAB_BYTES, tm, col, row0, &ibar[b]);
            ++in_issued;
        };
        auto in_take = [&](uint4 (&d)[4]) {
            const int b = in_read % NIB;
            mbar_wait_g(&ibar[b], (in_read / NIB) & 1);
            const unsigned char* rowp = islab + b * SLAB_BYTES + r * 64;
            const int sw = (r >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = *reinterpret_cast<const uint4*>(rowp + ((j ^ sw) << 4));
            ++in_read;
        };
        for (int it = 0;; ++it) {
            const int item = next_item(it);
            if (item >= n_items) break;
            int m_blk, n_blk, kc0, kc1;
            decode(item, m_blk, n_blk, kc0, kc1);
            const int buf = DOUBLE_ACC ? (it & 1) : 0;
            const int use = DOUBLE_ACC ? (it >> 1) : it;
            const int row0 = m_blk * BM;
            const long long row = (long long)row0 + r;
            const bool row_ok = row < p.M;
            const int n0 = n_blk * BN;
            const uint32_t acc_addr = lane_addr + buf * BN;
            const int cb = half * HC;                     // first column of this thread inside the tile
            float* sv = s_vec + (it & 1) * (3 * BN);      // column vectors of this tile: [0] bias, [1] ln_w, [2] ln_b
            float pv0 = 0.f, pv1 = 0.f, pv2 = 0.f;
            if (HAS_VEC && et < BN && n0 + et < p.N) {
                if (EPI != EPI_LNBWD_DROP) pv0 = __bfloat162float(p.bias[n0 + et]);
                if (EPI == EPI_LNBWD_DROP || (EPI == EPI_DROPRES_LN && p.ln_w)) pv1 = __ldg(p.ln_w + n0 + et);
                if (EPI == EPI_DROPRES_LN && p.ln_w) pv2 = __ldg(p.ln_b + n0 + et);
            }
            // per-row operand of the epilogue (residual / z / x): 64-byte slabs by TMA, two in flight per column half
            constexpr int N_IN = EPI == EPI_GELU_BWD ? NCH : 2 * NCH;       // input slabs per tile and half: 32 bf16 or 16 fp32 columns each
            constexpr int IN_W = EPI == EPI_GELU_BWD ? 32 : 16;
            if (epi_has_input(EPI) && slab_leader) {
                in_issue(&tmI0, n0 + cb, row0);
                if (NIB > 1) in_issue(&tmI0, n0 + cb + IN_W, row0);
            }
            float mu = 0.f, rs = 0.f;
            if (EPI == EPI_LNBWD_DROP && row_ok) { mu = p.mean[row]; rs = p.rstd[row]; }

            mbar_wait_g(&tfull[buf], use & 1);
            tc_fence_after();
            if (HAS_VEC) {
                if (et < BN) { sv[et] = pv0; sv[BN + et] = pv1; sv[2 * BN + et] = pv2; }
                epi_bar();
            }

            if constexpr (EPI == EPI_STORE || EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c0 = cb + ci * 32;
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
                    uint4 zq[4], uq[4];
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g8 * 8 + e]);
                        if (EPI != EPI_STORE) {
                            const float4 b0 = *reinterpret_cast<const float4*>(sv + c0 + g8 * 8), b1 = *reinterpret_cast<const float4*>(sv + c0 + g8 * 8 + 4);
                            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                        }
                        zq[g8] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        if (EPI == EPI_BIAS_GELU) {
                            if (p.gelu_grad) {
                                // u = gelu(z) and gelu'(z) of the fp32 z: the backward multiplies by the stored derivative
                                float gd[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) gelu_fast_both(f[e], f[e], gd[e]);
                                zq[g8] = make_uint4(pack_bf16(gd[0], gd[1]), pack_bf16(gd[2], gd[3]), pack_bf16(gd[4], gd[5]), pack_bf16(gd[6], gd[7]));
                            } else {
                                // u = gelu(z) of the ROUNDED z, the value a backward that re-evaluates gelu'(z) reads back
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = gelu_fast_val(round_bf16(f[e]));
                            }
                            uq[g8] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                    }
                    slab_store(&tmO0, n0 + c0, row0, zq, false);
                    if (EPI == EPI_BIAS_GELU) slab_store(&tmO1, n0 + c0, row0, uq, false);
                }
            } else if constexpr (EPI == EPI_WGRAD) {
#pragma unroll 2
                for (int c0 = cb; c0 < cb + HC; c0 += 16) {
                    uint32_t v[32];
                    tc_ld16(acc_addr + c0, v);
                    uint4 q[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) q[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    slab_store(&tmO0, n0 + c0, row0, q, p.use_atomics != 0);
                }
            } else if constexpr (EPI == EPI_GELU_BWD) {
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c0 = cb + ci * 32;
                    uint4 zq[4];
                    out_acquire();
                    in_take(zq);
                    half_bar();                                       // everybody has read the slab: refill it
                    if (slab_leader && ci + NIB < N_IN) in_issue(&tmI0, n0 + c0 + NIB * 32, row0);
                    uint32_t v[32];
                    tc_ld32(acc_addr + c0, v);
                    float cs[32];
                    uint4 dq[4];
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const float2 z01 = unpack_bf16(zq[g8].x), z23 = unpack_bf16(zq[g8].y), z45 = unpack_bf16(zq[g8].z), z67 = unpack_bf16(zq[g8].w);
                        const float z8[8] = {z01.x, z01.y, z23.x, z23.y, z45.x, z45.y, z67.x, z67.y};
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float val, grad = z8[e];
                            if (!p.gelu_grad) gelu_fast_both(z8[e], val, grad);
                            f[e] = __uint_as_float(v[g8 * 8 + e]) * grad;
                        }
                        dq[g8] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        const bool ok = row_ok && n0 + c0 + g8 * 8 < p.N;
#pragma unroll
                        for (int e = 0; e < 8; ++e) cs[g8 * 8 + e] = ok ? round_bf16(f[e]) : 0.f;      // sum what is stored
                    }
                    slab_store(&tmO0, n0 + c0, row0, dq, false, true);
                    const float tot = warp_col_reduce32(cs, lane);
                    atomicAdd(&s_cs[c0 + lane], tot);
                }
                epi_bar();
                if (et < BN) {
                    if (n0 + et < p.N && s_cs[et] != 0.f) atomicAdd(p.colsum0 + n0 + et, s_cs[et]);
                    s_cs[et] = 0.f;
                }
                epi_bar();
            } else if constexpr (EPI == EPI_DROPRES_LN) {
                const int xpar = it & 1;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
                for (int ci = 0; ci < 2 * NCH; ++ci) {
                    const int c0 = cb + ci * 16;
                    uint4 rq[4];
                    out_acquire();
                    in_take(rq);                                      // 16 fp32 of the residual row
                    half_bar();
                    if (slab_leader && ci + 2 < N_IN) in_issue(&tmI0, n0 + c0 + 32, row0);
                    uint32_t v[32];
                    tc_ld16(acc_addr + c0, v);
                    uint4 oq[4];
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        const int col = n0 + c0 + g4 * 4;
                        const float4 b4 = *reinterpret_cast<const float4*>(sv + c0 + g4 * 4);
                        float a[4] = {__uint_as_float(v[g4 * 4]) + b4.x, __uint_as_float(v[g4 * 4 + 1]) + b4.y,
                                      __uint_as_float(v[g4 * 4 + 2]) + b4.z, __uint_as_float(v[g4 * 4 + 3]) + b4.w};
                        if (p.thresh16) {
                            const long long idx = row * p.N + col;
                            const uint32_t h0 = ew_bits(key, (unsigned long long)idx), h1 = ew_bits(key, (unsigned long long)idx + 2);
                            a[0] = rng_keep(h0, 0, p.thresh16) ? a[0] * p.keep_scale : 0.f;
                            a[1] = rng_keep(h0, 1, p.thresh16) ? a[1] * p.keep_scale : 0.f;
                            a[2] = rng_keep(h1, 0, p.thresh16) ? a[2] * p.keep_scale : 0.f;
                            a[3] = rng_keep(h1, 1, p.thresh16) ? a[3] * p.keep_scale : 0.f;
                        }
                        float o[4] = {__uint_as_float(rq[g4].x) + a[0], __uint_as_float(rq[g4].y) + a[1], __uint_as_float(rq[g4].z) + a[2],
                                      __uint_as_float(rq[g4].w) + a[3]};
                        if (col >= p.N) { o[0] = o[1] = o[2] = o[3] = 0.f; }        // columns beyond the matrix do not enter the row statistics
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s1 += o[e];
                            s2 = fmaf(o[e], o[e], s2);
                            v[g4 * 4 + e] = __float_as_uint(o[e]);
                        }
                        oq[g4] = make_uint4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                    }
                    if (p.ln_w) tc_st16(acc_addr + c0, v);            // parked for the normalisation pass
                    slab_store(&tmO0, n0 + c0, row0, oq, false, true);      // xo
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
                    const float mean = S1 * inv_n;
                    const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, S2 * inv_n), 0.f) + p.eps);
                    if (row_ok && slot == 0) { p.mean[row] = mean; p.rstd[row] = rstd; }
#pragma unroll
                    for (int ci = 0; ci < NCH; ++ci) {
                        const int c0 = cb + ci * 32;
                        uint32_t v[32];
                        tc_ld32(acc_addr + c0, v);
                        uint4 yq[4];
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            const float4 w0 = *reinterpret_cast<const float4*>(sv + BN + c0 + g8 * 8), w1 = *reinterpret_cast<const float4*>(sv + BN + c0 + g8 * 8 + 4);
                            const float4 b0 = *reinterpret_cast<const float4*>(sv + 2 * BN + c0 + g8 * 8), b1 = *reinterpret_cast<const float4*>(sv + 2 * BN + c0 + g8 * 8 + 4);
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = fmaf((__uint_as_float(v[g8 * 8 + e]) - mean) * rstd, wv[e], bv[e]);
                            yq[g8] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                        slab_store(&tmO1, n0 + c0, row0, yq, false);  // LayerNorm output
                    }
                }
            } else if constexpr (EPI == EPI_LNBWD_DROP) {
                // 16-column chunks: dy, xhat and the two column-sum operands stay in registers
                constexpr int CW = 16, NCW = HC / CW;
                const int xpar = it & 1;
                const uint32_t scr_addr = lane_addr + BN;             // scratch columns [256, 512)
                const bool has_add = p.xo != nullptr;                 // dx_add present
                // column sums of 16 columns over the warp's 32 rows: lane l < 16 ends with the total of column l
                auto col_reduce16 = [&](float (&a)[16]) -> float {
#pragma unroll
                    for (int sft = 8; sft >= 1; sft >>= 1) {
                        const bool up = (lane & sft) != 0;
#pragma unroll
                        for (int i = 0; i < sft; ++i) {
                            const float keep = up ? a[i + sft] : a[i];
                            const float send = up ? a[i] : a[i + sft];
                            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                        }
                    }
                    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 16);
                };
                float c1 = 0.f, c2 = 0.f;
#pragma unroll 2
                for (int ci = 0; ci < NCW; ++ci) {
                    const int c0 = cb + ci * CW;
                    uint4 xq[4];
                    in_take(xq);                                      // 16 fp32 of the LayerNorm input row
                    half_bar();
                    if (slab_leader) {
                        if (ci + 2 < NCW) in_issue(&tmI0, n0 + c0 + 2 * CW, row0);
                        else if (has_add) in_issue(&tmI1, n0 + cb + (ci + 2 - NCW) * CW, row0);      // dx_add travels during the exchange
                    }
                    uint32_t v[32];
                    tc_ld16(acc_addr + c0, v);
                    float xh[CW], aw[CW], ab[CW];
#pragma unroll
                    for (int g4 = 0; g4 < CW / 4; ++g4) {
                        const bool ok = row_ok && n0 + c0 + g4 * 4 < p.N;
                        const float4 w4 = *reinterpret_cast<const float4*>(sv + BN + c0 + g4 * 4);
                        const float xv[4] = {__uint_as_float(xq[g4].x), __uint_as_float(xq[g4].y), __uint_as_float(xq[g4].z), __uint_as_float(xq[g4].w)};
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = g4 * 4 + e;
                            const float dy = ok ? __uint_as_float(v[i]) : 0.f;
                            const float h = ok ? (xv[e] - mu) * rs : 0.f;
                            const float g = dy * wv[e];
                            xh[i] = h;
                            aw[i] = dy * h;                           // -> d ln_w
                            ab[i] = dy;                               // -> d ln_b
                            c1 += g;
                            c2 = fmaf(g, h, c2);
                            v[i] = __float_as_uint(g);                // the accumulator slot now holds g = dy * w
                        }
                    }
                    tc_st16(acc_addr + c0, v);
                    {
                        uint32_t xu[32];
#pragma unroll
                        for (int i = 0; i < CW; ++i) xu[i] = __float_as_uint(xh[i]);
                        tc_st16(scr_addr + c0, xu);
                    }
                    const float t_w = col_reduce16(aw);
                    const float t_b = col_reduce16(ab);
                    if (lane < CW) {
                        atomicAdd(&s_cs[c0 + lane], t_w);
                        atomicAdd(&s_cs[BN + c0 + lane], t_b);
                    }
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
                for (int cj = 0; cj < NCW / 2; ++cj) {
                    uint4 daq[4];
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int ci = cj * 2 + sub;
                        const int c0 = cb + ci * CW;
                        uint4 aq[4];
                        if (has_add) {
                            out_acquire();
                            in_take(aq);
                            half_bar();
                            if (slab_leader && ci + 2 < NCW) in_issue(&tmI1, n0 + c0 + 2 * CW, row0);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) aq[j] = make_uint4(0u, 0u, 0u, 0u);
                        }
                        uint32_t gv[32], hv[32];
                        tc_ld16(acc_addr + c0, gv);
                        tc_ld16(scr_addr + c0, hv);
                        float ad[CW];
                        uint4 dxq[4];
#pragma unroll
                        for (int g4 = 0; g4 < CW / 4; ++g4) {
                            const int col = n0 + c0 + g4 * 4;
                            const bool ok = row_ok && col < p.N;
                            const float addv[4] = {__uint_as_float(aq[g4].x), __uint_as_float(aq[g4].y), __uint_as_float(aq[g4].z), __uint_as_float(aq[g4].w)};
                            float o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = g4 * 4 + e;
                                o[e] = addv[e] + rs * (__uint_as_float(gv[i]) - C1 - __uint_as_float(hv[i]) * C2);
                            }
                            dxq[g4] = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
                            if (p.thresh16) {
                                const long long idx = row * p.N + col;
                                const uint32_t h0 = ew_bits(key, (unsigned long long)idx), h1 = ew_bits(key, (unsigned long long)idx + 2);
                                o[0] = rng_keep(h0, 0, p.thresh16) ? o[0] * p.keep_scale : 0.f;
                                o[1] = rng_keep(h0, 1, p.thresh16) ? o[1] * p.keep_scale : 0.f;
                                o[2] = rng_keep(h1, 0, p.thresh16) ? o[2] * p.keep_scale : 0.f;
                                o[3] = rng_keep(h1, 1, p.thresh16) ? o[3] * p.keep_scale : 0.f;
                            }
                            const uint32_t u0 = pack_bf16(o[0], o[1]), u1 = pack_bf16(o[2], o[3]);
                            if (g4 & 1) { daq[sub * 2 + (g4 >> 1)].z = u0; daq[sub * 2 + (g4 >> 1)].w = u1; }
                            else { daq[sub * 2 + (g4 >> 1)].x = u0; daq[sub * 2 + (g4 >> 1)].y = u1; }
                            const float2 f0 = unpack_bf16(u0), f1 = unpack_bf16(u1);
                            ad[g4 * 4] = ok ? f0.x : 0.f; ad[g4 * 4 + 1] = ok ? f0.y : 0.f; ad[g4 * 4 + 2] = ok ? f1.x : 0.f; ad[g4 * 4 + 3] = ok ? f1.y : 0.f;
                        }
                        slab_store(&tmO0, n0 + c0, row0, dxq, false, has_add);        // dx (fp32, 16 columns)
                        const float t_d = col_reduce16(ad);
                        if (lane < CW) atomicAdd(&s_cs[2 * BN + c0 + lane], t_d);
                    }
                    slab_store(&tmO1, n0 + cb + cj * 2 * CW, row0, daq, false);       // da (bf16, 32 columns)
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
    if (warp >= 2 && ((threadIdx.x - 64) & 127) == 0) bulk_wait_all<0>();      // outstanding TMA stores of this column half
    __syncthreads();
    if (NCTA == 1 && threadIdx.x == 0) {
        // the last CTA to finish re-arms the scheduler counters for the next launch that uses this slot
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; }
    }
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

// output maps of the slab stores: bf16 (32 x 128) or fp32 (16 x 128) boxes, 64-byte swizzle
int make_out_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, bool fp32) {
    return make_map_2d(map, base, rows, cols, ld, fp32 ? 4 : 2, fp32 ? 16 : 32, BM, 64);
}

// Scheduler counters: a zero-initialised ring of {next item, CTAs done} pairs in device memory, one pair per launch.  The
// kernel re-arms its pair when its last CTA retires, so a CUDA-graph replay (which re-runs the same launch with the same
// pair) finds it zeroed; the ring is long enough that a pair is never shared by two launches in flight.
int* sched_slot() { return mmdti_sched_slot(); }

template <int AMN, int BMN, int EPI, int NCTA, int NEW = NUM_EPI_WARPS>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t st, const CUtensorMap* tmO0 = nullptr,
           const CUtensorMap* tmO1 = nullptr, const CUtensorMap* tmI0 = nullptr, const CUtensorMap* tmI1 = nullptr) {
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    p.kchunks = (p.K + BK - 1) / BK;
    if (p.ksplit < 1) p.ksplit = 1;
    p.ksplit = std::min(p.ksplit, p.kchunks);
    p.kchunks_per_split = (p.kchunks + p.ksplit - 1) / p.ksplit;
    p.ksplit = (p.kchunks + p.kchunks_per_split - 1) / p.kchunks_per_split;
    if (p.ksplit > 1) p.use_atomics = 1;
    auto kern = gemm_tc_kernel<AMN, BMN, EPI, NCTA, NEW>;
    const int smem_bytes = (int)SmemT<(NEW == 16 ? 3 : epi_stages(EPI)), NEW / 4, epi_has_input(EPI)>::total + 1024;
    MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int sms = num_sms();
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(64 + NEW * 32);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    if (NCTA == 1) {
        const long long items = (long long)p.tiles_m * p.tiles_n * p.ksplit;
        cfg.gridDim = dim3((unsigned)std::min<long long>(items, sms));
        cfg.numAttrs = 0;
        p.sched = sched_slot();
        MMDTI_REQUIRE(p.sched != nullptr, "gemm_tc: could not allocate the scheduler counters");
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
    const CUtensorMap& o0 = tmO0 ? *tmO0 : tmA;      // unused maps: any valid descriptor
    const CUtensorMap& o1 = tmO1 ? *tmO1 : o0;
    const CUtensorMap& i0 = tmI0 ? *tmI0 : tmA;
    const CUtensorMap& i1 = tmI1 ? *tmI1 : i0;
    MMDTI_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, o0, o1, i0, i1, p));
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
    CUtensorMap tmY;
    if (int rc = make_out_map(&tmY, Y, M, N, ldy, false)) return rc;
    if (bias) {
        MMDTI_REQUIRE(mmdti_aligned(bias, 16), "gemm_bias: bias must be 16-byte aligned");
        static const bool epi16 = getenv("MMDTI_GEMM_BIAS16") && atoi(getenv("MMDTI_GEMM_BIAS16")) != 0;
        if (epi16) return launch<0, 0, EPI_BIAS, 1, 16>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmY);
        return launch<0, 0, EPI_BIAS, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmY);
    }
    return launch<0, 0, EPI_STORE, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmY);
}

extern "C" int mmdti_gemm_bias_gelu(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* Z, int64_t ldz,
                                    void* U, int64_t ldu, int M, int N, int K, int store_grad, void* stream) {
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
    p.gelu_grad = store_grad ? 1 : 0;
    CUtensorMap tmZ, tmU;
    if (int rc = make_out_map(&tmZ, Z, M, N, ldz, false)) return rc;
    if (int rc = make_out_map(&tmU, U, M, N, ldu, false)) return rc;
    // the GELU epilogue is ALU-bound (2 MUFU + ~20 FMA per element): 16 epilogue warps (MMDTI_GEMM_EPI16=0: 8)
    static const bool epi16 = !(getenv("MMDTI_GEMM_EPI16") && atoi(getenv("MMDTI_GEMM_EPI16")) == 0);
    if (epi16) return launch<0, 0, EPI_BIAS_GELU, 1, 16>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmZ, &tmU);
    return launch<0, 0, EPI_BIAS_GELU, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmZ, &tmU);
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
    CUtensorMap tmXo, tmY, tmRes;
    if (int rc = make_out_map(&tmXo, xo, M, N, N, true)) return rc;
    if (int rc = make_out_map(&tmRes, res, M, N, N, true)) return rc;
    if (ln_w) { if (int rc = make_out_map(&tmY, Y, M, N, N, false)) return rc; }
    else tmY = tmXo;
    if (N > BN) return launch<0, 0, EPI_DROPRES_LN, 2>(tmA, tmB, p, st, &tmXo, &tmY, &tmRes);
    return launch<0, 0, EPI_DROPRES_LN, 1>(tmA, tmB, p, st, &tmXo, &tmY, &tmRes);
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
    CUtensorMap tmX;
    if (int rc = make_out_map(&tmX, dX, M, K, lddx, false)) return rc;
    return launch<0, 1, EPI_STORE, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmX);
}

extern "C" int mmdti_gemm_dgrad_gelu(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* Z, int64_t ldz, void* dZ,
                                     int64_t lddz, float* dbias, int M, int N, int K, int z_is_grad, void* stream) {
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
    p.gelu_grad = z_is_grad ? 1 : 0;
    CUtensorMap tmDZ, tmZ;
    if (int rc = make_out_map(&tmDZ, dZ, M, K, lddz, false)) return rc;
    if (int rc = make_out_map(&tmZ, Z, M, K, ldz, false)) return rc;
    static const bool epi16 = !(getenv("MMDTI_GEMM_EPI16") && atoi(getenv("MMDTI_GEMM_EPI16")) == 0);
    if (epi16) return launch<0, 1, EPI_GELU_BWD, 1, 16>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmDZ, nullptr, &tmZ);
    return launch<0, 1, EPI_GELU_BWD, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmDZ, nullptr, &tmZ);
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
    CUtensorMap tmDx, tmDa, tmX, tmAdd;
    if (int rc = make_out_map(&tmDx, dx, M, K, K, true)) return rc;
    if (int rc = make_out_map(&tmDa, da, M, K, K, false)) return rc;
    if (int rc = make_out_map(&tmX, x, M, K, K, true)) return rc;
    if (dx_add) { if (int rc = make_out_map(&tmAdd, dx_add, M, K, K, true)) return rc; }
    else tmAdd = tmX;
    if (K > BN) return launch<0, 1, EPI_LNBWD_DROP, 2>(tmA, tmB, p, st, &tmDx, &tmDa, &tmX, &tmAdd);
    return launch<0, 1, EPI_LNBWD_DROP, 1>(tmA, tmB, p, st, &tmDx, &tmDa, &tmX, &tmAdd);
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
    CUtensorMap tmW;
    if (int rc = make_out_map(&tmW, dW, N, K, lddw, true)) return rc;
    return launch<1, 1, EPI_WGRAD, 1>(tmA, tmB, p, st, &tmW);
}

extern "C" int mmdti_gemm_nn_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
                                 void* stream) {
    MMDTI_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && C && mmdti_aligned(C, 16) && ldc % 4 == 0,
                  "gemm_nn_f32: need M, N, K > 0, N, K multiples of 8, a 16-byte aligned C with ldc %% 4 == 0");
    if (int rc = check_ld(A, lda, "A")) return rc;
    if (int rc = check_ld(B, ldb, "B")) return rc;
    CUtensorMap tmA, tmB, tmC;
    if (int rc = make_map_bf16(&tmA, A, M, K, lda, BM)) return rc;           // A (M x K): reduction K, K-major
    if (int rc = make_map_bf16(&tmB, B, K, N, ldb, BK)) return rc;           // B (K x N): rows = reduction, MN-major
    if (int rc = make_out_map(&tmC, C, M, N, ldc, true)) return rc;
    GemmParams p = base_params(M, N, K);
    p.out0 = C; p.ld0 = ldc;
    return launch<0, 1, EPI_WGRAD, 1>(tmA, tmB, p, static_cast<cudaStream_t>(stream), &tmC);
}
