// K1 forward on the 5th-generation tensor cores: Gaussian pair-distance basis -> 2-layer MLP -> per-head pair bias in the
// padded (B,H,L,Lp) layout, bf16 operands / fp32 accumulation in TMEM.
//
// Reference: GaussianLayer.forward + gaussian() (models/mm_model.py:211-224,254-269), NonLinearHead.forward as gbf_proj
// (:117-128), permute/contiguous (:553-556), key-padding merge (models/transformers.py:122-132):
//   u = mul[et]*dist + bias[et];  g_k = exp(-0.5((u-mu_k)/sigma_k)^2) / (sqrt(2*3.14159) sigma_k),  sigma_k = |std_k| + 1e-5
//   out[b,h,i,j] = (W2 gelu(W1 g + b1) + b2)_h,  -inf at padded keys and in the layout's padding columns
//
// Work item = 128 consecutive positions q = i*Lp + j' of one molecule's PADDED (L x Lp) tile, so that for every head the
// item's outputs are 128 contiguous elements: the (64 heads x 128 positions) result leaves by ONE 3-D TMA store
// (dims q, h, b), clipped at the molecule's end.  Positions with j' >= L are the layout's -inf padding.
//   stage 1  all threads: gather u, build the basis row (128 kernels) straight into the swizzled K-major A tile
//            (3 instructions per value: the normalisation constant rides in the exponent)
//   MMA 1    one thread: z (128 x 128, TMEM) = G . W1^T                     8 x tcgen05.mma (M 128, N 128, K 16)
//   stage 2  all threads: tcgen05.ld z, + b1, exact-erf GELU, bf16 -> the same A tile
//   MMA 2    one thread: o (128 x 64, TMEM) = H . W2^T                      8 x tcgen05.mma (M 128, N 64, K 16)
//   stage 3  all threads: tcgen05.ld o, + b2, -inf masks, transposed into the (head, position) store tile; TMA store
// W1 / W2 stay resident in shared memory (bf16, swizzled) for the whole kernel; two CTAs per SM interleave their
// MUFU-bound stages with each other's MMA waits (the kernel is bound by the 2 x 71 M exp / erf evaluations, not by HBM).
#include "tc_common.cuh"

#include <algorithm>

using namespace tc;

namespace {

constexpr int KB = 128;          // Gaussian kernels
constexpr int NH = 64;           // heads
constexpr int TQ = 128;          // positions per work item (= UMMA M)
constexpr int NT = 512;          // threads: 4 per position (quarters of the kernels / columns / heads)
constexpr int NPART = NT / TQ;
constexpr int CHUNK_A = TQ * 128;        // one 64-wide K chunk of a 128-row operand tile: 16 KB
constexpr int CHUNK_W2 = NH * 128;       // 64 rows: 8 KB
constexpr uint32_t TMEM_COLS = 256;      // z: columns [0,128), o: [128,192)

struct K1Params {
    const float* dist;
    const long long* et;
    const float *means, *stds, *mul, *bias, *w1, *b1, *w2, *b2;
    const unsigned char* key_pad;
    int B, L, Lp, E, tiles_per_mol;
};

template <typename TP>
struct K1Smem {
    static constexpr uint32_t w1 = 0;                                  // 2 chunks x 16 KB
    static constexpr uint32_t w2 = w1 + 2 * CHUNK_A;                   // 2 chunks x 8 KB
    static constexpr uint32_t a = w2 + 2 * CHUNK_W2;                   // basis / hidden tile: 2 chunks x 16 KB
    static constexpr uint32_t ot = a + 2 * CHUNK_A;                    // [64 heads][128 positions] TP
    static constexpr uint32_t vec = ot + NH * TQ * sizeof(TP);         // mu', is', lc (3 x 128), b1 (128), b2 (64) floats
    static constexpr uint32_t tab = vec + (4 * KB + NH) * 4;           // mul[E], bias[E]
    // + 2 * E floats + barriers: computed at launch
};

// write 8 consecutive bf16 (16 bytes) of row `row`, K index k0 (multiple of 8) into a K-major 128-byte-swizzled operand
// tile made of 64-wide K chunks of `chunk_bytes` each
__device__ __forceinline__ void st_swz(unsigned char* tile, int chunk_bytes, int row, int k0, uint4 v) {
    unsigned char* p = tile + (k0 >> 6) * chunk_bytes + row * 128 + ((((k0 & 63) >> 3) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(p) = v;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <typename TP>
__global__ void __launch_bounds__(NT, 2) pair_bias_fwd_tc5_kernel(const __grid_constant__ CUtensorMap tmOut, const K1Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    using S = K1Smem<TP>;
    unsigned char* sW1 = smem + S::w1;
    unsigned char* sW2 = smem + S::w2;
    unsigned char* sA = smem + S::a;
    TP* sOt = reinterpret_cast<TP*>(smem + S::ot);
    float* s_mu = reinterpret_cast<float*>(smem + S::vec);      // mu_k * is'_k
    float* s_is = s_mu + KB;                                      // is'_k = sqrt(0.5 log2 e) / sigma_k
    float* s_lc = s_is + KB;                                      // log2(1 / (sqrt(2*3.14159) sigma_k))
    float* s_b1 = s_lc + KB;
    float* s_b2 = s_b1 + KB;
    float* s_mul = reinterpret_cast<float*>(smem + S::tab);
    float* s_bias = s_mul + p.E;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_bias + p.E) + 15) & ~uintptr_t(15));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, hf = warp >> 2;            // TMEM lane quadrant of this warp; which quarter of the columns
    const int r = quad * 32 + lane;                       // position (accumulator row) of this thread inside the item

    // ---- one-time: weights to bf16 swizzled operand tiles, per-kernel constants, tables, barriers, TMEM
    for (int i = tid; i < KB * KB / 8; i += NT) {
        const int n = i >> 4, k0 = (i & 15) * 8;
        const float4 a = *reinterpret_cast<const float4*>(p.w1 + n * KB + k0), b = *reinterpret_cast<const float4*>(p.w1 + n * KB + k0 + 4);
        st_swz(sW1, CHUNK_A, n, k0, make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w)));
    }
    for (int i = tid; i < NH * KB / 8; i += NT) {
        const int n = i >> 4, k0 = (i & 15) * 8;
        const float4 a = *reinterpret_cast<const float4*>(p.w2 + n * KB + k0), b = *reinterpret_cast<const float4*>(p.w2 + n * KB + k0 + 4);
        st_swz(sW2, CHUNK_W2, n, k0, make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w)));
    }
    for (int i = tid; i < KB; i += NT) {
        const float sg = fabsf(p.stds[i]) + 1e-5f;
        const float is = 0.84932180028801904272f / sg;            // sqrt(0.5 * log2(e)) / sigma: ex2(-((u - mu) is')^2) = exp(-0.5 ((u - mu)/sigma)^2)
        s_is[i] = is;
        s_mu[i] = p.means[i] * is;
        s_lc[i] = log2f(1.f / (sqrtf(2.f * 3.14159f) * sg));
        s_b1[i] = p.b1[i];
    }
    for (int i = tid; i < NH; i += NT) s_b2[i] = p.b2[i];
    for (int i = tid; i < p.E; i += NT) { s_mul[i] = p.mul[i]; s_bias[i] = p.bias[i]; }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    fence_proxy_async();                                  // the weight tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);

    constexpr uint32_t idesc1 = instr_desc_mn(TQ, KB, 0, 0), idesc2 = instr_desc_mn(TQ, NH, 0, 0);
    const uint64_t a_desc = smem_desc(smem_u32(sA), 16, 1024);
    const uint64_t w1_desc = smem_desc(smem_u32(sW1), 16, 1024);
    const uint64_t w2_desc = smem_desc(smem_u32(sW2), 16, 1024);

    const int LLp = p.L * p.Lp;
    const int n_items = p.B * p.tiles_per_mol;
    // the per-position inputs (edge type, distance, key-padding flag) of an item, fetched one item ahead
    struct Gather { long long e; float d; bool real, neg; };
    auto gather = [&](int item) {
        Gather gt;
        gt.e = 0; gt.d = 0.f; gt.real = false; gt.neg = true;
        if (item < n_items) {
            const int b = item / p.tiles_per_mol, t = item - b * p.tiles_per_mol;
            const int q = t * TQ + r;
            const int i = q / p.Lp, j = q - i * p.Lp;
            gt.real = q < LLp && j < p.L;                     // a pair of the molecule (not layout padding)
            if (gt.real) {
                const long long P = ((long long)b * p.L + i) * p.L + j;
                gt.e = p.et[P];
                gt.d = p.dist[P];
                gt.neg = p.key_pad && p.key_pad[b * p.L + j];
            }
        }
        return gt;
    };
    Gather cur = gather(blockIdx.x);
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, phase ^= 1u) {
        const int b = item / p.tiles_per_mol, t = item - b * p.tiles_per_mol;
        // ---- stage 1: u, basis row half [hf*64, hf*64 + 64) -> A tile
        float u = 0.f;
        const bool neg = cur.neg;
        if (cur.real) {
            const long long e = cur.e < 0 ? 0 : (cur.e >= p.E ? p.E - 1 : cur.e);
            u = fmaf(s_mul[e], cur.d, s_bias[e]);
        }
        const Gather nxt = gather(item + gridDim.x);          // in flight during this item's stages
        if (tid == 0) bulk_wait_read<0>();                    // the previous item's store has drained the store tile
#pragma unroll
        for (int k8 = 0; k8 < KB / NPART / 8; ++k8) {
            const int k0 = hf * (KB / NPART) + k8 * 8;
            float g[8];
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
                const float4 is4 = *reinterpret_cast<const float4*>(s_is + k0 + e), mu4 = *reinterpret_cast<const float4*>(s_mu + k0 + e);
                const float4 lc4 = *reinterpret_cast<const float4*>(s_lc + k0 + e);
                const float r0 = fmaf(u, is4.x, -mu4.x), r1 = fmaf(u, is4.y, -mu4.y), r2 = fmaf(u, is4.z, -mu4.z), r3 = fmaf(u, is4.w, -mu4.w);
                g[e] = fast_ex2(fmaf(-r0, r0, lc4.x));
                g[e + 1] = fast_ex2(fmaf(-r1, r1, lc4.y));
                g[e + 2] = fast_ex2(fmaf(-r2, r2, lc4.z));
                g[e + 3] = fast_ex2(fmaf(-r3, r3, lc4.w));
            }
            st_swz(sA, CHUNK_A, r, k0, make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7])));
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 1
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < KB / 16; ++s) {
                const uint64_t off = (uint64_t)((s >> 2) * (CHUNK_A >> 4) + (s & 3) * 2);
                tc_mma(tmem_base, a_desc + off, w1_desc + off, idesc1, s > 0 ? 1u : 0u);
            }
            tc_commit(&bars[0]);
        }
        mbar_wait_g(&bars[0], phase);
        tc_fence_after();
        // ---- stage 2: h = gelu(z + b1), columns [hf*64, hf*64 + 64) -> A tile (MMA 1 has finished reading it)
#pragma unroll
        for (int c32 = 0; c32 < KB / NPART / 32; ++c32) {
            const int c0 = hf * (KB / NPART) + c32 * 32;
            uint32_t v[32];
            tc_ld32(lane_addr + c0, v);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                const float4 b0 = *reinterpret_cast<const float4*>(s_b1 + c0 + g8 * 8), b1 = *reinterpret_cast<const float4*>(s_b1 + c0 + g8 * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = gelu_fast_val(__uint_as_float(v[g8 * 8 + e]) + bb[e]);
                st_swz(sA, CHUNK_A, r, c0 + g8 * 8, make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7])));
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- MMA 2
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < KB / 16; ++s) {
                const uint64_t offa = (uint64_t)((s >> 2) * (CHUNK_A >> 4) + (s & 3) * 2);
                const uint64_t offw = (uint64_t)((s >> 2) * (CHUNK_W2 >> 4) + (s & 3) * 2);
                tc_mma(tmem_base + KB, a_desc + offa, w2_desc + offw, idesc2, s > 0 ? 1u : 0u);
            }
            tc_commit(&bars[1]);
        }
        mbar_wait_g(&bars[1], phase);
        tc_fence_after();
        // ---- stage 3: heads [hf*32, hf*32 + 32) of this position -> store tile [head][position]
        {
            constexpr int HPT = NH / NPART;               // heads per thread
            uint32_t v[32];
            if (HPT == 32) tc_ld32(lane_addr + KB + hf * HPT, v);
            else tc_ld16(lane_addr + KB + hf * HPT, v);
            if constexpr (sizeof(TP) == 2) {
                // lanes l, l^1 hold adjacent positions: the even lane stores head e of both, the odd lane head e + 1 of both,
                // as one 32-bit word each (no sub-word bank conflicts, half the store instructions)
                const bool odd = lane & 1;
#pragma unroll
                for (int e = 0; e < HPT; e += 2) {
                    const float a0 = neg ? -INFINITY : __uint_as_float(v[e]) + s_b2[hf * HPT + e];
                    const float a1 = neg ? -INFINITY : __uint_as_float(v[e + 1]) + s_b2[hf * HPT + e + 1];
                    const float got = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);      // partner's value of the head this lane stores
                    const int h = hf * HPT + e + (odd ? 1 : 0);
                    const float lo = odd ? got : a0, hi = odd ? a1 : got;                  // positions (r & ~1), (r | 1)
                    *reinterpret_cast<uint32_t*>(sOt + h * TQ + (r & ~1)) = pack2<TP>(lo, hi);
                }
            } else {
#pragma unroll
                for (int e = 0; e < HPT; ++e) {
                    const int h = hf * HPT + e;
                    sOt[h * TQ + r] = from_f<TP>(neg ? -INFINITY : __uint_as_float(v[e]) + s_b2[h]);
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tma_store_3d(&tmOut, sOt, t * TQ, 0, b);
            bulk_commit();
        }
        cur = nxt;
    }
    if (tid == 0) bulk_wait_all<0>();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (B, H, L*Lp) output viewed as dims (q, h, b); box (128 positions, 64 heads, 1 molecule), no swizzle
int make_out_map3(CUtensorMap* map, void* out, int B, int LLp, int dtype) {
    static thread_local bool bound = false;
    if (!bound) { cudaFree(0); bound = true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) { mmdti_set_error("cuTensorMapEncodeTiled is not available from the driver"); return MMDTI_ERR_CUDA; }
    const size_t es = dtype == MMDTI_F32 ? 4 : 2;
    const cuuint64_t gdim[3] = {(cuuint64_t)LLp, (cuuint64_t)NH, (cuuint64_t)B};
    const cuuint64_t gstr[2] = {(cuuint64_t)LLp * es, (cuuint64_t)LLp * NH * es};
    const cuuint32_t box[3] = {(cuuint32_t)TQ, (cuuint32_t)NH, 1u};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = dtype == MMDTI_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                     : (dtype == MMDTI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    const CUresult r = fn(map, dt, 3, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { mmdti_set_error("cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r); return MMDTI_ERR_CUDA; }
    return MMDTI_OK;
}

template <typename TP>
int launch_k1(const CUtensorMap& tm, const K1Params& p, cudaStream_t st) {
    const size_t smem = K1Smem<TP>::tab + (size_t)2 * p.E * sizeof(float) + 16 + 64 + 1024;
    auto kern = pair_bias_fwd_tc5_kernel<TP>;
    MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // two CTAs per SM (each allocates 256 of the 512 TMEM columns) whenever their shared memory fits: 2 x (smem + 1 KB) <= 227 KB
    const int occ = 2 * (smem + 1024) <= 227 * 1024 ? 2 : 1;
    const long long items = (long long)p.B * p.tiles_per_mol;
    const int grid = (int)std::min<long long>(items, (long long)sms * occ);
    kern<<<grid, NT, smem, st>>>(tm, p);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

}  // namespace

// called by mmdti_pair_bias_fwd (pair_bias.cu) for the bf16-operand mode
int mmdti_pair_bias_fwd_tc5(const float* dist, const long long* et, const float* means, const float* stds, const float* mul,
                            const float* bias, const float* w1, const float* b1, const float* w2, const float* b2,
                            const unsigned char* key_pad, void* out, int B, int L, int Lp, int E, int pair_dtype, cudaStream_t st) {
    MMDTI_REQUIRE(mmdti_aligned(out, 16) && mmdti_aligned(w1, 16) && mmdti_aligned(w2, 16), "pair_bias_fwd: out / w1 / w2 must be 16-byte aligned");
    MMDTI_REQUIRE((size_t)2 * E * sizeof(float) <= 64 * 1024, "pair_bias_fwd: too many edge types (%d)", E);
    K1Params p;
    p.dist = dist; p.et = et; p.means = means; p.stds = stds; p.mul = mul; p.bias = bias; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.key_pad = key_pad; p.B = B; p.L = L; p.Lp = Lp; p.E = E;
    p.tiles_per_mol = (L * Lp + TQ - 1) / TQ;
    CUtensorMap tm;
    if (int rc = make_out_map3(&tm, out, B, L * Lp, pair_dtype)) return rc;
    if (pair_dtype == MMDTI_BF16) return launch_k1<bf16>(tm, p, st);
    if (pair_dtype == MMDTI_F16) return launch_k1<__half>(tm, p, st);
    if (pair_dtype == MMDTI_F32) return launch_k1<float>(tm, p, st);
    mmdti_set_error("pair_bias_fwd: bad pair_dtype %d", pair_dtype);
    return MMDTI_ERR_ARG;
}
