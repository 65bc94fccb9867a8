// K2 — pair-biased multi-head self-attention (head_dim 8), forward and backward.
//
// Data layout.  The pair tensor is carried as (B, H, L, Lp) with Lp = 8*NKB, NKB odd, so every
// row is 16-byte aligned, the (L x Lp) tile of one (molecule, head) is ONE contiguous 16-byte
// aligned range, and Lp == 8 (mod 16) makes the same image bank-conflict-free in shared
// memory.  Padding columns [L, Lp) hold -inf (an invariant every producer keeps), so they
// vanish in the softmax without any column predicate.
//
// Execution.  Persistent CTAs walk a contiguous range of (tile, row-chunk) work items through a
// 2-stage shared-memory ring: the pair slab of item w+1 arrives by ONE TMA bulk copy
// (cp.async.bulk + mbarrier complete_tx) and the q/k/v head rows by 16-byte cp.async while item
// w is computed; the updated slab leaves by one TMA bulk store.  Inside a CTA each warp owns
// 16 query rows x all keys in the register layout of mma.sync m16n8k8 (Q K^T: K = head_dim = 8)
// and m16n8k16 (A V: N = head_dim = 8).  32 flop per 4..8 bytes of pair traffic: the kernel
// is bound by HBM / instruction issue, not by the tensor pipe.
//
// Reference semantics: Uni-Core SelfMultiheadAttention(return_attn=True) as driven by
// models/transformers.py:136-139 (see include/mmdti_b200.h).
#include "common.cuh"

#include <math.h>
#include <algorithm>

namespace {

constexpr int HD = MMDTI_HEAD_DIM;  // 8
constexpr int NSTAGE = 2;      // backward ring
constexpr bool K2_BWD_CS_DEFAULT = true;   // column-split backward for L > 136 (MMDTI_K2_BWD_CS=0/1 overrides)
constexpr bool K2_FWD_CS_DEFAULT = true;   // column-split forward for L > 136 (MMDTI_K2_FWD_CS=0/1 overrides)
constexpr int NSTAGE_F = 3;    // forward ring: the bulk store of item w-1 may still be reading its stage while w+1 loads
constexpr float LOG2E = 1.4426950408889634f;

template <int NKB> struct Geo {
    static_assert(NKB % 2 == 1, "NKB must be odd so that the row stride is 8 mod 16");
    static constexpr int KP = NKB * 8;        // padded key count == pair row stride (elements)
    static constexpr int STRIDE = KP;
    static constexpr int NKB16 = (NKB + 1) / 2;
    static constexpr int KROWS = KP + 8;      // K/V smem rows (last 16-key block reads 8 rows past KP)
    static constexpr int MAXKB16 = NKB <= 9 ? 1 : (NKB <= 17 ? 3 : 6);   // 16-key blocks per warp (bwd phase 2)
};

struct FwdParams {
    const void *q, *k, *v;
    void* o;
    const void* pin;
    void* pout;
    long long ldqkv, ldo;
    int B, H, L;
    float scale, keep_scale;
    uint32_t thresh16;
    unsigned long long seed;
    const unsigned long long* seed_off;
    int crb, nchunks;     // 16-row blocks per chunk, chunks per tile
};

struct BwdParams {
    const void *q, *k, *v, *o, *s, *d_o, *dpout;
    void *dpin, *dq, *dk, *dv;
    long long ldqkv, lddo, lddqkv;
    int B, H, L;
    float scale, keep_scale;
    uint32_t thresh16;
    unsigned long long seed;
    const unsigned long long* seed_off;
    int crb, nchunks;
    int* sched;       // {next pool tile, CTAs done}: dynamic tile scheduler of the backward (NULL: static contiguous ranges)
    int pool;         // tiles handed out dynamically (the LAST `pool` tiles); the others are split into contiguous ranges
};

// two adjacent pair elements (col even) of a slab row -> floats
template <typename TP> __device__ __forceinline__ float2 slab_get2(const TP* row, int col) {
    if constexpr (sizeof(TP) == 2) {
        return unpack2<TP>(*reinterpret_cast<const uint32_t*>(row + col));
    } else {
        return *reinterpret_cast<const float2*>(row + col);
    }
}
// store two adjacent elements; returns the values as stored (rounded to TP)
template <typename TP> __device__ __forceinline__ float2 slab_put2(TP* row, int col, float a, float b) {
    if constexpr (sizeof(TP) == 2) {
        const uint32_t u = pack2<TP>(a, b);
        *reinterpret_cast<uint32_t*>(row + col) = u;
        return unpack2<TP>(u);
    } else {
        *reinterpret_cast<float2*>(row + col) = make_float2(a, b);
        return make_float2(a, b);
    }
}

// cp.async one head row (8 elements of T) global -> shared
template <typename T> __device__ __forceinline__ void cp_head_row(T* dst, const T* src) {
    cp_async_16(dst, src);
    if constexpr (sizeof(T) == 4) cp_async_16(dst + 4, src + 4);
}

template <typename T> __device__ __forceinline__ void zero_fill(T* p, int n, int tid, int nthr) {
    for (int i = tid; i < n; i += nthr) p[i] = from_f<T>(0.f);
}

// ===================================================================== forward
template <typename T, typename TP, int NKB>
struct FwdStage {
    using G = Geo<NKB>;
    T* K;      // [KROWS][8]
    T* V;      // [KROWS][8]
    T* Q;      // [NR][8]
    TP* slab;  // [NR][STRIDE]
    static __host__ __device__ size_t bytes(int NR) {
        return (size_t)(2 * G::KROWS + NR) * HD * sizeof(T) + (size_t)NR * G::STRIDE * sizeof(TP);
    }
    __device__ void carve(unsigned char* base, int NR) {
        K = reinterpret_cast<T*>(base);
        V = K + G::KROWS * HD;
        Q = V + G::KROWS * HD;
        slab = reinterpret_cast<TP*>(Q + NR * HD);
    }
};

// threads per CTA never exceed 32 * ceil(NKB/2) (one warp per 16-row block of a <= 8-block chunk); the
// register budget is capped so that ~640 threads stay resident per SM (4 CTAs of 160 threads at L = 66)
template <int NKB> struct FwdLaunch {
    static constexpr int MAXT = (NKB + 1) / 2 * 32 < 256 ? (NKB + 1) / 2 * 32 : 256;
    // NKB >= 25 (L > 136): a thread holds >= 100 score registers and shared memory limits the SM to 2 CTAs of <= 96
    // threads anyway, so the register cap is lifted (no spills) instead of kept at 128
    static constexpr int MINB = NKB >= 25 ? 1 : (640 / MAXT < 1 ? 1 : (640 / MAXT > 8 ? 8 : 640 / MAXT));
};

// CS (column split, bf16 activations, NKB >= 25): as in the backward, 4 warps share a 16-row block, each owning a quarter of the
// key blocks (<= 9 instead of 33 score blocks in registers): 32-row chunks, 8 warps, two CTAs per SM.  The row max is exchanged
// before the exponentials, the row sums and the partial A'V products after them.
template <typename T, typename TP, int NKB, bool CS = false>
__global__ void __launch_bounds__(CS ? 256 : FwdLaunch<NKB>::MAXT, CS ? 2 : FwdLaunch<NKB>::MINB)
pair_attn_fwd_kernel(const FwdParams p) {
    using G = Geo<NKB>;
    constexpr bool F32 = std::is_same<T, float>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NSTAGE_F];

    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q4 = lane & 3;
    const int L = p.L, NR = p.crb * 16;

    const size_t stage_bytes = (FwdStage<T, TP, NKB>::bytes(NR) + 127) & ~size_t(127);
    // stage views are re-derived from the shared-memory base on every use: keeps the pointers provably
    // shared (LDS/STS instead of generic LD/ST through a local array of pointers)
    auto stage = [&](int s) {
        FwdStage<T, TP, NKB> x;
        x.carve(smem_raw + (size_t)s * stage_bytes, NR);
        return x;
    };

    // one-time init: zero the whole ring (K/V rows beyond L stay zero for ever; slab rows the TMA never
    // writes must hold finite bit patterns, not NaNs), barriers
    for (size_t i = tid; i < NSTAGE_F * stage_bytes / 4; i += nthr) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE_F; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int ntiles = p.B * p.H;
    const int t0 = (int)((long long)ntiles * blockIdx.x / gridDim.x), t1 = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);
    const int w0 = t0 * p.nchunks, w1 = t1 * p.nchunks;
    const size_t tile_elems = (size_t)L * G::STRIDE;
    const FastDiv div_h((uint32_t)p.H), div_c((uint32_t)p.nchunks);
    // who gathers the q / k / v rows (see prefetch): the last warp alone when its row block is a half block of a
    // single-chunk tile, everybody otherwise
    const int nwarps = nthr >> 5;
    const bool light_last = !CS && p.nchunks == 1 && nwarps > 1 && (L - (nwarps - 1) * 16) <= 8 && (L - (nwarps - 1) * 16) > 0;
    const int pf_lane0 = light_last ? (warp == nwarps - 1 ? lane : -1) : tid;
    const int pf_step = light_last ? 32 : nthr;

    auto prefetch = [&](int w, int s) {
        const int tile = (int)div_c.div((uint32_t)w);
        const int chunk = w - tile * p.nchunks;
        const int b = (int)div_h.div((uint32_t)tile), h = tile - b * p.H;
        const int row0 = chunk * NR, nrows = min(L - row0, NR);
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)((size_t)nrows * G::STRIDE * sizeof(TP));
            mbar_arrive_expect_tx(&full_bar[s], bytes);
            bulk_g2s(stage(s).slab, static_cast<const TP*>(p.pin) + (size_t)tile * tile_elems + (size_t)row0 * G::STRIDE, bytes,
                     &full_bar[s]);
        }
        const T* qg = static_cast<const T*>(p.q) + (size_t)b * L * p.ldqkv + h * HD;
        const T* kg = static_cast<const T*>(p.k) + (size_t)b * L * p.ldqkv + h * HD;
        const T* vg = static_cast<const T*>(p.v) + (size_t)b * L * p.ldqkv + h * HD;
        // light_last: the CTA's last warp owns a half row block (L = 66: rows 64, 65) and would idle at the barrier for
        // three quarters of an item; it fetches the q / k / v head rows of the next item ALONE, so that the full-block
        // warps spend no instruction on the gather
        const int i0 = pf_lane0, istep = pf_step;
        if (i0 >= 0) {
            for (int i = i0; i < 2 * L + nrows; i += istep) {
                if (i < L) cp_head_row(stage(s).K + i * HD, kg + (size_t)i * p.ldqkv);
                else if (i < 2 * L) cp_head_row(stage(s).V + (i - L) * HD, vg + (size_t)(i - L) * p.ldqkv);
                else cp_head_row(stage(s).Q + (i - 2 * L) * HD, qg + (size_t)(row0 + i - 2 * L) * p.ldqkv);
            }
        }
        cp_async_commit();
    };

    const bool do_drop = p.thresh16 != 0;
    const unsigned long long eff_seed = do_drop ? rng_effective_seed(p.seed, p.seed_off) : 0ull;      // once per kernel
    if (w0 < w1) prefetch(w0, 0);

    int it = 0;
    int s = 0, par = 0;                 // ring position of item w and the phase parity of its mbarrier
    for (int w = w0; w < w1; ++w, ++it) {
        const int tile = (int)div_c.div((uint32_t)w);
        const int chunk = w - tile * p.nchunks;
        const int b = (int)div_h.div((uint32_t)tile), h = tile - b * p.H;
        const int row0 = chunk * NR, nrows = min(L - row0, NR);

        cp_async_wait<0>();
        mbar_wait(&full_bar[s], par);
        __syncthreads();                    // item w is resident; everyone is done with item w-1
        const int sn = s + 1 == NSTAGE_F ? 0 : s + 1;
        if (w + 1 < w1) {
            // stage sn was last read by the bulk store of item w-2: allow the store of item w-1 to stay in flight
            if (tid == 0) bulk_wait_read<1>();
            prefetch(w + 1, sn);
        }

        // One 16-row block per warp.  HB = false: rows +8..15 of the block lie beyond L (the leftover block of L = 66 holds 2
        // rows), so every per-element instruction of that half is dropped; its mma operands are zero.
        auto row_block = [&](auto hb_tag) {
            constexpr bool HB = decltype(hb_tag)::value;
            const T* Ks = stage(s).K;
            const T* Vs = stage(s).V;
            const T* Qs = stage(s).Q;
            const int la = warp * 16 + g, lb = la + 8;               // local rows
            const int ra = row0 + la, rbb = row0 + lb;                // global rows
            TP* srow_a = stage(s).slab + la * G::STRIDE;
            TP* srow_b = stage(s).slab + lb * G::STRIDE;
            float sc[NKB][4];

            // ---- S = Q K^T
            if constexpr (!F32) {
                const uint32_t* Q32 = reinterpret_cast<const uint32_t*>(Qs);
                const uint32_t* K32 = reinterpret_cast<const uint32_t*>(Ks);
                const uint32_t qa0 = Q32[la * 4 + q4], qa1 = HB ? Q32[lb * 4 + q4] : 0u;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    sc[kb][0] = sc[kb][1] = sc[kb][2] = sc[kb][3] = 0.f;
                    mma_bf16_1688(sc[kb], qa0, qa1, K32[(kb * 8 + g) * 4 + q4]);
                }
            } else {
                float qa[HD], qb[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) { qa[d] = Qs[la * HD + d]; qb[d] = HB ? Qs[lb * HD + d] : 0.f; }
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* kr = Ks + (kb * 8 + 2 * q4 + e) * HD;
                        float da = 0.f, db = 0.f;
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            da = fmaf(qa[d], kr[d], da);
                            db = fmaf(qb[d], kr[d], db);
                        }
                        sc[kb][e] = da;
                        sc[kb][2 + e] = db;
                    }
                }
            }

            // ---- S = scale*S + P ; P' := S (stored, rounded to TP) ; row max
            float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const int col = kb * 8 + 2 * q4;
                const float2 pa = slab_get2<TP>(srow_a, col);
                const float2 va = slab_put2<TP>(srow_a, col, fmaf(sc[kb][0], p.scale, pa.x), fmaf(sc[kb][1], p.scale, pa.y));
                sc[kb][0] = va.x; sc[kb][1] = va.y;
                ma = fmaxf(ma, fmaxf(va.x, va.y));
                if constexpr (HB) {
                    const float2 pb = slab_get2<TP>(srow_b, col);
                    const float2 vb = slab_put2<TP>(srow_b, col, fmaf(sc[kb][2], p.scale, pb.x), fmaf(sc[kb][3], p.scale, pb.y));
                    sc[kb][2] = vb.x; sc[kb][3] = vb.y;
                    mb = fmaxf(mb, fmaxf(vb.x, vb.y));
                } else {
                    sc[kb][2] = sc[kb][3] = 0.f;
                }
            }
            ma = quad_max(ma);
            if constexpr (HB) mb = quad_max(mb);

            // ---- softmax numerators, row sums, dropout
            float suma = 0.f, sumb = 0.f;
            if constexpr (F32) {
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    sc[kb][0] = expf(sc[kb][0] - ma); sc[kb][1] = expf(sc[kb][1] - ma);
                    if constexpr (HB) { sc[kb][2] = expf(sc[kb][2] - mb); sc[kb][3] = expf(sc[kb][3] - mb); }
                }
            } else {
                const float ka = ma * LOG2E, kbm = mb * LOG2E;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    sc[kb][0] = fast_ex2(fmaf(sc[kb][0], LOG2E, -ka));
                    sc[kb][1] = fast_ex2(fmaf(sc[kb][1], LOG2E, -ka));
                    if constexpr (HB) {
                        sc[kb][2] = fast_ex2(fmaf(sc[kb][2], LOG2E, -kbm));
                        sc[kb][3] = fast_ex2(fmaf(sc[kb][3], LOG2E, -kbm));
                    }
                }
            }
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                suma += sc[kb][0] + sc[kb][1];
                if constexpr (HB) sumb += sc[kb][2] + sc[kb][3];
            }
            // dropout: one hash per 4 elements; the odd column's 16 bits are compared in place (bits >= t << 16),
            // the even column's after one shift
            if (do_drop) {
                const uint32_t rkey = rng_stream_key(eff_seed, (uint32_t)tile);
                const uint32_t thi = p.thresh16 << 16;
#pragma unroll
                for (int kb = 0; kb < NKB; kb += 2) {
                    const uint2 wa = rng_quad_bits(rkey, ra, kb * 8 + 2 * q4);
                    if ((wa.x << 16) < thi) sc[kb][0] = 0.f;
                    if (wa.x < thi) sc[kb][1] = 0.f;
                    if (kb + 1 < NKB) {
                        if ((wa.y << 16) < thi) sc[kb + 1][0] = 0.f;
                        if (wa.y < thi) sc[kb + 1][1] = 0.f;
                    }
                    if constexpr (HB) {
                        const uint2 wb = rng_quad_bits(rkey, rbb, kb * 8 + 2 * q4);
                        if ((wb.x << 16) < thi) sc[kb][2] = 0.f;
                        if (wb.x < thi) sc[kb][3] = 0.f;
                        if (kb + 1 < NKB) {
                            if ((wb.y << 16) < thi) sc[kb + 1][2] = 0.f;
                            if (wb.y < thi) sc[kb + 1][3] = 0.f;
                        }
                    }
                }
            }
            suma = quad_sum(suma);
            const float inva = F32 ? p.keep_scale / suma : p.keep_scale * fast_rcp(suma);
            float invb = 0.f;
            if constexpr (HB) {
                sumb = quad_sum(sumb);
                invb = F32 ? p.keep_scale / sumb : p.keep_scale * fast_rcp(sumb);
            }
            T* og = static_cast<T*>(p.o) + (size_t)b * L * p.ldo + h * HD;

            // ---- O = A' V
            if constexpr (!F32) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < G::NKB16; j += 2) {
                    uint32_t b0, b1, b2 = 0u, b3 = 0u;
                    if (j + 1 < G::NKB16) ldmatrix_x4_trans(b0, b1, b2, b3, Vs + (j * 16 + lane) * HD);
                    else ldmatrix_x2_trans(b0, b1, Vs + (j * 16 + (lane & 15)) * HD);
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        if (j + jj < G::NKB16) {
                            const int kb0 = 2 * (j + jj), kb1 = kb0 + 1;
                            const uint32_t a0 = pack_bf16(sc[kb0][0], sc[kb0][1]);
                            const uint32_t a1 = pack_bf16(sc[kb0][2], sc[kb0][3]);
                            uint32_t a2 = 0u, a3 = 0u;
                            if (kb1 < NKB) {
                                a2 = pack_bf16(sc[kb1][0], sc[kb1][1]);
                                a3 = pack_bf16(sc[kb1][2], sc[kb1][3]);
                            }
                            mma_bf16_16816(o, a0, a1, a2, a3, jj ? b2 : b0, jj ? b3 : b1);
                        }
                    }
                }
                if (ra < L)
                    *reinterpret_cast<uint32_t*>(og + (size_t)ra * p.ldo + 2 * q4) = pack_bf16(o[0] * inva, o[1] * inva);
                if (HB && rbb < L)
                    *reinterpret_cast<uint32_t*>(og + (size_t)rbb * p.ldo + 2 * q4) = pack_bf16(o[2] * invb, o[3] * invb);
            } else {
                float oa[HD], ob[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) oa[d] = ob[d] = 0.f;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* vr = Vs + (kb * 8 + 2 * q4 + e) * HD;
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            oa[d] = fmaf(sc[kb][e], vr[d], oa[d]);
                            ob[d] = fmaf(sc[kb][2 + e], vr[d], ob[d]);
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < HD; ++d) {
                    oa[d] = quad_sum(oa[d]) * inva;
                    if constexpr (HB) ob[d] = quad_sum(ob[d]) * invb;
                }
                if (q4 == 0) {
                    if (ra < L) {
                        float4* dst = reinterpret_cast<float4*>(og + (size_t)ra * p.ldo);
                        dst[0] = make_float4(oa[0], oa[1], oa[2], oa[3]);
                        dst[1] = make_float4(oa[4], oa[5], oa[6], oa[7]);
                    }
                    if (HB && rbb < L) {
                        float4* dst = reinterpret_cast<float4*>(og + (size_t)rbb * p.ldo);
                        dst[0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
                        dst[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
                    }
                }
            }
        };
        if constexpr (CS) {
            constexpr int NSPLIT = 4;
            constexpr int KBW = (NKB + NSPLIT - 1) / NSPLIT;          // key blocks per warp
            __shared__ float cs_max[2][NSPLIT][16];
            __shared__ float cs_sum[2][NSPLIT][16];
            __shared__ float cs_o[2][NSPLIT - 1][16][HD];
            const int rb = warp >> 2, cs = warp & 3;                   // launched with 8 warps, crb = 2
            if (rb < p.crb && row0 + rb * 16 < L) {
                const int kb0 = cs * KBW;
                const T* Ks = stage(s).K;
                const T* Vs = stage(s).V;
                const T* Qs = stage(s).Q;
                const int la = rb * 16 + g, lb = la + 8;
                const int ra = row0 + la, rbb = row0 + lb;
                TP* srow_a = stage(s).slab + la * G::STRIDE;
                TP* srow_b = stage(s).slab + lb * G::STRIDE;
                const unsigned bar_id = 1u + (unsigned)rb;
                float sc[KBW][4];

                // ---- S = Q K^T over this warp's key blocks; S = scale*S + P; P' := S (stored); row max
                const uint32_t* Q32 = reinterpret_cast<const uint32_t*>(Qs);
                const uint32_t* K32 = reinterpret_cast<const uint32_t*>(Ks);
                const uint32_t qa0 = Q32[la * 4 + q4], qa1 = Q32[lb * 4 + q4];
                float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
                for (int j = 0; j < KBW; ++j) {
                    const int kb = kb0 + j;
                    if (kb < NKB) {
                        const int col = kb * 8 + 2 * q4;
                        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                        mma_bf16_1688(sc[j], qa0, qa1, K32[(kb * 8 + g) * 4 + q4]);
                        const float2 pa = slab_get2<TP>(srow_a, col), pb = slab_get2<TP>(srow_b, col);
                        const float2 va = slab_put2<TP>(srow_a, col, fmaf(sc[j][0], p.scale, pa.x), fmaf(sc[j][1], p.scale, pa.y));
                        const float2 vb = slab_put2<TP>(srow_b, col, fmaf(sc[j][2], p.scale, pb.x), fmaf(sc[j][3], p.scale, pb.y));
                        sc[j][0] = va.x; sc[j][1] = va.y; sc[j][2] = vb.x; sc[j][3] = vb.y;
                        ma = fmaxf(ma, fmaxf(va.x, va.y));
                        mb = fmaxf(mb, fmaxf(vb.x, vb.y));
                    } else {
                        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = -INFINITY;
                    }
                }
                ma = quad_max(ma);
                mb = quad_max(mb);
                if (q4 == 0) { cs_max[rb][cs][g] = ma; cs_max[rb][cs][g + 8] = mb; }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128) : "memory");
#pragma unroll
                for (int c = 0; c < NSPLIT; ++c) {
                    ma = fmaxf(ma, cs_max[rb][c][g]);
                    mb = fmaxf(mb, cs_max[rb][c][g + 8]);
                }

                // ---- softmax numerators, partial row sums (before dropout, like the row-split form), dropout
                float suma = 0.f, sumb = 0.f;
                {
                    const float ka = ma * LOG2E, kbm = mb * LOG2E;
#pragma unroll
                    for (int j = 0; j < KBW; ++j) {
                        if (kb0 + j < NKB) {
                            sc[j][0] = fast_ex2(fmaf(sc[j][0], LOG2E, -ka));
                            sc[j][1] = fast_ex2(fmaf(sc[j][1], LOG2E, -ka));
                            sc[j][2] = fast_ex2(fmaf(sc[j][2], LOG2E, -kbm));
                            sc[j][3] = fast_ex2(fmaf(sc[j][3], LOG2E, -kbm));
                            suma += sc[j][0] + sc[j][1];
                            sumb += sc[j][2] + sc[j][3];
                        } else {
                            sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                        }
                    }
                }
                if (do_drop) {
                    const uint32_t rkey = rng_stream_key(eff_seed, (uint32_t)tile);
                    const uint32_t thi = p.thresh16 << 16;
                    uint2 qwa = make_uint2(0u, 0u), qwb = make_uint2(0u, 0u);
#pragma unroll
                    for (int j = 0; j < KBW; ++j) {
                        const int kb = kb0 + j;
                        if (kb < NKB) {
                            {
                                // one hash serves the key-block pair (2m, 2m+1): .x even block, .y odd block
                                const int col = kb * 8 + 2 * q4;
                                if (j == 0 || (kb & 1) == 0) { qwa = rng_quad_bits(rkey, ra, col); qwb = rng_quad_bits(rkey, rbb, col); }
                                const uint32_t ba = (kb & 1) ? qwa.y : qwa.x, bb = (kb & 1) ? qwb.y : qwb.x;
                                if ((ba << 16) < thi) sc[j][0] = 0.f;
                                if (ba < thi) sc[j][1] = 0.f;
                                if ((bb << 16) < thi) sc[j][2] = 0.f;
                                if (bb < thi) sc[j][3] = 0.f;
                            }
                        }
                    }
                }
                suma = quad_sum(suma);
                sumb = quad_sum(sumb);

                // ---- partial O = A' V over this warp's key blocks (pairs of 8-key blocks from kb0), summed by warp 0
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int jp = 0; jp < (KBW + 1) / 2; ++jp) {
                    const int kbA = kb0 + 2 * jp;
                    if (kbA < NKB) {
                        uint32_t b0, b1;
                        ldmatrix_x2_trans(b0, b1, Vs + (kbA * 8 + (lane & 15)) * HD);
                        const uint32_t a0 = pack_bf16(sc[2 * jp][0], sc[2 * jp][1]);
                        const uint32_t a1 = pack_bf16(sc[2 * jp][2], sc[2 * jp][3]);
                        uint32_t a2 = 0u, a3 = 0u;
                        if (2 * jp + 1 < KBW) {
                            a2 = pack_bf16(sc[2 * jp + 1][0], sc[2 * jp + 1][1]);      // zeros beyond NKB
                            a3 = pack_bf16(sc[2 * jp + 1][2], sc[2 * jp + 1][3]);
                        }
                        mma_bf16_16816(o, a0, a1, a2, a3, b0, b1);
                    }
                }
                if (q4 == 0) { cs_sum[rb][cs][g] = suma; cs_sum[rb][cs][g + 8] = sumb; }
                if (cs > 0) {
                    *reinterpret_cast<float2*>(&cs_o[rb][cs - 1][g][2 * q4]) = make_float2(o[0], o[1]);
                    *reinterpret_cast<float2*>(&cs_o[rb][cs - 1][g + 8][2 * q4]) = make_float2(o[2], o[3]);
                }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128) : "memory");
                if (cs == 0) {
                    suma = 0.f;
                    sumb = 0.f;
#pragma unroll
                    for (int c = 0; c < NSPLIT; ++c) {
                        suma += cs_sum[rb][c][g];
                        sumb += cs_sum[rb][c][g + 8];
                    }
#pragma unroll
                    for (int c = 0; c < NSPLIT - 1; ++c) {
                        const float2 x = *reinterpret_cast<const float2*>(&cs_o[rb][c][g][2 * q4]);
                        const float2 y = *reinterpret_cast<const float2*>(&cs_o[rb][c][g + 8][2 * q4]);
                        o[0] += x.x; o[1] += x.y; o[2] += y.x; o[3] += y.y;
                    }
                    const float inva = p.keep_scale * fast_rcp(suma), invb = p.keep_scale * fast_rcp(sumb);
                    T* og = static_cast<T*>(p.o) + (size_t)b * L * p.ldo + h * HD;
                    if (ra < L)
                        *reinterpret_cast<uint32_t*>(og + (size_t)ra * p.ldo + 2 * q4) = pack_bf16(o[0] * inva, o[1] * inva);
                    if (rbb < L)
                        *reinterpret_cast<uint32_t*>(og + (size_t)rbb * p.ldo + 2 * q4) = pack_bf16(o[2] * invb, o[3] * invb);
                }
            }
        } else if (warp < p.crb && row0 + warp * 16 < L) {
            if (row0 + warp * 16 + 8 < L) row_block(std::true_type{});
            else row_block(std::false_type{});
        }
        // (tried: every warp bulk-storing its own 16 rows without the CTA barrier -- no gain: 80.3 vs 77 us)
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(static_cast<TP*>(p.pout) + tile * tile_elems + (size_t)row0 * G::STRIDE, stage(s).slab,
                     (uint32_t)((size_t)nrows * G::STRIDE * sizeof(TP)));
            bulk_commit();
        }
        if (sn == 0) par ^= 1;
        s = sn;
    }
    if (tid == 0) bulk_wait_all<0>();
}

// ===================================================================== backward
// delta_i = rowsum(dA o A) is taken as dO_i . O_i (O from the forward).
// Contract: the incoming d_pair_out is 0 wherever S = -inf (masked keys / padding columns);
// A is exactly 0 there, so dS = A o (dA - delta) + d_pair_out is 0 there too, unpredicated.
template <typename T, typename TP, typename TG, int NKB>
struct BwdStage {
    using G = Geo<NKB>;
    T *K, *V;            // [KROWS][8]
    T *Q, *dO, *O;       // [NR][8]
    TP* sS;              // [NR][STRIDE]
    TG* sG;              // [NR][STRIDE]
    static __host__ __device__ size_t bytes(int NR) {
        return (size_t)(2 * G::KROWS + 3 * NR) * HD * sizeof(T) + (size_t)NR * G::STRIDE * (sizeof(TP) + sizeof(TG));
    }
    __device__ void carve(unsigned char* base, int NR) {
        K = reinterpret_cast<T*>(base);
        V = K + G::KROWS * HD;
        Q = V + G::KROWS * HD;
        dO = Q + NR * HD;
        O = dO + NR * HD;
        sS = reinterpret_cast<TP*>(O + NR * HD);
        sG = reinterpret_cast<TG*>(sS + NR * G::STRIDE);
    }
};

// CS (column split, bf16 activations, NKB >= 25 i.e. L > 136): shared memory limits a chunk to 32 rows at two CTAs per SM, so
// with one warp per 16-row block only 2 warps per CTA would work in phase 1.  Instead 4 warps share a row block, each owning
// a quarter of the key blocks: row max / row sum / dQ partials are exchanged through shared memory behind a named barrier.
template <typename T, typename TP, typename TG, int NKB, bool CS = false>
__global__ void __launch_bounds__(256, CS ? 2 : 1) pair_attn_bwd_kernel(const BwdParams p) {
    using G = Geo<NKB>;
    constexpr bool F32 = std::is_same<T, float>::value;
    // dS in the mma operand type: aliases the dP slab when that already has this type
    constexpr bool ALIAS_DS = std::is_same<T, TG>::value;
    constexpr int MAXKB16 = CS ? (G::NKB16 + 7) / 8 : G::MAXKB16;      // CS: always 8 warps
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NSTAGE];

    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = nthr >> 5;
    const int g = lane >> 2, q4 = lane & 3;
    const int L = p.L, NR = p.crb * 16;
    const int nkb16 = (L + 15) >> 4;

    const size_t stage_bytes = (BwdStage<T, TP, TG, NKB>::bytes(NR) + 127) & ~size_t(127);
    auto stage = [&](int s) {
        BwdStage<T, TP, TG, NKB> x;
        x.carve(smem_raw + (size_t)s * stage_bytes, NR);
        return x;
    };
    unsigned char* extra = smem_raw + NSTAGE * stage_bytes;
    T* Apt = reinterpret_cast<T*>(extra);                            // [NR][STRIDE] dropped probabilities
    T* dSt_own = Apt + NR * G::STRIDE;                               // [NR][STRIDE] (only when !ALIAS_DS)
    float* dKacc = reinterpret_cast<float*>(ALIAS_DS ? dSt_own : dSt_own + NR * G::STRIDE);   // f32 path only
    float* dVacc = dKacc + G::KP * HD;

    // one-time init: zero the whole ring and the A'/dS buffers: K/V rows beyond L stay zero for ever, and
    // slab rows the TMA never writes must hold finite bit patterns (a NaN there would leak into dK/dV)
    {
        const size_t words = (NSTAGE * stage_bytes + (size_t)NR * G::STRIDE * sizeof(T) * (ALIAS_DS ? 1 : 2)) / 4;
        for (size_t i = tid; i < words; i += nthr) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int ntiles = p.B * p.H;
    const int t0 = (int)((long long)ntiles * blockIdx.x / gridDim.x), t1 = (int)((long long)ntiles * (blockIdx.x + 1) / gridDim.x);
    const FastDiv div_h((uint32_t)p.H);
    const size_t tile_elems = (size_t)L * G::STRIDE;
    const bool has_dpo = p.dpout != nullptr;

    auto prefetch = [&](int tile, int chunk, int s) {
        const int b = (int)div_h.div((uint32_t)tile), h = tile - b * p.H;
        const int row0 = chunk * NR, nrows = min(L - row0, NR);
        if (tid == 0) {
            const uint32_t bs = (uint32_t)((size_t)nrows * G::STRIDE * sizeof(TP));
            const uint32_t bg = has_dpo ? (uint32_t)((size_t)nrows * G::STRIDE * sizeof(TG)) : 0u;
            mbar_arrive_expect_tx(&full_bar[s], bs + bg);
            bulk_g2s(stage(s).sS, static_cast<const TP*>(p.s) + (size_t)tile * tile_elems + (size_t)row0 * G::STRIDE, bs, &full_bar[s]);
            if (has_dpo)
                bulk_g2s(stage(s).sG, static_cast<const TG*>(p.dpout) + (size_t)tile * tile_elems + (size_t)row0 * G::STRIDE, bg,
                         &full_bar[s]);
        }
        const size_t tok = (size_t)b * L;
        const T* qg = static_cast<const T*>(p.q) + tok * p.ldqkv + h * HD;
        const T* kg = static_cast<const T*>(p.k) + tok * p.ldqkv + h * HD;
        const T* vg = static_cast<const T*>(p.v) + tok * p.ldqkv + h * HD;
        const T* og = static_cast<const T*>(p.o) + tok * p.lddo + h * HD;
        const T* dog = static_cast<const T*>(p.d_o) + tok * p.lddo + h * HD;
        for (int i = tid; i < 2 * L + 3 * nrows; i += nthr) {
            if (i < L) cp_head_row(stage(s).K + i * HD, kg + (size_t)i * p.ldqkv);
            else if (i < 2 * L) cp_head_row(stage(s).V + (i - L) * HD, vg + (size_t)(i - L) * p.ldqkv);
            else {
                const int j = i - 2 * L, which = j >= 2 * nrows ? 2 : (j >= nrows ? 1 : 0), r = j - which * nrows;
                if (which == 0) cp_head_row(stage(s).Q + r * HD, qg + (size_t)(row0 + r) * p.ldqkv);
                else if (which == 1) cp_head_row(stage(s).dO + r * HD, dog + (size_t)(row0 + r) * p.lddo);
                else cp_head_row(stage(s).O + r * HD, og + (size_t)(row0 + r) * p.lddo);
            }
        }
        cp_async_commit();
    };

    const bool do_drop = p.thresh16 != 0;
    const unsigned long long eff_seed = do_drop ? rng_effective_seed(p.seed, p.seed_off) : 0ull;      // once per kernel
    float dk_acc[MAXKB16][4], dv_acc[MAXKB16][4];

    // Work distribution.  A tile = one (molecule, head) with all its row chunks (dK / dV accumulate across the chunks).
    // p.sched == NULL: every CTA walks a contiguous range of tiles (consecutive heads of a molecule share the cache lines of
    // their q / k / v rows).  Otherwise only the first ntiles - p.pool tiles are split that way and the last p.pool tiles are
    // handed out DYNAMICALLY from a global counter once a CTA has finished its range, so CTAs that share their SM with a kernel
    // of another stream -- the NCCL all-reduce of the gradient buckets, the weight-gradient GEMMs of the side stream -- shed
    // work instead of holding the whole kernel back.  Tile ids travel two ahead through s_tile[] (thread 0 fetches the id of
    // the tile after next while the current tile computes), so the atomic's latency is never exposed.
    __shared__ int s_tile[2];
    const bool dyn = p.sched != nullptr;
    const int nstat_all = ntiles - p.pool;
    const int st0 = (int)((long long)nstat_all * blockIdx.x / gridDim.x), st1 = (int)((long long)nstat_all * (blockIdx.x + 1) / gridDim.x);
    int n_fetched = 0;                                   // thread 0: ids taken so far
    auto next_id = [&]() -> int {                        // thread 0 only
        const int k = n_fetched++;
        return st0 + k < st1 ? st0 + k : nstat_all + atomicAdd(p.sched, 1);
    };
    int tile = t0, chunk = 0, kt = 0, fetched = 0;
    if (dyn) {
        if (tid == 0) {
            s_tile[0] = next_id();
            s_tile[1] = next_id();
        }
        __syncthreads();
        tile = s_tile[0];
    }
    const int tile_end = dyn ? ntiles : t1;
    if (tile < tile_end) prefetch(tile, 0, 0);
    for (int it = 0; tile < tile_end; ++it) {
        const int s = it & 1;
        const int b = (int)div_h.div((uint32_t)tile), h = tile - b * p.H;
        const int row0 = chunk * NR, nrows = min(L - row0, NR);
        const int nrb_chunk = (nrows + 15) >> 4;

        cp_async_wait<0>();
        mbar_wait(&full_bar[s], (it >> 1) & 1);
        __syncthreads();        // item resident; the previous item (incl. its phase 2 readers) finished everywhere
        int ntile = tile, nchunk = chunk + 1;
        if (nchunk == p.nchunks) {
            nchunk = 0;
            ntile = dyn ? s_tile[(kt + 1) & 1] : tile + 1;
        }
        if (ntile < tile_end) {
            if (tid == 0) bulk_wait_read<0>();
            prefetch(ntile, nchunk, s ^ 1);
        }
        const bool fetch_now = dyn && chunk == 0 && tid == 0;
        if (fetch_now) fetched = next_id();                  // id of tile kt + 2; stored at the end of this item

        if (chunk == 0) {
#pragma unroll
            for (int i = 0; i < MAXKB16; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) dk_acc[i][c] = dv_acc[i][c] = 0.f;
            if constexpr (F32) {
                for (int i = tid; i < G::KP * HD; i += nthr) dKacc[i] = dVacc[i] = 0.f;
            }
        }

        const T* Ks = stage(s).K;
        const T* Vs = stage(s).V;
        const T* Qs = stage(s).Q;
        const T* dOs = stage(s).dO;
        const T* Os = stage(s).O;
        T* dSt = ALIAS_DS ? reinterpret_cast<T*>(stage(s).sG) : dSt_own;

        // ================= phase 1: per 16-row block: A, dA, dS, dQ
        // HB = false: rows +8..15 of the warp's 16-row block lie beyond L (the leftover block of L = 66 holds 2 rows): the
        // per-element work of that half is dropped; its A' / dS rows are stored as zeros (phase 2 reads whole blocks).
        auto row_block = [&](auto hb_tag) {
            constexpr bool HB = decltype(hb_tag)::value;
            const int la = warp * 16 + g, lb = la + 8;
            const int ra = row0 + la, rbb = row0 + lb;
            const TP* srow_a = stage(s).sS + la * G::STRIDE;
            const TP* srow_b = stage(s).sS + lb * G::STRIDE;
            TG* grow_a = stage(s).sG + la * G::STRIDE;
            TG* grow_b = stage(s).sG + lb * G::STRIDE;
            T* ap_a = Apt + la * G::STRIDE;
            T* ap_b = Apt + lb * G::STRIDE;
            T* ds_a = dSt + la * G::STRIDE;
            T* ds_b = dSt + lb * G::STRIDE;
            const bool va_ok = ra < L, vb_ok = HB && rbb < L;

            // ---- softmax recompute: sc[][] := exp(S - max)
            float sc[NKB][4];
            float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const int col = kb * 8 + 2 * q4;
                const float2 xa = slab_get2<TP>(srow_a, col);
                sc[kb][0] = xa.x; sc[kb][1] = xa.y;
                ma = fmaxf(ma, fmaxf(xa.x, xa.y));
                if constexpr (HB) {
                    const float2 xb = slab_get2<TP>(srow_b, col);
                    sc[kb][2] = xb.x; sc[kb][3] = xb.y;
                    mb = fmaxf(mb, fmaxf(xb.x, xb.y));
                } else {
                    sc[kb][2] = sc[kb][3] = 0.f;
                }
            }
            ma = quad_max(ma);
            if constexpr (HB) mb = quad_max(mb);
            if (!(ma > -INFINITY)) ma = 0.f;      // stale / fully masked rows: keep everything finite
            if (!(mb > -INFINITY)) mb = 0.f;
            float suma = 0.f, sumb = 0.f;
            if constexpr (F32) {
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    sc[kb][0] = expf(sc[kb][0] - ma); sc[kb][1] = expf(sc[kb][1] - ma);
                    if constexpr (HB) { sc[kb][2] = expf(sc[kb][2] - mb); sc[kb][3] = expf(sc[kb][3] - mb); }
                }
            } else {
                const float ka = ma * LOG2E, kbm = mb * LOG2E;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    sc[kb][0] = fast_ex2(fmaf(sc[kb][0], LOG2E, -ka));
                    sc[kb][1] = fast_ex2(fmaf(sc[kb][1], LOG2E, -ka));
                    if constexpr (HB) {
                        sc[kb][2] = fast_ex2(fmaf(sc[kb][2], LOG2E, -kbm));
                        sc[kb][3] = fast_ex2(fmaf(sc[kb][3], LOG2E, -kbm));
                    }
                }
            }
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                suma += sc[kb][0] + sc[kb][1];
                if constexpr (HB) sumb += sc[kb][2] + sc[kb][3];
            }
            suma = quad_sum(suma);
            if constexpr (HB) sumb = quad_sum(sumb);
            // rows beyond L (stale slab rows) get A = 0
            const float inva = (va_ok && suma > 0.f) ? (F32 ? 1.f / suma : fast_rcp(suma)) : 0.f;
            const float invb = (vb_ok && sumb > 0.f) ? (F32 ? 1.f / sumb : fast_rcp(sumb)) : 0.f;

            // ---- delta = dO . O per row; dO fragments
            float dela, delb;
            uint32_t ga0 = 0u, ga1 = 0u;
            float gfa[HD], gfb[HD];
            if constexpr (!F32) {
                const uint32_t* dO32 = reinterpret_cast<const uint32_t*>(dOs);
                const uint32_t* O32 = reinterpret_cast<const uint32_t*>(Os);
                ga0 = dO32[la * 4 + q4];
                const float2 x = unpack_bf16(ga0), y = unpack_bf16(O32[la * 4 + q4]);
                dela = x.x * y.x + x.y * y.y;
                delb = 0.f;
                if constexpr (HB) {
                    ga1 = dO32[lb * 4 + q4];
                    const float2 x2 = unpack_bf16(ga1), y2 = unpack_bf16(O32[lb * 4 + q4]);
                    delb = x2.x * y2.x + x2.y * y2.y;
                }
            } else {
                const float* dOf = reinterpret_cast<const float*>(dOs);
                const float* Of = reinterpret_cast<const float*>(Os);
#pragma unroll
                for (int d = 0; d < HD; ++d) { gfa[d] = dOf[la * HD + d]; gfb[d] = HB ? dOf[lb * HD + d] : 0.f; }
                dela = dOf[la * HD + 2 * q4] * Of[la * HD + 2 * q4] + dOf[la * HD + 2 * q4 + 1] * Of[la * HD + 2 * q4 + 1];
                delb = HB ? dOf[lb * HD + 2 * q4] * Of[lb * HD + 2 * q4] + dOf[lb * HD + 2 * q4 + 1] * Of[lb * HD + 2 * q4 + 1] : 0.f;
            }
            dela = quad_sum(dela);
            if constexpr (HB) delb = quad_sum(delb);

            // ---- per key block: dA' = dO V^T, A', dS; dS kept in sc[][] for dQ
            const uint32_t rkey = rng_stream_key(eff_seed, (uint32_t)tile);
            const uint32_t* Vs32 = reinterpret_cast<const uint32_t*>(Vs);
            uint2 qwa = make_uint2(0u, 0u), qwb = make_uint2(0u, 0u);
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const int col = kb * 8 + 2 * q4;
                float da[4];
                if constexpr (!F32) {
                    da[0] = da[1] = da[2] = da[3] = 0.f;
                    mma_bf16_1688(da, ga0, ga1, Vs32[(kb * 8 + g) * 4 + q4]);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* vr = reinterpret_cast<const float*>(Vs) + (kb * 8 + 2 * q4 + e) * HD;
                        float xa = 0.f, xb = 0.f;
#pragma unroll
                        for (int d = 0; d < HD; ++d) { xa = fmaf(gfa[d], vr[d], xa); xb = fmaf(gfb[d], vr[d], xb); }
                        da[e] = xa;
                        da[2 + e] = xb;
                    }
                }
                float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
                if (do_drop) {
                    if ((kb & 1) == 0) {
                        qwa = rng_quad_bits(rkey, ra, col);
                        if constexpr (HB) qwb = rng_quad_bits(rkey, rbb, col);
                    }
                    const uint32_t ba = (kb & 1) ? qwa.y : qwa.x, bb = (kb & 1) ? qwb.y : qwb.x;
                    const uint32_t thi = p.thresh16 << 16;
                    k0 = (ba << 16) >= thi ? p.keep_scale : 0.f;
                    k1 = ba >= thi ? p.keep_scale : 0.f;
                    if constexpr (HB) {
                        k2 = (bb << 16) >= thi ? p.keep_scale : 0.f;
                        k3 = bb >= thi ? p.keep_scale : 0.f;
                    }
                }
                const float A0 = sc[kb][0] * inva, A1 = sc[kb][1] * inva;
                if constexpr (F32) *reinterpret_cast<float2*>(ap_a + col) = make_float2(A0 * k0, A1 * k1);
                else *reinterpret_cast<uint32_t*>(ap_a + col) = pack_bf16(A0 * k0, A1 * k1);
                float2 ua = make_float2(0.f, 0.f);
                if (has_dpo) ua = slab_get2<TG>(grow_a, col);
                float d0 = fmaf(A0, fmaf(da[0], k0, -dela), ua.x);
                float d1 = fmaf(A1, fmaf(da[1], k1, -dela), ua.y);
                if (!va_ok) { d0 = 0.f; d1 = 0.f; }      // rows beyond L: stale slab rows
                // stored (rounded) values are what the previous layer sees; use them for dQ/dK too
                const float2 sa = slab_put2<TG>(grow_a, col, d0, d1);
                sc[kb][0] = sa.x; sc[kb][1] = sa.y;
                if constexpr (!ALIAS_DS) *reinterpret_cast<uint32_t*>(ds_a + col) = pack_bf16(sa.x, sa.y);
                if constexpr (HB) {
                    const float A2 = sc[kb][2] * invb, A3 = sc[kb][3] * invb;
                    if constexpr (F32) *reinterpret_cast<float2*>(ap_b + col) = make_float2(A2 * k2, A3 * k3);
                    else *reinterpret_cast<uint32_t*>(ap_b + col) = pack_bf16(A2 * k2, A3 * k3);
                    float2 ub = make_float2(0.f, 0.f);
                    if (has_dpo) ub = slab_get2<TG>(grow_b, col);
                    float d2 = fmaf(A2, fmaf(da[2], k2, -delb), ub.x);
                    float d3 = fmaf(A3, fmaf(da[3], k3, -delb), ub.y);
                    if (!vb_ok) { d2 = 0.f; d3 = 0.f; }
                    const float2 sb = slab_put2<TG>(grow_b, col, d2, d3);
                    sc[kb][2] = sb.x; sc[kb][3] = sb.y;
                    if constexpr (!ALIAS_DS) *reinterpret_cast<uint32_t*>(ds_b + col) = pack_bf16(sb.x, sb.y);
                } else {
                    // zero rows: with several chunks per tile the same local rows held real values one chunk earlier
                    if constexpr (F32) *reinterpret_cast<float2*>(ap_b + col) = make_float2(0.f, 0.f);
                    else *reinterpret_cast<uint32_t*>(ap_b + col) = 0u;
                    slab_put2<TG>(grow_b, col, 0.f, 0.f);
                    sc[kb][2] = sc[kb][3] = 0.f;
                    if constexpr (!ALIAS_DS) *reinterpret_cast<uint32_t*>(ds_b + col) = 0u;
                }
            }

            // ---- dQ = scale * dS K
            T* dqg = static_cast<T*>(p.dq) + (size_t)b * L * p.lddqkv + h * HD;
            if constexpr (!F32) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < G::NKB16; j += 2) {
                    uint32_t b0, b1, b2 = 0u, b3 = 0u;
                    if (j + 1 < G::NKB16) ldmatrix_x4_trans(b0, b1, b2, b3, Ks + (j * 16 + lane) * HD);
                    else ldmatrix_x2_trans(b0, b1, Ks + (j * 16 + (lane & 15)) * HD);
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        if (j + jj < G::NKB16) {
                            const int kb0 = 2 * (j + jj), kb1 = kb0 + 1;
                            const uint32_t a0 = pack_bf16(sc[kb0][0], sc[kb0][1]);
                            const uint32_t a1 = pack_bf16(sc[kb0][2], sc[kb0][3]);
                            uint32_t a2 = 0u, a3 = 0u;
                            if (kb1 < NKB) {
                                a2 = pack_bf16(sc[kb1][0], sc[kb1][1]);
                                a3 = pack_bf16(sc[kb1][2], sc[kb1][3]);
                            }
                            mma_bf16_16816(o, a0, a1, a2, a3, jj ? b2 : b0, jj ? b3 : b1);
                        }
                    }
                }
                if (va_ok)
                    *reinterpret_cast<uint32_t*>(dqg + (size_t)ra * p.lddqkv + 2 * q4) = pack_bf16(o[0] * p.scale, o[1] * p.scale);
                if (vb_ok)
                    *reinterpret_cast<uint32_t*>(dqg + (size_t)rbb * p.lddqkv + 2 * q4) = pack_bf16(o[2] * p.scale, o[3] * p.scale);
            } else {
                float oa[HD], ob[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) oa[d] = ob[d] = 0.f;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* kr = reinterpret_cast<const float*>(Ks) + (kb * 8 + 2 * q4 + e) * HD;
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            oa[d] = fmaf(sc[kb][e], kr[d], oa[d]);
                            ob[d] = fmaf(sc[kb][2 + e], kr[d], ob[d]);
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < HD; ++d) {
                    oa[d] = quad_sum(oa[d]) * p.scale;
                    ob[d] = quad_sum(ob[d]) * p.scale;
                }
                if (q4 == 0) {
                    if (va_ok) {
                        float4* dst = reinterpret_cast<float4*>(dqg + (size_t)ra * p.lddqkv);
                        dst[0] = make_float4(oa[0], oa[1], oa[2], oa[3]);
                        dst[1] = make_float4(oa[4], oa[5], oa[6], oa[7]);
                    }
                    if (vb_ok) {
                        float4* dst = reinterpret_cast<float4*>(dqg + (size_t)rbb * p.lddqkv);
                        dst[0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
                        dst[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
                    }
                }
            }
        };
        if constexpr (CS) {
            constexpr int NSPLIT = 4;
            constexpr int KBW = (NKB + NSPLIT - 1) / NSPLIT;          // key blocks per warp
            __shared__ float cs_max[2][NSPLIT][16];
            __shared__ float cs_sum[2][NSPLIT][16];
            __shared__ float cs_dq[2][NSPLIT - 1][16][HD];
            const int rb = warp >> 2, cs = warp & 3;                   // launched with 8 warps, crb = 2
            if (rb < nrb_chunk) {
                const int kb0 = cs * KBW;
                const int la = rb * 16 + g, lb = la + 8;
                const int ra = row0 + la, rbb = row0 + lb;
                const TP* srow_a = stage(s).sS + la * G::STRIDE;
                const TP* srow_b = stage(s).sS + lb * G::STRIDE;
                TG* grow_a = stage(s).sG + la * G::STRIDE;
                TG* grow_b = stage(s).sG + lb * G::STRIDE;
                T* ap_a = Apt + la * G::STRIDE;
                T* ap_b = Apt + lb * G::STRIDE;
                T* ds_a = dSt + la * G::STRIDE;
                T* ds_b = dSt + lb * G::STRIDE;
                const bool va_ok = ra < L, vb_ok = rbb < L;
                const unsigned bar_id = 1u + (unsigned)rb;

                // ---- softmax recompute over this warp's key blocks; row max and row sum across the 4 warps
                float sc[KBW][4];
                float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
                for (int j = 0; j < KBW; ++j) {
                    const int kb = kb0 + j;
                    if (kb < NKB) {
                        const int col = kb * 8 + 2 * q4;
                        const float2 xa = slab_get2<TP>(srow_a, col), xb = slab_get2<TP>(srow_b, col);
                        sc[j][0] = xa.x; sc[j][1] = xa.y; sc[j][2] = xb.x; sc[j][3] = xb.y;
                        ma = fmaxf(ma, fmaxf(xa.x, xa.y));
                        mb = fmaxf(mb, fmaxf(xb.x, xb.y));
                    } else {
                        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = -INFINITY;
                    }
                }
                ma = quad_max(ma);
                mb = quad_max(mb);
                if (q4 == 0) { cs_max[rb][cs][g] = ma; cs_max[rb][cs][g + 8] = mb; }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128) : "memory");
#pragma unroll
                for (int c = 0; c < NSPLIT; ++c) {
                    ma = fmaxf(ma, cs_max[rb][c][g]);
                    mb = fmaxf(mb, cs_max[rb][c][g + 8]);
                }
                if (!(ma > -INFINITY)) ma = 0.f;      // stale / fully masked rows: keep everything finite
                if (!(mb > -INFINITY)) mb = 0.f;
                float suma = 0.f, sumb = 0.f;
                {
                    const float ka = ma * LOG2E, kbm = mb * LOG2E;
#pragma unroll
                    for (int j = 0; j < KBW; ++j) {
                        sc[j][0] = fast_ex2(fmaf(sc[j][0], LOG2E, -ka));
                        sc[j][1] = fast_ex2(fmaf(sc[j][1], LOG2E, -ka));
                        sc[j][2] = fast_ex2(fmaf(sc[j][2], LOG2E, -kbm));
                        sc[j][3] = fast_ex2(fmaf(sc[j][3], LOG2E, -kbm));
                        suma += sc[j][0] + sc[j][1];
                        sumb += sc[j][2] + sc[j][3];
                    }
                }
                suma = quad_sum(suma);
                sumb = quad_sum(sumb);
                if (q4 == 0) { cs_sum[rb][cs][g] = suma; cs_sum[rb][cs][g + 8] = sumb; }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128) : "memory");
                suma = 0.f;
                sumb = 0.f;
#pragma unroll
                for (int c = 0; c < NSPLIT; ++c) {
                    suma += cs_sum[rb][c][g];
                    sumb += cs_sum[rb][c][g + 8];
                }
                const float inva = (va_ok && suma > 0.f) ? fast_rcp(suma) : 0.f;
                const float invb = (vb_ok && sumb > 0.f) ? fast_rcp(sumb) : 0.f;

                // ---- delta = dO . O per row (every warp of the row block evaluates it); dO fragments
                const uint32_t* dO32 = reinterpret_cast<const uint32_t*>(dOs);
                const uint32_t* O32 = reinterpret_cast<const uint32_t*>(Os);
                const uint32_t ga0 = dO32[la * 4 + q4], ga1 = dO32[lb * 4 + q4];
                float dela, delb;
                {
                    const float2 x = unpack_bf16(ga0), y = unpack_bf16(O32[la * 4 + q4]);
                    const float2 x2 = unpack_bf16(ga1), y2 = unpack_bf16(O32[lb * 4 + q4]);
                    dela = quad_sum(x.x * y.x + x.y * y.y);
                    delb = quad_sum(x2.x * y2.x + x2.y * y2.y);
                }

                // ---- per key block: dA' = dO V^T, A', dS; dS kept in sc[][] for dQ
                const uint32_t rkey = rng_stream_key(eff_seed, (uint32_t)tile);
                const uint32_t thi = p.thresh16 << 16;
                const uint32_t* Vs32 = reinterpret_cast<const uint32_t*>(Vs);
                uint2 qwa = make_uint2(0u, 0u), qwb = make_uint2(0u, 0u);
#pragma unroll
                for (int j = 0; j < KBW; ++j) {
                    const int kb = kb0 + j;
                    if (kb < NKB) {
                        const int col = kb * 8 + 2 * q4;
                        float da[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_bf16_1688(da, ga0, ga1, Vs32[(kb * 8 + g) * 4 + q4]);
                        float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
                        if (do_drop) {
                            // one hash serves the key-block pair (2m, 2m+1): .x even block, .y odd block
                            if (j == 0 || (kb & 1) == 0) { qwa = rng_quad_bits(rkey, ra, col); qwb = rng_quad_bits(rkey, rbb, col); }
                            const uint32_t ba = (kb & 1) ? qwa.y : qwa.x, bb = (kb & 1) ? qwb.y : qwb.x;
                            k0 = (ba << 16) >= thi ? p.keep_scale : 0.f;
                            k1 = ba >= thi ? p.keep_scale : 0.f;
                            k2 = (bb << 16) >= thi ? p.keep_scale : 0.f;
                            k3 = bb >= thi ? p.keep_scale : 0.f;
                        }
                        const float A0 = sc[j][0] * inva, A1 = sc[j][1] * inva, A2 = sc[j][2] * invb, A3 = sc[j][3] * invb;
                        *reinterpret_cast<uint32_t*>(ap_a + col) = pack_bf16(A0 * k0, A1 * k1);
                        *reinterpret_cast<uint32_t*>(ap_b + col) = pack_bf16(A2 * k2, A3 * k3);
                        float2 ua = make_float2(0.f, 0.f), ub = make_float2(0.f, 0.f);
                        if (has_dpo) {
                            ua = slab_get2<TG>(grow_a, col);
                            ub = slab_get2<TG>(grow_b, col);
                        }
                        float d0 = fmaf(A0, fmaf(da[0], k0, -dela), ua.x);
                        float d1 = fmaf(A1, fmaf(da[1], k1, -dela), ua.y);
                        float d2 = fmaf(A2, fmaf(da[2], k2, -delb), ub.x);
                        float d3 = fmaf(A3, fmaf(da[3], k3, -delb), ub.y);
                        if (!va_ok) { d0 = 0.f; d1 = 0.f; }      // rows beyond L: stale slab rows
                        if (!vb_ok) { d2 = 0.f; d3 = 0.f; }
                        const float2 sa = slab_put2<TG>(grow_a, col, d0, d1);
                        const float2 sb = slab_put2<TG>(grow_b, col, d2, d3);
                        sc[j][0] = sa.x; sc[j][1] = sa.y; sc[j][2] = sb.x; sc[j][3] = sb.y;
                        if constexpr (!ALIAS_DS) {
                            *reinterpret_cast<uint32_t*>(ds_a + col) = pack_bf16(sa.x, sa.y);
                            *reinterpret_cast<uint32_t*>(ds_b + col) = pack_bf16(sb.x, sb.y);
                        }
                    }
                }

                // ---- dQ = scale * dS K: partial over this warp's key blocks (pairs of 8-key blocks from kb0), summed by warp 0
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int jp = 0; jp < (KBW + 1) / 2; ++jp) {
                    const int kbA = kb0 + 2 * jp;
                    if (kbA < NKB) {
                        uint32_t b0, b1;
                        ldmatrix_x2_trans(b0, b1, Ks + (kbA * 8 + (lane & 15)) * HD);
                        const uint32_t a0 = pack_bf16(sc[2 * jp][0], sc[2 * jp][1]);
                        const uint32_t a1 = pack_bf16(sc[2 * jp][2], sc[2 * jp][3]);
                        uint32_t a2 = 0u, a3 = 0u;
                        if (2 * jp + 1 < KBW) {
                            if (kbA + 1 < NKB) {
                                a2 = pack_bf16(sc[2 * jp + 1][0], sc[2 * jp + 1][1]);
                                a3 = pack_bf16(sc[2 * jp + 1][2], sc[2 * jp + 1][3]);
                            }
                        }
                        mma_bf16_16816(o, a0, a1, a2, a3, b0, b1);
                    }
                }
                if (cs > 0) {
                    *reinterpret_cast<float2*>(&cs_dq[rb][cs - 1][g][2 * q4]) = make_float2(o[0], o[1]);
                    *reinterpret_cast<float2*>(&cs_dq[rb][cs - 1][g + 8][2 * q4]) = make_float2(o[2], o[3]);
                }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(128) : "memory");
                if (cs == 0) {
#pragma unroll
                    for (int c = 0; c < NSPLIT - 1; ++c) {
                        const float2 x = *reinterpret_cast<const float2*>(&cs_dq[rb][c][g][2 * q4]);
                        const float2 y = *reinterpret_cast<const float2*>(&cs_dq[rb][c][g + 8][2 * q4]);
                        o[0] += x.x; o[1] += x.y; o[2] += y.x; o[3] += y.y;
                    }
                    T* dqg = static_cast<T*>(p.dq) + (size_t)b * L * p.lddqkv + h * HD;
                    if (va_ok)
                        *reinterpret_cast<uint32_t*>(dqg + (size_t)ra * p.lddqkv + 2 * q4) = pack_bf16(o[0] * p.scale, o[1] * p.scale);
                    if (vb_ok)
                        *reinterpret_cast<uint32_t*>(dqg + (size_t)rbb * p.lddqkv + 2 * q4) = pack_bf16(o[2] * p.scale, o[3] * p.scale);
                }
            }
        } else if (warp < nrb_chunk) {
            if (row0 + warp * 16 + 8 < L) row_block(std::true_type{});
            else row_block(std::false_type{});
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(static_cast<TG*>(p.dpin) + (size_t)tile * tile_elems + (size_t)row0 * G::STRIDE, stage(s).sG,
                     (uint32_t)((size_t)nrows * G::STRIDE * sizeof(TG)));
            bulk_commit();
        }

        // ================= phase 2: dK += dS^T Q, dV += A'^T dO over this chunk's rows
        if constexpr (!F32) {
#pragma unroll
            for (int i = 0; i < MAXKB16; ++i) {
                const int kblk = warp + i * nwarps;
                if (kblk < nkb16) {
                    const int key0 = kblk * 16;
                    const bool hi_ok = key0 + 8 < G::KP;     // the second 8-key half may lie beyond the row stride
                    // ldmatrix.x4.trans, matrix m = lane/8:  m0 rows +0..7 keys key0..+7 | m1 rows +0..7 keys +8..
                    //                                        m2 rows +8..15 keys key0..  | m3 rows +8..15 keys +8..
                    const int m = lane >> 3, rr = lane & 7;
                    const int soff = (rr + ((m & 2) ? 8 : 0)) * G::STRIDE + key0 + (((m & 1) && hi_ok) ? 8 : 0);
                    for (int r = 0; r < nrb_chunk; ++r) {
                        const int lr0 = r * 16;
                        const int brow = lr0 + (lane & 15);
                        uint32_t a0, a1, a2, a3, b0, b1;
                        ldmatrix_x4_trans(a0, a1, a2, a3, dSt + lr0 * G::STRIDE + soff);
                        if (!hi_ok) { a1 = 0u; a3 = 0u; }
                        ldmatrix_x2_trans(b0, b1, Qs + brow * HD);
                        mma_bf16_16816(dk_acc[i], a0, a1, a2, a3, b0, b1);
                        ldmatrix_x4_trans(a0, a1, a2, a3, Apt + lr0 * G::STRIDE + soff);
                        if (!hi_ok) { a1 = 0u; a3 = 0u; }
                        ldmatrix_x2_trans(b0, b1, dOs + brow * HD);
                        mma_bf16_16816(dv_acc[i], a0, a1, a2, a3, b0, b1);
                    }
                }
            }
        } else {
            for (int idx = tid; idx < L * HD; idx += nthr) {
                const int key = idx >> 3, d = idx & 7;
                float xk = 0.f, xv = 0.f;
                for (int r = 0; r < nrows; ++r) {
                    xk = fmaf(reinterpret_cast<const float*>(dSt)[r * G::STRIDE + key], reinterpret_cast<const float*>(Qs)[r * HD + d], xk);
                    xv = fmaf(reinterpret_cast<const float*>(Apt)[r * G::STRIDE + key], reinterpret_cast<const float*>(dOs)[r * HD + d], xv);
                }
                dKacc[idx] += xk;
                dVacc[idx] += xv;
            }
        }

        // ---- last chunk of the tile: write dK, dV
        if (chunk == p.nchunks - 1) {
            T* dkg = static_cast<T*>(p.dk) + (size_t)b * L * p.lddqkv + h * HD;
            T* dvg = static_cast<T*>(p.dv) + (size_t)b * L * p.lddqkv + h * HD;
            if constexpr (!F32) {
#pragma unroll
                for (int i = 0; i < MAXKB16; ++i) {
                    const int kblk = warp + i * nwarps;
                    if (kblk < nkb16) {
                        const int ka = kblk * 16 + g, kb_ = ka + 8;
                        if (ka < L) {
                            *reinterpret_cast<uint32_t*>(dkg + (size_t)ka * p.lddqkv + 2 * q4) = pack_bf16(dk_acc[i][0] * p.scale, dk_acc[i][1] * p.scale);
                            *reinterpret_cast<uint32_t*>(dvg + (size_t)ka * p.lddqkv + 2 * q4) = pack_bf16(dv_acc[i][0], dv_acc[i][1]);
                        }
                        if (kb_ < L) {
                            *reinterpret_cast<uint32_t*>(dkg + (size_t)kb_ * p.lddqkv + 2 * q4) = pack_bf16(dk_acc[i][2] * p.scale, dk_acc[i][3] * p.scale);
                            *reinterpret_cast<uint32_t*>(dvg + (size_t)kb_ * p.lddqkv + 2 * q4) = pack_bf16(dv_acc[i][2], dv_acc[i][3]);
                        }
                    }
                }
            } else {
                // each (key,d) accumulator is owned by one thread: no sync needed
                for (int idx = tid; idx < L * HD; idx += nthr) {
                    const int key = idx >> 3, d = idx & 7;
                    dkg[(size_t)key * p.lddqkv + d] = dKacc[idx] * p.scale;
                    dvg[(size_t)key * p.lddqkv + d] = dVacc[idx];
                }
            }
        }
        // slot kt & 1 held this tile's id, which everybody read before the barrier at the top of this item
        if (fetch_now) s_tile[kt & 1] = fetched;
        if (nchunk == 0) ++kt;
        tile = ntile;
        chunk = nchunk;
    }
    if (tid == 0) {
        bulk_wait_all<0>();
        if (dyn) {
            // the last CTA to finish re-arms the counters for the next launch that uses this pair (CUDA-graph replays)
            __threadfence();
            if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; }
        }
    }
}

// ------------------------------------------------------------------ debug: keep mask (dense (B,H,L,L))
__global__ void dropout_mask_kernel(uint8_t* keep, int H, int L, uint32_t thresh16, unsigned long long seed,
                                    const unsigned long long* seed_off) {
    seed = rng_effective_seed(seed, seed_off);
    const int bh = blockIdx.x;
    const uint32_t rkey = rng_stream_key(seed, (uint32_t)bh);
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        const int r = e / L, c = e - r * L;
        keep[(size_t)bh * L * L + e] = rng_keep(rng_pair_bits(rkey, r, c), c, thresh16) ? 1 : 0;
    }
}

// ------------------------------------------------------------------ host-side dispatch
inline void drop_params(float p, uint32_t& thresh16, float& keep_scale) {
    double t = floor((double)p * 65536.0 + 0.5);
    if (t < 0) t = 0;
    if (t > 65535) t = 65535;
    thresh16 = (uint32_t)t;
    keep_scale = (float)(65536.0 / (65536.0 - t));
}

constexpr size_t SMEM_CAP = 220 * 1024;

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

// rows per chunk: the whole tile when it has at most 8 sixteen-row blocks, else balanced chunks of <= 3 blocks
inline void pick_chunks(int L, int& crb, int& nchunks) {
    const int nrb = (L + 15) / 16;
    if (nrb <= 8) { crb = nrb; nchunks = 1; return; }
    nchunks = (nrb + 2) / 3;
    crb = (nrb + nchunks - 1) / nchunks;
    nchunks = (nrb + crb - 1) / crb;
}

template <typename T, typename TP, int NKB>
int launch_fwd(FwdParams p, cudaStream_t st) {
    pick_chunks(p.L, p.crb, p.nchunks);
    auto smem_of = [](int crb) { return NSTAGE_F * ((FwdStage<T, TP, NKB>::bytes(crb * 16) + 127) & ~size_t(127)); };
    while (p.crb > 1 && smem_of(p.crb) > SMEM_CAP) --p.crb;
    p.nchunks = ((p.L + 15) / 16 + p.crb - 1) / p.crb;
    if constexpr (!std::is_same<T, float>::value && NKB >= 25) {
        // column-split form (see the kernel): 8 warps, 32-row chunks, two CTAs per SM
        const char* cs_env = getenv("MMDTI_K2_FWD_CS");       // read per launch so that a test can compare both forms
        const bool use_cs = cs_env ? atoi(cs_env) != 0 : K2_FWD_CS_DEFAULT;
        if (use_cs && smem_of(2) <= 112 * 1024) {
            p.crb = 2;
            p.nchunks = ((p.L + 15) / 16 + 1) / 2;
            const size_t smem = smem_of(2);
            auto kern = pair_attn_fwd_kernel<T, TP, NKB, true>;
            static int occ_dev[MMDTI_MAX_DEVICES] = {};      // per device: the attribute and the occupancy are device properties
            int& occ_cs = occ_dev[mmdti_device_slot()];
            if (!occ_cs) {
                MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                MMDTI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cs, kern, 256, smem));
                if (occ_cs < 1) occ_cs = 1;
            }
            const long long ntiles = (long long)p.B * p.H;
            const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * occ_cs);
            kern<<<grid, 256, smem, st>>>(p);
            MMDTI_LAUNCH_OK();
            return MMDTI_OK;
        }
    }
    const size_t smem = smem_of(p.crb);
    MMDTI_REQUIRE(smem <= SMEM_CAP, "pair_attn_fwd: shared memory %zu exceeds cap (L=%d)", smem, p.L);
    auto kern = pair_attn_fwd_kernel<T, TP, NKB>;
    static int occ_dev[MMDTI_MAX_DEVICES] = {};
    static size_t occ_smem_dev[MMDTI_MAX_DEVICES] = {};
    const int dev_slot = mmdti_device_slot();
    int& occ = occ_dev[dev_slot];
    size_t& occ_smem = occ_smem_dev[dev_slot];
    const int threads = p.crb * 32;
    if (!occ || occ_smem != smem) {
        MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MMDTI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        if (occ < 1) occ = 1;
        occ_smem = smem;
    }
    const long long ntiles = (long long)p.B * p.H;
    const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * occ);
    kern<<<grid, threads, smem, st>>>(p);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

template <typename T, typename TP, typename TG, int NKB>
int launch_bwd(BwdParams p, cudaStream_t st) {
    using G = Geo<NKB>;
    pick_chunks(p.L, p.crb, p.nchunks);
    constexpr bool F32 = std::is_same<T, float>::value;
    constexpr bool ALIAS_DS = std::is_same<T, TG>::value;
    auto smem_of = [](int crb) {
        const int NR = crb * 16;
        size_t s = NSTAGE * ((BwdStage<T, TP, TG, NKB>::bytes(NR) + 127) & ~size_t(127)) +
                   (size_t)NR * G::STRIDE * sizeof(T) * (ALIAS_DS ? 1 : 2);
        if (F32) s += 2 * (size_t)G::KP * HD * sizeof(float);
        return s;
    };
    while (p.crb > 1 && smem_of(p.crb) > SMEM_CAP) --p.crb;
    // two resident CTAs beat one with a longer chunk (L = 258: 3 warps/SM -> 6, 722 -> 576 us)
    if (p.nchunks > 1)
        while (p.crb > 2 && smem_of(p.crb) > 113 * 1024) --p.crb;
    if (const char* e = getenv("MMDTI_K2_BWD_CRB")) p.crb = std::max(1, std::min(p.crb, atoi(e)));      // tuning knob
    // dynamic tile scheduler (see the kernel): MMDTI_K2_BWD_DYN = fraction of the tiles in the dynamic pool, 0 .. 1.  Default 0
    // (static contiguous ranges: fastest when nothing else runs on the SMs).  Measured at N = 4 under the overlapped NCCL
    // all-reduce: 0.3 brings the launch from 203 to 141 us, but the step stays at 9.6 ms (the all-reduce, not K2, is what the
    // backward waits for), so it stays off.
    {
        const char* e = getenv("MMDTI_K2_BWD_DYN");
        const double frac = e ? std::min(1.0, std::max(0.0, atof(e))) : 0.0;
        p.pool = (int)(frac * (double)p.B * (double)p.H);
        p.sched = p.pool > 0 ? mmdti_sched_slot() : nullptr;
    }
    const int nkb16 = (p.L + 15) / 16;
    if constexpr (!F32 && NKB >= 25) {
        // column-split phase 1 (see the kernel): 8 warps, 32-row chunks, two CTAs per SM
        const char* cs_env = getenv("MMDTI_K2_BWD_CS");       // read per launch so that a test can compare both forms
        const bool use_cs = cs_env ? atoi(cs_env) != 0 : K2_BWD_CS_DEFAULT;
        if (use_cs && smem_of(2) <= 112 * 1024) {
            p.crb = 2;
            p.nchunks = (nkb16 + 1) / 2;
            const size_t smem = smem_of(2);
            auto kern = pair_attn_bwd_kernel<T, TP, TG, NKB, true>;
            static int occ_dev[MMDTI_MAX_DEVICES] = {};      // per device: the attribute and the occupancy are device properties
            int& occ_cs = occ_dev[mmdti_device_slot()];
            if (!occ_cs) {
                MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                MMDTI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cs, kern, 256, smem));
                if (occ_cs < 1) occ_cs = 1;
            }
            const long long ntiles = (long long)p.B * p.H;
            const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * occ_cs);
            kern<<<grid, 256, smem, st>>>(p);
            MMDTI_LAUNCH_OK();
            return MMDTI_OK;
        }
    }
    p.nchunks = (nkb16 + p.crb - 1) / p.crb;
    // phase 1 uses one warp per 16-row block of the chunk; phase 2 spreads the 16-key blocks over all warps
    const int nwarps = std::max(p.crb, (nkb16 + G::MAXKB16 - 1) / G::MAXKB16);
    MMDTI_REQUIRE(nwarps <= 8, "pair_attn_bwd: unsupported L=%d", p.L);
    const size_t smem = smem_of(p.crb);
    MMDTI_REQUIRE(smem <= SMEM_CAP, "pair_attn_bwd: shared memory %zu exceeds cap (L=%d)", smem, p.L);
    auto kern = pair_attn_bwd_kernel<T, TP, TG, NKB>;
    static int occ_dev[MMDTI_MAX_DEVICES] = {};
    static size_t occ_smem_dev[MMDTI_MAX_DEVICES] = {};
    const int dev_slot = mmdti_device_slot();
    int& occ = occ_dev[dev_slot];
    size_t& occ_smem = occ_smem_dev[dev_slot];
    const int threads = nwarps * 32;
    if (!occ || occ_smem != smem) {
        MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MMDTI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        if (occ < 1) occ = 1;
        occ_smem = smem;
    }
    const long long ntiles = (long long)p.B * p.H;
    const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * occ);
    kern<<<grid, threads, smem, st>>>(p);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

template <typename T, typename TP>
int dispatch_fwd_nkb(const FwdParams& p, cudaStream_t st) {
    switch (mmdti_pair_nkb(p.L)) {
        case 3: return launch_fwd<T, TP, 3>(p, st);
        case 5: return launch_fwd<T, TP, 5>(p, st);
        case 9: return launch_fwd<T, TP, 9>(p, st);
        case 13: return launch_fwd<T, TP, 13>(p, st);
        case 17: return launch_fwd<T, TP, 17>(p, st);
        case 25: return launch_fwd<T, TP, 25>(p, st);
        case 33: return launch_fwd<T, TP, 33>(p, st);
    }
    mmdti_set_error("pair_attn_fwd: L=%d exceeds the supported maximum 264", p.L);
    return MMDTI_ERR_ARG;
}
template <typename T, typename TP, typename TG>
int dispatch_bwd_nkb(const BwdParams& p, cudaStream_t st) {
    switch (mmdti_pair_nkb(p.L)) {
        case 3: return launch_bwd<T, TP, TG, 3>(p, st);
        case 5: return launch_bwd<T, TP, TG, 5>(p, st);
        case 9: return launch_bwd<T, TP, TG, 9>(p, st);
        case 13: return launch_bwd<T, TP, TG, 13>(p, st);
        case 17: return launch_bwd<T, TP, TG, 17>(p, st);
        case 25: return launch_bwd<T, TP, TG, 25>(p, st);
        case 33: return launch_bwd<T, TP, TG, 33>(p, st);
    }
    mmdti_set_error("pair_attn_bwd: L=%d exceeds the supported maximum 264", p.L);
    return MMDTI_ERR_ARG;
}

int check_common(const void* q, const void* k, const void* v, long long ld, int B, int H, int L, int act_dtype) {
    MMDTI_REQUIRE(B > 0 && H > 0 && L > 0, "pair_attn: empty problem B=%d H=%d L=%d", B, H, L);
    MMDTI_REQUIRE(act_dtype == MMDTI_F32 || act_dtype == MMDTI_BF16, "pair_attn: act_dtype must be f32 or bf16");
    MMDTI_REQUIRE(mmdti_pair_nkb(L) > 0, "pair_attn: L=%d exceeds the supported maximum 264", L);
    const size_t esz = act_dtype == MMDTI_F32 ? 4 : 2;
    MMDTI_REQUIRE(q && k && v && mmdti_aligned(q, 16) && mmdti_aligned(k, 16) && mmdti_aligned(v, 16) && (ld * esz) % 16 == 0,
                  "pair_attn: q/k/v must be 16-byte aligned with a 16-byte-multiple row stride");
    return MMDTI_OK;
}

}  // namespace

extern "C" int mmdti_pair_ld(int L) {
    const int nkb = mmdti_pair_nkb(L);
    return nkb > 0 ? nkb * 8 : -1;
}

extern "C" int mmdti_pair_attn_fwd(const void* q, const void* k, const void* v, int64_t ldqkv, const void* pair_in,
                                   void* pair_out, void* o, int64_t ldo, int B, int H, int L, float scale,
                                   float dropout_p, uint64_t seed, int act_dtype, int pair_dtype, void* stream) {
    if (int rc = check_common(q, k, v, ldqkv, B, H, L, act_dtype)) return rc;
    MMDTI_REQUIRE(pair_in && pair_out && o, "pair_attn_fwd: null buffer");
    MMDTI_REQUIRE(mmdti_aligned(pair_in, 16) && mmdti_aligned(pair_out, 16) && mmdti_aligned(o, 16),
                  "pair_attn_fwd: pair/o buffers must be 16-byte aligned");
    MMDTI_REQUIRE((ldo * (act_dtype == MMDTI_F32 ? 4 : 2)) % 16 == 0, "pair_attn_fwd: ldo must be a 16-byte multiple");
    MMDTI_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "pair_attn_fwd: dropout_p out of range");
    FwdParams p;
    p.q = q; p.k = k; p.v = v; p.o = o; p.pin = pair_in; p.pout = pair_out;
    p.ldqkv = ldqkv; p.ldo = ldo; p.B = B; p.H = H; p.L = L; p.scale = scale; p.seed = seed; p.seed_off = mmdti_seed_offset_ptr();
    p.crb = 0; p.nchunks = 0;
    drop_params(dropout_p, p.thresh16, p.keep_scale);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == MMDTI_F32) {
        MMDTI_REQUIRE(pair_dtype == MMDTI_F32, "pair_attn_fwd: f32 activations require an f32 pair tensor");
        return dispatch_fwd_nkb<float, float>(p, st);
    }
    switch (pair_dtype) {
        case MMDTI_BF16: return dispatch_fwd_nkb<bf16, bf16>(p, st);
        case MMDTI_F16: return dispatch_fwd_nkb<bf16, __half>(p, st);
        case MMDTI_F32: return dispatch_fwd_nkb<bf16, float>(p, st);
    }
    mmdti_set_error("pair_attn_fwd: bad pair_dtype %d", pair_dtype);
    return MMDTI_ERR_ARG;
}

extern "C" int mmdti_pair_attn_bwd(const void* q, const void* k, const void* v, int64_t ldqkv, const void* s,
                                   const void* o, const void* d_o, int64_t lddo, const void* d_pair_out, void* d_pair_in,
                                   void* dq, void* dk, void* dv, int64_t lddqkv, int B, int H, int L, float scale,
                                   float dropout_p, uint64_t seed, int act_dtype, int pair_dtype, int gpair_dtype,
                                   void* stream) {
    if (int rc = check_common(q, k, v, ldqkv, B, H, L, act_dtype)) return rc;
    MMDTI_REQUIRE(s && o && d_o && d_pair_in && dq && dk && dv, "pair_attn_bwd: null buffer");
    const size_t esz = act_dtype == MMDTI_F32 ? 4 : 2;
    MMDTI_REQUIRE(mmdti_aligned(s, 16) && mmdti_aligned(o, 16) && mmdti_aligned(d_o, 16) && mmdti_aligned(d_pair_in, 16) &&
                      mmdti_aligned(d_pair_out, 16) && mmdti_aligned(dq, 16) && mmdti_aligned(dk, 16) &&
                      mmdti_aligned(dv, 16) && (lddo * esz) % 16 == 0 && (lddqkv * esz) % 16 == 0,
                  "pair_attn_bwd: buffers must be 16-byte aligned with 16-byte-multiple row strides");
    MMDTI_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "pair_attn_bwd: dropout_p out of range");
    BwdParams p;
    p.q = q; p.k = k; p.v = v; p.s = s; p.o = o; p.d_o = d_o; p.dpout = d_pair_out; p.dpin = d_pair_in;
    p.dq = dq; p.dk = dk; p.dv = dv; p.ldqkv = ldqkv; p.lddo = lddo; p.lddqkv = lddqkv;
    p.B = B; p.H = H; p.L = L; p.scale = scale; p.seed = seed; p.seed_off = mmdti_seed_offset_ptr();
    p.crb = 0; p.nchunks = 0;
    drop_params(dropout_p, p.thresh16, p.keep_scale);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (act_dtype == MMDTI_F32) {
        MMDTI_REQUIRE(pair_dtype == MMDTI_F32 && gpair_dtype == MMDTI_F32,
                      "pair_attn_bwd: f32 activations require f32 pair tensors");
        return dispatch_bwd_nkb<float, float, float>(p, st);
    }
    if (pair_dtype == MMDTI_BF16 && gpair_dtype == MMDTI_BF16) return dispatch_bwd_nkb<bf16, bf16, bf16>(p, st);
    if (pair_dtype == MMDTI_F16 && gpair_dtype == MMDTI_F16) return dispatch_bwd_nkb<bf16, __half, __half>(p, st);
    if (pair_dtype == MMDTI_F32 && gpair_dtype == MMDTI_F32) return dispatch_bwd_nkb<bf16, float, float>(p, st);
    mmdti_set_error("pair_attn_bwd: unsupported (pair_dtype=%d, gpair_dtype=%d) combination", pair_dtype, gpair_dtype);
    return MMDTI_ERR_ARG;
}

extern "C" int mmdti_pair_attn_dropout_mask(uint8_t* keep, int B, int H, int L, float dropout_p, uint64_t seed,
                                            void* stream) {
    MMDTI_REQUIRE(keep && B > 0 && H > 0 && L > 0, "dropout_mask: bad arguments");
    uint32_t thresh16;
    float ks;
    drop_params(dropout_p, thresh16, ks);
    dropout_mask_kernel<<<B * H, 256, 0, static_cast<cudaStream_t>(stream)>>>(keep, H, L, thresh16, seed, mmdti_seed_offset_ptr());
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
