// K2 — pair-biased multi-head self-attention (head_dim 8), forward and backward.
//
// One CTA per (molecule b, head h).  The (L,L) pair tile of that head is one contiguous
// L*L-element range of the (B,H,L,L) tensor, so it is staged through shared memory with
// fully coalesced flat copies (global -> smem rows padded to a bank-conflict-free stride),
// consumed / produced in the register layout of mma.sync m16n8k8 / m16n8k16 (head_dim 8 is
// exactly the K of m16n8k8 for Q K^T and the N of m16n8k16 for A V), and written back
// with the same flat coalesced copy.  The kernel is bound by the pair-tensor HBM traffic
// (read P + write P' forward; read S, read dP', write dP backward): 32 flop per 4..8 bytes.
//
// Reference semantics: Uni-Core SelfMultiheadAttention(return_attn=True) as driven by
// models/transformers.py:136-139 (see include/mmdti_b200.h).
#include "common.cuh"

#include <math.h>

namespace {

constexpr int HD = MMDTI_HEAD_DIM;  // 8

// row stride (elements) of the smem pair slabs: multiple of 8, and == 8 (mod 16) so that
// (a) 32-bit fragment accesses of 8 rows x 4 column pairs hit 32 distinct banks and
// (b) ldmatrix rows (16 B) of 8 consecutive slab rows hit 8 distinct 16-byte bank groups.
template <int NKB> struct Geo {
    static constexpr int KP = NKB * 8;                                // padded key count
    static constexpr int STRIDE = (KP % 16 == 8) ? KP : KP + 8;
    static constexpr int NKB16 = (NKB + 1) / 2;
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
    int crb;   // 16-row blocks staged per chunk (<= warps per CTA)
};

struct BwdParams {
    const void *q, *k, *v, *o, *s, *d_o, *dpout;
    void *dpin, *dq, *dk, *dv;
    long long ldqkv, lddo, lddqkv;
    int B, H, L;
    float scale, keep_scale;
    uint32_t thresh16;
    unsigned long long seed;
    int crb;
};

// ------------------------------------------------------------------ slab copies
// global (flat, contiguous nrows*L elements) <-> smem [nrows][STRIDE]
template <typename TP, int STRIDE>
__device__ __forceinline__ void slab_load(TP* __restrict__ slab, const TP* __restrict__ g, int nrows, int L,
                                          int tid, int nthr) {
    if (sizeof(TP) == 2 && (L & 1) == 0) {
        const uint32_t* g32 = reinterpret_cast<const uint32_t*>(g);
        uint32_t* s32 = reinterpret_cast<uint32_t*>(slab);
        const int Lh = L >> 1, n = nrows * Lh;
        const FastDiv fd(Lh);
#pragma unroll 4
        for (int e = tid; e < n; e += nthr) {
            const int r = fd.div(e), c = e - r * Lh;
            s32[r * (STRIDE / 2) + c] = __ldg(g32 + e);
        }
    } else {
        const int n = nrows * L;
        const FastDiv fd(L);
#pragma unroll 4
        for (int e = tid; e < n; e += nthr) {
            const int r = fd.div(e), c = e - r * L;
            slab[r * STRIDE + c] = g[e];
        }
    }
}
template <typename TP, int STRIDE>
__device__ __forceinline__ void slab_store(TP* __restrict__ g, const TP* __restrict__ slab, int nrows, int L,
                                           int tid, int nthr) {
    if (sizeof(TP) == 2 && (L & 1) == 0) {
        uint32_t* g32 = reinterpret_cast<uint32_t*>(g);
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(slab);
        const int Lh = L >> 1, n = nrows * Lh;
        const FastDiv fd(Lh);
#pragma unroll 4
        for (int e = tid; e < n; e += nthr) {
            const int r = fd.div(e), c = e - r * Lh;
            g32[e] = s32[r * (STRIDE / 2) + c];
        }
    } else {
        const int n = nrows * L;
        const FastDiv fd(L);
#pragma unroll 4
        for (int e = tid; e < n; e += nthr) {
            const int r = fd.div(e), c = e - r * L;
            g[e] = slab[r * STRIDE + c];
        }
    }
}

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

template <typename T> __device__ __forceinline__ void load_head_row(float (&dst)[HD], const T* p) {
    if constexpr (std::is_same<T, float>::value) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
        dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        float2 f;
        f = unpack_bf16(u.x); dst[0] = f.x; dst[1] = f.y;
        f = unpack_bf16(u.y); dst[2] = f.x; dst[3] = f.y;
        f = unpack_bf16(u.z); dst[4] = f.x; dst[5] = f.y;
        f = unpack_bf16(u.w); dst[6] = f.x; dst[7] = f.y;
    }
}

// ===================================================================== forward
// smem layout (T = bf16):  Ks [KP][8] bf16 | Vt [8][STRIDE] bf16 | slab [NW*16][STRIDE] TP
//             (T = float): Ks [KP][8] f32  | Vs [KP][8] f32      | slab
template <typename T, typename TP, int NKB>
__global__ void __launch_bounds__(256) pair_attn_fwd_kernel(const FwdParams p) {
    using G = Geo<NKB>;
    constexpr bool F32 = std::is_same<T, float>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Ks = reinterpret_cast<T*>(smem_raw);
    T* Vx = Ks + G::KP * HD;                                            // Vt (bf16) or Vs (f32)
    constexpr int VX_ELEMS = F32 ? G::KP * HD : HD * G::STRIDE;
    TP* slab = reinterpret_cast<TP*>(Vx + VX_ELEMS);

    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = nthr >> 5;
    const int g = lane >> 2, q4 = lane & 3;
    const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
    const int L = p.L;
    const T* qg = static_cast<const T*>(p.q) + (size_t)b * L * p.ldqkv + h * HD;
    const T* kg = static_cast<const T*>(p.k) + (size_t)b * L * p.ldqkv + h * HD;
    const T* vg = static_cast<const T*>(p.v) + (size_t)b * L * p.ldqkv + h * HD;
    T* og = static_cast<T*>(p.o) + (size_t)b * L * p.ldo + h * HD;
    const TP* pin = static_cast<const TP*>(p.pin) + (size_t)bh * L * L;
    TP* pout = static_cast<TP*>(p.pout) + (size_t)bh * L * L;

    // ---- stage K (row-major) and V (transposed for the bf16 path) of this head
    for (int key = tid; key < G::KP; key += nthr) {
        float kr[HD], vr[HD];
        if (key < L) {
            load_head_row(kr, kg + (size_t)key * p.ldqkv);
            load_head_row(vr, vg + (size_t)key * p.ldqkv);
        } else {
#pragma unroll
            for (int d = 0; d < HD; ++d) kr[d] = vr[d] = 0.f;
        }
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            Ks[key * HD + d] = from_f<T>(kr[d]);
            if constexpr (F32) Vx[key * HD + d] = vr[d];
            else Vx[d * G::STRIDE + key] = from_f<T>(vr[d]);
        }
    }
    if constexpr (!F32 && (G::STRIDE > G::KP)) {   // zero the stride padding of Vt (never read; kept clean)
        constexpr int PADC = G::STRIDE - G::KP;
        for (int i = tid; i < HD * PADC; i += nthr) {
            const int d = i / PADC, c = i - d * PADC;
            Vx[d * G::STRIDE + G::KP + c] = from_f<T>(0.f);
        }
    }

    const uint32_t rkey = rng_stream_key(p.seed, (uint32_t)bh);
    const bool do_drop = p.thresh16 != 0;
    const int nrb = (L + 15) >> 4;

    for (int rb0 = 0; rb0 < nrb; rb0 += p.crb) {
        const int row0 = rb0 * 16;
        const int nrows = min(L - row0, p.crb * 16);
        __syncthreads();
        slab_load<TP, G::STRIDE>(slab, pin + (size_t)row0 * L, nrows, L, tid, nthr);
        __syncthreads();

        const int rb = rb0 + warp;
        if (warp < p.crb && rb < nrb) {
            const int ra = rb * 16 + g, rbb = ra + 8;            // global query rows of this thread
            TP* srow_a = slab + (ra - row0) * G::STRIDE;
            TP* srow_b = srow_a + 8 * G::STRIDE;
            float s[NKB][4];

            // ---- S = Q K^T
            if constexpr (!F32) {
                const uint32_t qa0 = ra < L ? *reinterpret_cast<const uint32_t*>(qg + (size_t)ra * p.ldqkv + 2 * q4) : 0u;
                const uint32_t qa1 = rbb < L ? *reinterpret_cast<const uint32_t*>(qg + (size_t)rbb * p.ldqkv + 2 * q4) : 0u;
                const uint32_t* Ks32 = reinterpret_cast<const uint32_t*>(Ks);
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
                    s[kb][0] = s[kb][1] = s[kb][2] = s[kb][3] = 0.f;
                    mma_bf16_1688(s[kb], qa0, qa1, Ks32[(kb * 8 + g) * 4 + q4]);
                }
            } else {
                float qa[HD], qb[HD];
                if (ra < L) load_head_row(qa, qg + (size_t)ra * p.ldqkv);
                else {
#pragma unroll
                    for (int d = 0; d < HD; ++d) qa[d] = 0.f;
                }
                if (rbb < L) load_head_row(qb, qg + (size_t)rbb * p.ldqkv);
                else {
#pragma unroll
                    for (int d = 0; d < HD; ++d) qb[d] = 0.f;
                }
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
                        s[kb][e] = da;
                        s[kb][2 + e] = db;
                    }
                }
            }

            // ---- S = scale*S + P ; P' := S (stored, rounded to TP) ; running row max
            float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const int col = kb * 8 + 2 * q4;
                const float2 pa = slab_get2<TP>(srow_a, col), pb = slab_get2<TP>(srow_b, col);
                float2 va = slab_put2<TP>(srow_a, col, fmaf(s[kb][0], p.scale, pa.x), fmaf(s[kb][1], p.scale, pa.y));
                float2 vb = slab_put2<TP>(srow_b, col, fmaf(s[kb][2], p.scale, pb.x), fmaf(s[kb][3], p.scale, pb.y));
                if (col >= L) { va.x = -INFINITY; vb.x = -INFINITY; }
                if (col + 1 >= L) { va.y = -INFINITY; vb.y = -INFINITY; }
                s[kb][0] = va.x; s[kb][1] = va.y; s[kb][2] = vb.x; s[kb][3] = vb.y;
                ma = fmaxf(ma, fmaxf(va.x, va.y));
                mb = fmaxf(mb, fmaxf(vb.x, vb.y));
            }
            ma = quad_max(ma);
            mb = quad_max(mb);

            // ---- softmax numerators, row sums, dropout
            float suma = 0.f, sumb = 0.f;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                if constexpr (F32) {
                    s[kb][0] = expf(s[kb][0] - ma); s[kb][1] = expf(s[kb][1] - ma);
                    s[kb][2] = expf(s[kb][2] - mb); s[kb][3] = expf(s[kb][3] - mb);
                } else {
                    s[kb][0] = __expf(s[kb][0] - ma); s[kb][1] = __expf(s[kb][1] - ma);
                    s[kb][2] = __expf(s[kb][2] - mb); s[kb][3] = __expf(s[kb][3] - mb);
                }
                suma += s[kb][0] + s[kb][1];
                sumb += s[kb][2] + s[kb][3];
                if (do_drop) {
                    const int col = kb * 8 + 2 * q4;
                    const uint32_t ba = rng_pair_bits(rkey, ra, col), bb = rng_pair_bits(rkey, rbb, col);
                    if (!rng_keep(ba, 0, p.thresh16)) s[kb][0] = 0.f;
                    if (!rng_keep(ba, 1, p.thresh16)) s[kb][1] = 0.f;
                    if (!rng_keep(bb, 0, p.thresh16)) s[kb][2] = 0.f;
                    if (!rng_keep(bb, 1, p.thresh16)) s[kb][3] = 0.f;
                }
            }
            suma = quad_sum(suma);
            sumb = quad_sum(sumb);
            const float inva = p.keep_scale / suma, invb = p.keep_scale / sumb;

            // ---- O = A' V
            if constexpr (!F32) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t* Vt32 = reinterpret_cast<const uint32_t*>(Vx);
#pragma unroll
                for (int j = 0; j < G::NKB16; ++j) {
                    const int kb0 = 2 * j, kb1 = 2 * j + 1;
                    const uint32_t a0 = pack_bf16(s[kb0][0], s[kb0][1]);
                    const uint32_t a1 = pack_bf16(s[kb0][2], s[kb0][3]);
                    uint32_t a2 = 0u, a3 = 0u, b1 = 0u;
                    const uint32_t b0 = Vt32[(g * G::STRIDE + kb0 * 8 + 2 * q4) >> 1];
                    if (kb1 < NKB) {
                        a2 = pack_bf16(s[kb1][0], s[kb1][1]);
                        a3 = pack_bf16(s[kb1][2], s[kb1][3]);
                        b1 = Vt32[(g * G::STRIDE + kb1 * 8 + 2 * q4) >> 1];
                    }
                    mma_bf16_16816(o, a0, a1, a2, a3, b0, b1);
                }
                if (ra < L)
                    *reinterpret_cast<uint32_t*>(og + (size_t)ra * p.ldo + 2 * q4) = pack_bf16(o[0] * inva, o[1] * inva);
                if (rbb < L)
                    *reinterpret_cast<uint32_t*>(og + (size_t)rbb * p.ldo + 2 * q4) = pack_bf16(o[2] * invb, o[3] * invb);
            } else {
                float oa[HD], ob[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) oa[d] = ob[d] = 0.f;
#pragma unroll
                for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float* vr = Vx + (kb * 8 + 2 * q4 + e) * HD;
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            oa[d] = fmaf(s[kb][e], vr[d], oa[d]);
                            ob[d] = fmaf(s[kb][2 + e], vr[d], ob[d]);
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < HD; ++d) {
                    oa[d] = quad_sum(oa[d]) * inva;
                    ob[d] = quad_sum(ob[d]) * invb;
                }
                if (q4 == 0) {
                    if (ra < L) {
                        float4* dst = reinterpret_cast<float4*>(og + (size_t)ra * p.ldo);
                        dst[0] = make_float4(oa[0], oa[1], oa[2], oa[3]);
                        dst[1] = make_float4(oa[4], oa[5], oa[6], oa[7]);
                    }
                    if (rbb < L) {
                        float4* dst = reinterpret_cast<float4*>(og + (size_t)rbb * p.ldo);
                        dst[0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
                        dst[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
                    }
                }
            }
        }
        __syncthreads();
        slab_store<TP, G::STRIDE>(pout + (size_t)row0 * L, slab, nrows, L, tid, nthr);
    }
}

// ===================================================================== backward
// smem layout, T = bf16 (NR = crb*16 rows per chunk, LR = 16*ceil(L/16) rows):
//   Kt [8][STRIDE] bf16 | Vs [KP][8] bf16 | Qs [LR][8] bf16 | dOs [LR][8] bf16 |
//   sS [NR][STRIDE] TP | sG [NR][STRIDE] TG | dSt [NR][STRIDE] bf16 | Apt [NR][STRIDE] bf16
// T = float: Ks,Vs,Qs,dOs f32 [..][8]; dSt/Apt f32; dKacc,dVacc [KP][8] f32 at the end.
// delta_i = rowsum(dA o A) is taken as dO_i . O_i (O from the forward), so a row's dA never
// has to be resident all at once.
template <typename T, typename TP, typename TG, int NKB>
__global__ void __launch_bounds__(256) pair_attn_bwd_kernel(const BwdParams p) {
    using G = Geo<NKB>;
    constexpr bool F32 = std::is_same<T, float>::value;
    constexpr int MAXKB16 = 3;   // 16-key blocks per warp in phase 2 (host: ceil(nrb/nwarps) <= 3)
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = nthr >> 5;
    const int g = lane >> 2, q4 = lane & 3;
    const int bh = blockIdx.x, b = bh / p.H, h = bh - b * p.H;
    const int L = p.L;
    const int nrb = (L + 15) >> 4, LR = nrb * 16, NR = p.crb * 16;

    T* Kx = reinterpret_cast<T*>(smem_raw);                      // Kt (bf16) / Ks (f32)
    constexpr int KX_ELEMS = F32 ? G::KP * HD : HD * G::STRIDE;
    T* Vs = Kx + KX_ELEMS;
    T* Qs = Vs + G::KP * HD;
    T* dOs = Qs + LR * HD;
    TP* sS = reinterpret_cast<TP*>(dOs + LR * HD);
    TG* sG = reinterpret_cast<TG*>(sS + NR * G::STRIDE);
    T* dSt = reinterpret_cast<T*>(sG + NR * G::STRIDE);
    T* Apt = dSt + NR * G::STRIDE;
    float* dKacc = reinterpret_cast<float*>(Apt + NR * G::STRIDE);   // f32 path only
    float* dVacc = dKacc + G::KP * HD;

    const T* qg = static_cast<const T*>(p.q) + (size_t)b * L * p.ldqkv + h * HD;
    const T* kg = static_cast<const T*>(p.k) + (size_t)b * L * p.ldqkv + h * HD;
    const T* vg = static_cast<const T*>(p.v) + (size_t)b * L * p.ldqkv + h * HD;
    const T* og = static_cast<const T*>(p.o) + (size_t)b * L * p.lddo + h * HD;
    const T* dog = static_cast<const T*>(p.d_o) + (size_t)b * L * p.lddo + h * HD;
    T* dqg = static_cast<T*>(p.dq) + (size_t)b * L * p.lddqkv + h * HD;
    T* dkg = static_cast<T*>(p.dk) + (size_t)b * L * p.lddqkv + h * HD;
    T* dvg = static_cast<T*>(p.dv) + (size_t)b * L * p.lddqkv + h * HD;
    const TP* sg = static_cast<const TP*>(p.s) + (size_t)bh * L * L;
    const TG* dpo = p.dpout ? static_cast<const TG*>(p.dpout) + (size_t)bh * L * L : nullptr;
    TG* dpi = static_cast<TG*>(p.dpin) + (size_t)bh * L * L;

    // ---- stage K, V, Q, dO of this head (zero rows beyond L)
    for (int r = tid; r < max(G::KP, LR); r += nthr) {
        float kr[HD], vr[HD], qr[HD], gr[HD];
        if (r < L) {
            load_head_row(kr, kg + (size_t)r * p.ldqkv);
            load_head_row(vr, vg + (size_t)r * p.ldqkv);
            load_head_row(qr, qg + (size_t)r * p.ldqkv);
            load_head_row(gr, dog + (size_t)r * p.lddo);
        } else {
#pragma unroll
            for (int d = 0; d < HD; ++d) kr[d] = vr[d] = qr[d] = gr[d] = 0.f;
        }
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            if (r < G::KP) {
                if constexpr (F32) Kx[r * HD + d] = kr[d];
                else Kx[d * G::STRIDE + r] = from_f<T>(kr[d]);
                Vs[r * HD + d] = from_f<T>(vr[d]);
            }
            if (r < LR) {
                Qs[r * HD + d] = from_f<T>(qr[d]);
                dOs[r * HD + d] = from_f<T>(gr[d]);
            }
        }
    }
    if constexpr (F32) {
        for (int i = tid; i < G::KP * HD; i += nthr) dKacc[i] = dVacc[i] = 0.f;
    }

    const uint32_t rkey = rng_stream_key(p.seed, (uint32_t)bh);
    const bool do_drop = p.thresh16 != 0;

    float dk_acc[MAXKB16][4], dv_acc[MAXKB16][4];
#pragma unroll
    for (int i = 0; i < MAXKB16; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) dk_acc[i][c] = dv_acc[i][c] = 0.f;

    for (int rb0 = 0; rb0 < nrb; rb0 += p.crb) {
        const int row0 = rb0 * 16;
        const int nrows = min(L - row0, NR);
        const int nrb_chunk = min(nrb - rb0, p.crb);
        __syncthreads();
        slab_load<TP, G::STRIDE>(sS, sg + (size_t)row0 * L, nrows, L, tid, nthr);
        if (dpo) slab_load<TG, G::STRIDE>(sG, dpo + (size_t)row0 * L, nrows, L, tid, nthr);
        __syncthreads();

        // ================= phase 1: per 16-row block: A, dA, dS, dQ
        const int rb = rb0 + warp;
        if (warp < p.crb && rb < nrb) {
            const int ra = rb * 16 + g, rbb = ra + 8;
            const int la = ra - row0, lb = la + 8;
            const TP* srow_a = sS + la * G::STRIDE;
            const TP* srow_b = sS + lb * G::STRIDE;
            TG* grow_a = sG + la * G::STRIDE;
            TG* grow_b = sG + lb * G::STRIDE;
            T* ap_a = Apt + la * G::STRIDE;
            T* ap_b = Apt + lb * G::STRIDE;
            T* ds_a = dSt + la * G::STRIDE;
            T* ds_b = dSt + lb * G::STRIDE;
            const bool va_ok = ra < L, vb_ok = rbb < L;

            // ---- softmax recompute: s[][] := A (normalised probabilities)
            float s[NKB][4];
            float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const int col = kb * 8 + 2 * q4;
                float2 xa = slab_get2<TP>(srow_a, col), xb = slab_get2<TP>(srow_b, col);
                if (col >= L || !va_ok) xa.x = -INFINITY;
                if (col + 1 >= L || !va_ok) xa.y = -INFINITY;
                if (col >= L || !vb_ok) xb.x = -INFINITY;
                if (col + 1 >= L || !vb_ok) xb.y = -INFINITY;
                s[kb][0] = xa.x; s[kb][1] = xa.y; s[kb][2] = xb.x; s[kb][3] = xb.y;
                ma = fmaxf(ma, fmaxf(xa.x, xa.y));
                mb = fmaxf(mb, fmaxf(xb.x, xb.y));
            }
            ma = quad_max(ma);
            mb = quad_max(mb);
            if (ma == -INFINITY) ma = 0.f;      // rows beyond L: keep everything finite (A = 0)
            if (mb == -INFINITY) mb = 0.f;
            float suma = 0.f, sumb = 0.f;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                if constexpr (F32) {
                    s[kb][0] = expf(s[kb][0] - ma); s[kb][1] = expf(s[kb][1] - ma);
                    s[kb][2] = expf(s[kb][2] - mb); s[kb][3] = expf(s[kb][3] - mb);
                } else {
                    s[kb][0] = __expf(s[kb][0] - ma); s[kb][1] = __expf(s[kb][1] - ma);
                    s[kb][2] = __expf(s[kb][2] - mb); s[kb][3] = __expf(s[kb][3] - mb);
                }
                suma += s[kb][0] + s[kb][1];
                sumb += s[kb][2] + s[kb][3];
            }
            suma = quad_sum(suma);
            sumb = quad_sum(sumb);
            const float inva = suma > 0.f ? 1.f / suma : 0.f, invb = sumb > 0.f ? 1.f / sumb : 0.f;

            // ---- delta = dO . O per row; dO fragments
            float dela = 0.f, delb = 0.f;
            uint32_t ga0 = 0u, ga1 = 0u;          // bf16 A-operand fragments of dO rows ra / rbb
            float gfa[HD], gfb[HD];               // f32 path
            if constexpr (!F32) {
                const uint32_t* dO32 = reinterpret_cast<const uint32_t*>(dOs);
                ga0 = dO32[ra * 4 + q4];
                ga1 = dO32[rbb * 4 + q4];
                if (va_ok) {
                    const float2 x = unpack_bf16(ga0);
                    const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(og + (size_t)ra * p.lddo + 2 * q4));
                    dela = x.x * y.x + x.y * y.y;
                }
                if (vb_ok) {
                    const float2 x = unpack_bf16(ga1);
                    const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(og + (size_t)rbb * p.lddo + 2 * q4));
                    delb = x.x * y.x + x.y * y.y;
                }
            } else {
#pragma unroll
                for (int d = 0; d < HD; ++d) { gfa[d] = dOs[ra * HD + d]; gfb[d] = dOs[rbb * HD + d]; }
                if (va_ok) {
                    dela = gfa[2 * q4] * og[(size_t)ra * p.lddo + 2 * q4] + gfa[2 * q4 + 1] * og[(size_t)ra * p.lddo + 2 * q4 + 1];
                }
                if (vb_ok) {
                    delb = gfb[2 * q4] * og[(size_t)rbb * p.lddo + 2 * q4] + gfb[2 * q4 + 1] * og[(size_t)rbb * p.lddo + 2 * q4 + 1];
                }
            }
            dela = quad_sum(dela);
            delb = quad_sum(delb);

            // ---- per key block: dA' = dO V^T, A', dS; dS kept in s[][] for dQ
            const uint32_t* Vs32 = reinterpret_cast<const uint32_t*>(Vs);
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
                    const uint32_t ba = rng_pair_bits(rkey, ra, col), bb = rng_pair_bits(rkey, rbb, col);
                    k0 = rng_keep(ba, 0, p.thresh16) ? p.keep_scale : 0.f;
                    k1 = rng_keep(ba, 1, p.thresh16) ? p.keep_scale : 0.f;
                    k2 = rng_keep(bb, 0, p.thresh16) ? p.keep_scale : 0.f;
                    k3 = rng_keep(bb, 1, p.thresh16) ? p.keep_scale : 0.f;
                }
                const float A0 = s[kb][0] * inva, A1 = s[kb][1] * inva, A2 = s[kb][2] * invb, A3 = s[kb][3] * invb;
                // A' (dropped probabilities) for dV
                if constexpr (F32) {
                    *reinterpret_cast<float2*>(ap_a + col) = make_float2(A0 * k0, A1 * k1);
                    *reinterpret_cast<float2*>(ap_b + col) = make_float2(A2 * k2, A3 * k3);
                } else {
                    *reinterpret_cast<uint32_t*>(ap_a + col) = pack_bf16(A0 * k0, A1 * k1);
                    *reinterpret_cast<uint32_t*>(ap_b + col) = pack_bf16(A2 * k2, A3 * k3);
                }
                // dS = A o (dA - delta) + dP'
                float2 ua = make_float2(0.f, 0.f), ub = make_float2(0.f, 0.f);
                if (dpo) { ua = slab_get2<TG>(grow_a, col); ub = slab_get2<TG>(grow_b, col); }
                float d0 = A0 * (da[0] * k0 - dela) + ua.x;
                float d1 = A1 * (da[1] * k1 - dela) + ua.y;
                float d2 = A2 * (da[2] * k2 - delb) + ub.x;
                float d3 = A3 * (da[3] * k3 - delb) + ub.y;
                // masked (S = -inf), padded columns and padded rows carry no gradient
                const float2 xa = slab_get2<TP>(srow_a, col), xb = slab_get2<TP>(srow_b, col);
                if (col >= L || !va_ok || xa.x == -INFINITY) d0 = 0.f;
                if (col + 1 >= L || !va_ok || xa.y == -INFINITY) d1 = 0.f;
                if (col >= L || !vb_ok || xb.x == -INFINITY) d2 = 0.f;
                if (col + 1 >= L || !vb_ok || xb.y == -INFINITY) d3 = 0.f;
                // stored (rounded) values are what the previous layer sees; use them for dQ/dK too
                const float2 sa = slab_put2<TG>(grow_a, col, d0, d1);
                const float2 sb = slab_put2<TG>(grow_b, col, d2, d3);
                s[kb][0] = sa.x; s[kb][1] = sa.y; s[kb][2] = sb.x; s[kb][3] = sb.y;
                if constexpr (F32) {
                    *reinterpret_cast<float2*>(ds_a + col) = sa;
                    *reinterpret_cast<float2*>(ds_b + col) = sb;
                } else {
                    *reinterpret_cast<uint32_t*>(ds_a + col) = pack_bf16(sa.x, sa.y);
                    *reinterpret_cast<uint32_t*>(ds_b + col) = pack_bf16(sb.x, sb.y);
                }
            }

            // ---- dQ = scale * dS K
            if constexpr (!F32) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t* Kt32 = reinterpret_cast<const uint32_t*>(Kx);
#pragma unroll
                for (int j = 0; j < G::NKB16; ++j) {
                    const int kb0 = 2 * j, kb1 = 2 * j + 1;
                    const uint32_t a0 = pack_bf16(s[kb0][0], s[kb0][1]);
                    const uint32_t a1 = pack_bf16(s[kb0][2], s[kb0][3]);
                    uint32_t a2 = 0u, a3 = 0u, b1 = 0u;
                    const uint32_t b0 = Kt32[(g * G::STRIDE + kb0 * 8 + 2 * q4) >> 1];
                    if (kb1 < NKB) {
                        a2 = pack_bf16(s[kb1][0], s[kb1][1]);
                        a3 = pack_bf16(s[kb1][2], s[kb1][3]);
                        b1 = Kt32[(g * G::STRIDE + kb1 * 8 + 2 * q4) >> 1];
                    }
                    mma_bf16_16816(o, a0, a1, a2, a3, b0, b1);
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
                        const float* kr = reinterpret_cast<const float*>(Kx) + (kb * 8 + 2 * q4 + e) * HD;
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            oa[d] = fmaf(s[kb][e], kr[d], oa[d]);
                            ob[d] = fmaf(s[kb][2 + e], kr[d], ob[d]);
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
        }
        __syncthreads();

        // ================= phase 2: dK += dS^T Q, dV += A'^T dO over this chunk's rows
        slab_store<TG, G::STRIDE>(dpi + (size_t)row0 * L, sG, nrows, L, tid, nthr);
        if constexpr (!F32) {
#pragma unroll
            for (int i = 0; i < MAXKB16; ++i) {
                const int kblk = warp + i * nwarps;
                if (kblk < nrb && kblk * 16 < G::KP) {
                    const int key0 = kblk * 16;
                    // the second 8-key half lies beyond the padded key range when NKB is odd
                    const bool hi_ok = key0 + 8 < G::KP;
                    // ldmatrix.x4.trans: lane -> row address of matrix m = lane/8
                    //   m0: rows +0..7, keys key0..+7      m1: rows +0..7,  keys key0+8..
                    //   m2: rows +8..15, keys key0..+7     m3: rows +8..15, keys key0+8..
                    const int m = lane >> 3, rr = lane & 7;
                    const int soff = (rr + ((m & 2) ? 8 : 0)) * G::STRIDE + key0 + (((m & 1) && hi_ok) ? 8 : 0);
                    for (int r = 0; r < nrb_chunk; ++r) {
                        const int lr0 = r * 16;                 // local row of the chunk
                        const int brow = row0 + lr0 + (lane & 15);   // Q / dO row for ldmatrix.x2
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
                    xk = fmaf(reinterpret_cast<const float*>(dSt)[r * G::STRIDE + key], reinterpret_cast<const float*>(Qs)[(row0 + r) * HD + d], xk);
                    xv = fmaf(reinterpret_cast<const float*>(Apt)[r * G::STRIDE + key], reinterpret_cast<const float*>(dOs)[(row0 + r) * HD + d], xv);
                }
                dKacc[idx] += xk;
                dVacc[idx] += xv;
            }
        }
    }

    // ---- write dK, dV
    if constexpr (!F32) {
#pragma unroll
        for (int i = 0; i < MAXKB16; ++i) {
            const int kblk = warp + i * nwarps;
            if (kblk < nrb && kblk * 16 < G::KP) {
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
        __syncthreads();
        for (int idx = tid; idx < L * HD; idx += nthr) {
            const int key = idx >> 3, d = idx & 7;
            dkg[(size_t)key * p.lddqkv + d] = dKacc[idx] * p.scale;
            dvg[(size_t)key * p.lddqkv + d] = dVacc[idx];
        }
    }
}

// ------------------------------------------------------------------ debug: keep mask
__global__ void dropout_mask_kernel(uint8_t* keep, int H, int L, uint32_t thresh16, unsigned long long seed) {
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

template <typename T, int NKB> size_t fwd_smem(int crb, size_t sz_tp) {
    using G = Geo<NKB>;
    constexpr bool F32 = std::is_same<T, float>::value;
    size_t kv = (size_t)G::KP * HD * sizeof(T) + (F32 ? (size_t)G::KP * HD : (size_t)HD * G::STRIDE) * sizeof(T);
    return kv + (size_t)crb * 16 * G::STRIDE * sz_tp;
}
template <typename T, int NKB> size_t bwd_smem(int crb, int L, size_t sz_tp, size_t sz_tg) {
    using G = Geo<NKB>;
    constexpr bool F32 = std::is_same<T, float>::value;
    const int LR = ((L + 15) / 16) * 16, NR = crb * 16;
    size_t s = (F32 ? (size_t)G::KP * HD : (size_t)HD * G::STRIDE) * sizeof(T);
    s += (size_t)G::KP * HD * sizeof(T) + 2 * (size_t)LR * HD * sizeof(T);
    s += (size_t)NR * G::STRIDE * (sz_tp + sz_tg + 2 * sizeof(T));
    if (F32) s += 2 * (size_t)G::KP * HD * sizeof(float);
    return s;
}

constexpr size_t SMEM_CAP = 200 * 1024;

// warps per CTA: one per 16-row block, at most 8, balanced over the chunks
inline int pick_warps(int nrb) {
    const int chunks = (nrb + 7) / 8;
    return (nrb + chunks - 1) / chunks;
}

template <typename T, typename TP, int NKB>
int launch_fwd(FwdParams p, cudaStream_t st) {
    const int nrb = (p.L + 15) / 16;
    const int nw = pick_warps(nrb);
    int crb = nw;
    while (crb > 1 && fwd_smem<T, NKB>(crb, sizeof(TP)) > SMEM_CAP) --crb;
    const size_t smem = fwd_smem<T, NKB>(crb, sizeof(TP));
    MMDTI_REQUIRE(smem <= SMEM_CAP, "pair_attn_fwd: shared memory %zu exceeds cap", smem);
    p.crb = crb;
    auto kern = pair_attn_fwd_kernel<T, TP, NKB>;
    MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<p.B * p.H, nw * 32, smem, st>>>(p);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
template <typename T, typename TP, typename TG, int NKB>
int launch_bwd(BwdParams p, cudaStream_t st) {
    const int nrb = (p.L + 15) / 16;
    const int nw = pick_warps(nrb);      // ceil(nrb/nw) <= 3 since nw >= nrb/ceil(nrb/8) and nrb <= 17
    MMDTI_REQUIRE((nrb + nw - 1) / nw <= 3, "pair_attn_bwd: L=%d too large", p.L);
    int crb = nw;
    while (crb > 1 && bwd_smem<T, NKB>(crb, p.L, sizeof(TP), sizeof(TG)) > SMEM_CAP) --crb;
    const size_t smem = bwd_smem<T, NKB>(crb, p.L, sizeof(TP), sizeof(TG));
    MMDTI_REQUIRE(smem <= 227 * 1024, "pair_attn_bwd: shared memory %zu exceeds 227 KB (L=%d)", smem, p.L);
    p.crb = crb;
    auto kern = pair_attn_bwd_kernel<T, TP, TG, NKB>;
    MMDTI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<p.B * p.H, nw * 32, smem, st>>>(p);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

template <typename T, typename TP>
int dispatch_fwd_nkb(const FwdParams& p, cudaStream_t st) {
    const int nkb = (p.L + 7) / 8;
    if (nkb <= 4) return launch_fwd<T, TP, 4>(p, st);
    if (nkb <= 9) return launch_fwd<T, TP, 9>(p, st);
    if (nkb <= 17) return launch_fwd<T, TP, 17>(p, st);
    if (nkb <= 33) return launch_fwd<T, TP, 33>(p, st);
    mmdti_set_error("pair_attn_fwd: L=%d exceeds the supported maximum 264", p.L);
    return MMDTI_ERR_ARG;
}
template <typename T, typename TP, typename TG>
int dispatch_bwd_nkb(const BwdParams& p, cudaStream_t st) {
    const int nkb = (p.L + 7) / 8;
    if (nkb <= 4) return launch_bwd<T, TP, TG, 4>(p, st);
    if (nkb <= 9) return launch_bwd<T, TP, TG, 9>(p, st);
    if (nkb <= 17) return launch_bwd<T, TP, TG, 17>(p, st);
    if (nkb <= 33) return launch_bwd<T, TP, TG, 33>(p, st);
    mmdti_set_error("pair_attn_bwd: L=%d exceeds the supported maximum 264", p.L);
    return MMDTI_ERR_ARG;
}

int check_common(const void* q, const void* k, const void* v, long long ld, int B, int H, int L, int act_dtype) {
    MMDTI_REQUIRE(B > 0 && H > 0 && L > 0, "pair_attn: empty problem B=%d H=%d L=%d", B, H, L);
    MMDTI_REQUIRE(act_dtype == MMDTI_F32 || act_dtype == MMDTI_BF16, "pair_attn: act_dtype must be f32 or bf16");
    const size_t esz = act_dtype == MMDTI_F32 ? 4 : 2;
    MMDTI_REQUIRE(mmdti_aligned(q, 16) && mmdti_aligned(k, 16) && mmdti_aligned(v, 16) && (ld * esz) % 16 == 0,
                  "pair_attn: q/k/v must be 16-byte aligned with a 16-byte-multiple row stride");
    return MMDTI_OK;
}

}  // namespace

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
    p.ldqkv = ldqkv; p.ldo = ldo; p.B = B; p.H = H; p.L = L; p.scale = scale; p.seed = seed;
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
                                   const void* o, const void* d_o, int64_t lddo, const void* d_pair_out, void* d_pair_in, void* dq,
                                   void* dk, void* dv, int64_t lddqkv, int B, int H, int L, float scale,
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
    p.B = B; p.H = H; p.L = L; p.scale = scale; p.seed = seed;
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
    dropout_mask_kernel<<<B * H, 256, 0, static_cast<cudaStream_t>(stream)>>>(keep, H, L, thresh16, seed);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
