// Cross-modal attention of the fusion block: BertCoAttention.forward, models/mm_module.py:493-522,
// called twice per step by CrossAttentionModel.forward (models/mm_model.py:386-406): graph tokens attend to SMILES tokens
// and SMILES tokens attend to graph tokens, 16 heads x 32, additive key mask (1 - mask) * -10000, softmax, dropout on the
// probabilities, context = P V.  Also the masked mean pooling that follows it (models/mm_model.py:571-576).
//
// bf16 path: flash-style kernels on mma.sync m16n8k16 (the problem is tiny -- 1 GFLOP per direction at config 2 -- and
// launch/latency bound; the operands of one (molecule, head) fit in shared memory, nothing is re-read from HBM):
//   forward      one CTA per (b, h, 64 query rows), 4 warps x 16 rows, keys streamed in blocks of 64, online softmax in
//                the log2 domain, row log-sum-exp saved;
//   backward dQ  same decomposition: recompute P from the saved LSE, dP = dO V^T, dS = P (dP - delta), dQ = scale dS K;
//                also produces delta = rowsum(dO * O);
//   backward dKV one CTA per (b, h, 64 keys), warps own 16 keys, queries streamed: S^T = K Q^T, dV = P_drop^T dO,
//                dK = scale dS^T Q.  No atomics, no cross-warp reductions: results are deterministic.
// fp32 path (validation mode): one warp per row, plain FMA loops, same dropout stream.
// Dropout: counter-based hash of (seed, b*H+h, query, key) -- rng_quad_bits of common.cuh with a per-(b,h) key, so forward,
// both backward kernels and mmdti_cross_attn_dropout_mask agree under any tiling.
#include "common.cuh"

#include <math.h>

namespace {
constexpr float CA_LOG2E = 1.4426950408889634f;
constexpr float CA_MASKED = -10000.0f;          // models/mm_model.py:394,400: (1.0 - mask) * -10000.0

inline void ca_drop_params(float p, uint32_t& thresh16, float& keep_scale) {
    double t = floor((double)p * 65536.0 + 0.5);
    if (t < 0) t = 0;
    if (t > 65535) t = 65535;
    thresh16 = (uint32_t)t;
    keep_scale = (float)(65536.0 / (65536.0 - t));
}
inline uint32_t ca_seed_key(uint64_t seed) { return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x2545F491U)); }
__device__ __forceinline__ uint32_t ca_bh_key(uint32_t key, const unsigned long long* off, uint32_t bh) {
    return mix32(rng_effective_key(key, off) ^ (bh * 0x9E3779B1U + 0x7F4A7C15U));
}
__device__ __forceinline__ bool ca_keep(uint32_t key, uint32_t qrow, uint32_t kcol, uint32_t thresh16) {
    return rng_keep(rng_pair_bits(key, qrow, kcol), kcol, thresh16);
}

// 64 rows x HD bf16 from a (rows, ld) matrix into padded shared rows; rows >= nvalid are zero-filled
template <int HD> __device__ __forceinline__ void ca_load_rows(bf16* dst, const bf16* src, long long ld, int row0, int nvalid, int tid) {
    constexpr int LD = HD + 8, CPR = HD / 8;
    for (int i = tid; i < 64 * CPR; i += 128) {
        const int r = i / CPR, c = i % CPR;
        if (row0 + r < nvalid) cp_async_16(dst + r * LD + c * 8, src + (long long)(row0 + r) * ld + c * 8);
        else *reinterpret_cast<uint4*>(dst + r * LD + c * 8) = make_uint4(0, 0, 0, 0);
    }
}
// A fragments (16 rows x HD) of a padded shared tile
template <int HD> __device__ __forceinline__ void ca_a_frags(uint32_t (&a)[HD / 16][4], const bf16* tile, int row0, int lane) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk)
        ldmatrix_x4(a[kk][0], a[kk][1], a[kk][2], a[kk][3], tile + (row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + kk * 16 + (lane >> 4) * 8);
}
// c (16 x 8) += A (16 x HD) . Y[n0 .. n0+8)[0 .. HD)^T with Y a padded shared tile (rows = n, row-major in k)
template <int HD> __device__ __forceinline__ void ca_mma_nt(float (&c)[4], const uint32_t (&a)[HD / 16][4], const bf16* Y, int n0, int lane) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int k2 = 0; k2 < HD / 32; ++k2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, Y + (n0 + (lane & 7)) * LD + k2 * 32 + (lane >> 3) * 8);
        mma_bf16_16816(c, a[2 * k2][0], a[2 * k2][1], a[2 * k2][2], a[2 * k2][3], b0, b1);
        mma_bf16_16816(c, a[2 * k2 + 1][0], a[2 * k2 + 1][1], a[2 * k2 + 1][2], a[2 * k2 + 1][3], b2, b3);
    }
}
// acc (16 x HD) += A (16 x 16, one k-step) . Z[k0 .. k0+16)[0 .. HD) with Z a padded shared tile (rows = k)
template <int HD> __device__ __forceinline__ void ca_mma_nn(float (&acc)[HD / 8][4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                            const bf16* Z, int k0, int lane) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int n2 = 0; n2 < HD / 16; ++n2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3, Z + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + n2 * 16 + (lane >> 4) * 8);
        mma_bf16_16816(acc[2 * n2], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(acc[2 * n2 + 1], a0, a1, a2, a3, b2, b3);
    }
}

struct CAParams {
    const bf16 *q, *k, *v;
    long long ldq, ldkv;
    const uint8_t* kmask;        // (B, Lk) 1 = attend
    int H, Lq, Lk;
    float scale, scale_log2, keep_scale;
    uint32_t key, thresh16;
    const unsigned long long* seed_off;
};

// ------------------------------------------------------------------------------------------------ forward (bf16)
template <int HD>
__global__ void __launch_bounds__(128) cross_attn_fwd_kernel(CAParams P, bf16* __restrict__ o, long long ldo, float* __restrict__ lse2) {
    constexpr int LD = HD + 8;
    __shared__ __align__(16) bf16 Qs[64 * LD], Ks[64 * LD], Vs[64 * LD];
    __shared__ float madd[64];
    const int bh = blockIdx.x, b = bh / P.H, h = bh % P.H, q0 = blockIdx.y * 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const uint32_t key = ca_bh_key(P.key, P.seed_off, (uint32_t)bh);
    const uint32_t th2 = P.thresh16 * 0x10001U;
    const bf16* qb = P.q + (long long)b * P.Lq * P.ldq + h * HD;
    const bf16* kb = P.k + (long long)b * P.Lk * P.ldkv + h * HD;
    const bf16* vb = P.v + (long long)b * P.Lk * P.ldkv + h * HD;
    ca_load_rows<HD>(Qs, qb, P.ldq, q0, P.Lq, tid);
    uint32_t qa[HD / 16][4];
    float oacc[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int row_g = q0 + warp * 16 + g;
    const bool active = q0 + warp * 16 < P.Lq;
    for (int k0 = 0; k0 < P.Lk; k0 += 64) {
        if (k0) __syncthreads();
        ca_load_rows<HD>(Ks, kb, P.ldkv, k0, P.Lk, tid);
        ca_load_rows<HD>(Vs, vb, P.ldkv, k0, P.Lk, tid);
        if (tid < 64) {
            const int kc = k0 + tid;
            madd[tid] = kc < P.Lk ? (P.kmask[(long long)b * P.Lk + kc] ? 0.f : CA_MASKED * CA_LOG2E) : -INFINITY;
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (!active) continue;                                 // no valid query row in this warp: it only helps loading
        if (k0 == 0) ca_a_frags<HD>(qa, Qs, warp * 16, lane);
        const int nt = min(8, (P.Lk - k0 + 7) >> 3);             // 8-key tiles of this block that hold valid keys
        float s[8][4];
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            if (j < nt) {
                ca_mma_nt<HD>(s[j], qa, Ks, 8 * j, lane);
                const float a0 = madd[8 * j + 2 * q4], a1 = madd[8 * j + 2 * q4 + 1];
                s[j][0] = fmaf(s[j][0], P.scale_log2, a0);
                s[j][1] = fmaf(s[j][1], P.scale_log2, a1);
                s[j][2] = fmaf(s[j][2], P.scale_log2, a0);
                s[j][3] = fmaf(s[j][3], P.scale_log2, a1);
                mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
                mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
            }
        }
        mx0 = quad_max(mx0);
        mx1 = quad_max(mx1);
        const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
        const float al0 = fast_ex2(m_run[0] - mn0), al1 = fast_ex2(m_run[1] - mn1);
        m_run[0] = mn0;
        m_run[1] = mn1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nt) {
                s[j][0] = fast_ex2(s[j][0] - mn0);
                s[j][1] = fast_ex2(s[j][1] - mn0);
                s[j][2] = fast_ex2(s[j][2] - mn1);
                s[j][3] = fast_ex2(s[j][3] - mn1);
                ps0 += s[j][0] + s[j][1];
                ps1 += s[j][2] + s[j][3];
            }
        }
        l_run[0] = l_run[0] * al0 + ps0;
        l_run[1] = l_run[1] * al1 + ps1;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            oacc[n][0] *= al0;
            oacc[n][1] *= al0;
            oacc[n][2] *= al1;
            oacc[n][3] *= al1;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (2 * t < nt) {                                  // tile 2t+1 beyond nt holds zeros
                const uint2 w0 = rng_quad_bits(key, (uint32_t)row_g, (uint32_t)(k0 + 16 * t + 2 * q4));
                const uint2 w1 = rng_quad_bits(key, (uint32_t)(row_g + 8), (uint32_t)(k0 + 16 * t + 2 * q4));
                const uint32_t a0 = pack_bf16(s[2 * t][0], s[2 * t][1]) & rng_keep_mask2(w0.x, th2);
                const uint32_t a1 = pack_bf16(s[2 * t][2], s[2 * t][3]) & rng_keep_mask2(w1.x, th2);
                const uint32_t a2 = pack_bf16(s[2 * t + 1][0], s[2 * t + 1][1]) & rng_keep_mask2(w0.y, th2);
                const uint32_t a3 = pack_bf16(s[2 * t + 1][2], s[2 * t + 1][3]) & rng_keep_mask2(w1.y, th2);
                ca_mma_nn<HD>(oacc, a0, a1, a2, a3, Vs, 16 * t, lane);
            }
        }
    }
    if (!active) return;
    const float l0 = quad_sum(l_run[0]), l1 = quad_sum(l_run[1]);
    const float i0 = P.keep_scale / l0, i1 = P.keep_scale / l1;
    if (row_g < P.Lq) {
        bf16* orow = o + ((long long)b * P.Lq + row_g) * ldo + h * HD + 2 * q4;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(orow + 8 * n) = pack_bf16(oacc[n][0] * i0, oacc[n][1] * i0);
        if (q4 == 0) lse2[(long long)bh * P.Lq + row_g] = m_run[0] + log2f(l0);
    }
    if (row_g + 8 < P.Lq) {
        bf16* orow = o + ((long long)b * P.Lq + row_g + 8) * ldo + h * HD + 2 * q4;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(orow + 8 * n) = pack_bf16(oacc[n][2] * i1, oacc[n][3] * i1);
        if (q4 == 0) lse2[(long long)bh * P.Lq + row_g + 8] = m_run[1] + log2f(l1);
    }
}

// ------------------------------------------------------------------------------------------------ backward dQ (bf16)
template <int HD>
__global__ void __launch_bounds__(128) cross_attn_bwd_q_kernel(CAParams P, const bf16* __restrict__ o, const bf16* __restrict__ d_o, long long ldo,
                                                               const float* __restrict__ lse2, float* __restrict__ delta,
                                                               bf16* __restrict__ dq, long long lddq) {
    constexpr int LD = HD + 8;
    __shared__ __align__(16) bf16 Qs[64 * LD], Gs[64 * LD], Ks[64 * LD], Vs[64 * LD];
    __shared__ float madd[64];
    const int bh = blockIdx.x, b = bh / P.H, h = bh % P.H, q0 = blockIdx.y * 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const uint32_t key = ca_bh_key(P.key, P.seed_off, (uint32_t)bh);
    const bf16* qb = P.q + (long long)b * P.Lq * P.ldq + h * HD;
    const bf16* gb = d_o + (long long)b * P.Lq * ldo + h * HD;
    const bf16* ob = o + (long long)b * P.Lq * ldo + h * HD;
    const bf16* kb = P.k + (long long)b * P.Lk * P.ldkv + h * HD;
    const bf16* vb = P.v + (long long)b * P.Lk * P.ldkv + h * HD;
    ca_load_rows<HD>(Qs, qb, P.ldq, q0, P.Lq, tid);
    ca_load_rows<HD>(Gs, gb, ldo, q0, P.Lq, tid);
    // delta = rowsum(dO * O): lane -> (row lane % 16, half lane / 16 of the head dims)
    float dl;
    {
        const int r = q0 + warp * 16 + (lane & 15), half = lane >> 4;
        float acc = 0.f;
        if (r < P.Lq) {
            const bf16* orow = ob + (long long)r * ldo + half * (HD / 2);
            const bf16* grow = gb + (long long)r * ldo + half * (HD / 2);
#pragma unroll
            for (int c = 0; c < HD / 16; ++c) {
                const uint4 ov = *reinterpret_cast<const uint4*>(orow + 8 * c), gv = *reinterpret_cast<const uint4*>(grow + 8 * c);
                const uint32_t ou[4] = {ov.x, ov.y, ov.z, ov.w}, gu[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 a = unpack_bf16(ou[e]), c2 = unpack_bf16(gu[e]);
                    acc = fmaf(a.x, c2.x, fmaf(a.y, c2.y, acc));
                }
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < 16 && r < P.Lq) delta[(long long)bh * P.Lq + r] = acc;
        dl = acc;
    }
    const float dl0 = __shfl_sync(0xffffffffu, dl, g), dl1 = __shfl_sync(0xffffffffu, dl, g + 8);
    const int row_g = q0 + warp * 16 + g;
    const bool active = q0 + warp * 16 < P.Lq;
    const float ls0 = row_g < P.Lq ? lse2[(long long)bh * P.Lq + row_g] : INFINITY;
    const float ls1 = row_g + 8 < P.Lq ? lse2[(long long)bh * P.Lq + row_g + 8] : INFINITY;
    uint32_t qa[HD / 16][4], ga[HD / 16][4];
    float acc[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    for (int k0 = 0; k0 < P.Lk; k0 += 64) {
        if (k0) __syncthreads();
        ca_load_rows<HD>(Ks, kb, P.ldkv, k0, P.Lk, tid);
        ca_load_rows<HD>(Vs, vb, P.ldkv, k0, P.Lk, tid);
        if (tid < 64) {
            const int kc = k0 + tid;
            madd[tid] = kc < P.Lk ? (P.kmask[(long long)b * P.Lk + kc] ? 0.f : CA_MASKED * CA_LOG2E) : -INFINITY;
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (!active) continue;
        if (k0 == 0) {
            ca_a_frags<HD>(qa, Qs, warp * 16, lane);
            ca_a_frags<HD>(ga, Gs, warp * 16, lane);
        }
        const int nt = min(8, (P.Lk - k0 + 7) >> 3);
        float ds[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ds[j][0] = ds[j][1] = ds[j][2] = ds[j][3] = 0.f;
            if (j >= nt) continue;
            float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
            ca_mma_nt<HD>(s, qa, Ks, 8 * j, lane);
            ca_mma_nt<HD>(dp, ga, Vs, 8 * j, lane);
            const float a0 = madd[8 * j + 2 * q4], a1 = madd[8 * j + 2 * q4 + 1];
            const float p0 = fast_ex2(fmaf(s[0], P.scale_log2, a0) - ls0), p1 = fast_ex2(fmaf(s[1], P.scale_log2, a1) - ls0);
            const float p2 = fast_ex2(fmaf(s[2], P.scale_log2, a0) - ls1), p3 = fast_ex2(fmaf(s[3], P.scale_log2, a1) - ls1);
            const uint32_t col = (uint32_t)(k0 + 8 * j + 2 * q4);
            const uint32_t w0 = rng_pair_bits(key, (uint32_t)row_g, col), w1 = rng_pair_bits(key, (uint32_t)(row_g + 8), col);
            const float ks = P.keep_scale;
            ds[j][0] = p0 * (((w0 & 0xffffu) >= P.thresh16 ? dp[0] * ks : 0.f) - dl0);
            ds[j][1] = p1 * (((w0 >> 16) >= P.thresh16 ? dp[1] * ks : 0.f) - dl0);
            ds[j][2] = p2 * (((w1 & 0xffffu) >= P.thresh16 ? dp[2] * ks : 0.f) - dl1);
            ds[j][3] = p3 * (((w1 >> 16) >= P.thresh16 ? dp[3] * ks : 0.f) - dl1);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (2 * t < nt)
                ca_mma_nn<HD>(acc, pack_bf16(ds[2 * t][0], ds[2 * t][1]), pack_bf16(ds[2 * t][2], ds[2 * t][3]),
                              pack_bf16(ds[2 * t + 1][0], ds[2 * t + 1][1]), pack_bf16(ds[2 * t + 1][2], ds[2 * t + 1][3]), Ks, 16 * t, lane);
    }
    if (row_g < P.Lq) {
        bf16* r = dq + ((long long)b * P.Lq + row_g) * lddq + h * HD + 2 * q4;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(r + 8 * n) = pack_bf16(acc[n][0] * P.scale, acc[n][1] * P.scale);
    }
    if (row_g + 8 < P.Lq) {
        bf16* r = dq + ((long long)b * P.Lq + row_g + 8) * lddq + h * HD + 2 * q4;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(r + 8 * n) = pack_bf16(acc[n][2] * P.scale, acc[n][3] * P.scale);
    }
}

// ------------------------------------------------------------------------------------------------ backward dK, dV (bf16)
template <int HD>
__global__ void __launch_bounds__(128) cross_attn_bwd_kv_kernel(CAParams P, const bf16* __restrict__ d_o, long long ldo, const float* __restrict__ lse2,
                                                                const float* __restrict__ delta, bf16* __restrict__ dk, bf16* __restrict__ dv,
                                                                long long lddkv) {
    constexpr int LD = HD + 8;
    __shared__ __align__(16) bf16 Qs[64 * LD], Gs[64 * LD], Ks[64 * LD], Vs[64 * LD];
    __shared__ float lse_s[64], dl_s[64];
    const int bh = blockIdx.x, b = bh / P.H, h = bh % P.H, k0 = blockIdx.y * 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const uint32_t key = ca_bh_key(P.key, P.seed_off, (uint32_t)bh);
    const bf16* qb = P.q + (long long)b * P.Lq * P.ldq + h * HD;
    const bf16* gb = d_o + (long long)b * P.Lq * ldo + h * HD;
    const bf16* kb = P.k + (long long)b * P.Lk * P.ldkv + h * HD;
    const bf16* vb = P.v + (long long)b * P.Lk * P.ldkv + h * HD;
    ca_load_rows<HD>(Ks, kb, P.ldkv, k0, P.Lk, tid);
    ca_load_rows<HD>(Vs, vb, P.ldkv, k0, P.Lk, tid);
    const int key_g = k0 + warp * 16 + g;                   // this thread's key rows: key_g and key_g + 8
    float add0 = 0.f, add1 = 0.f;
    if (key_g < P.Lk) add0 = P.kmask[(long long)b * P.Lk + key_g] ? 0.f : CA_MASKED * CA_LOG2E;
    if (key_g + 8 < P.Lk) add1 = P.kmask[(long long)b * P.Lk + key_g + 8] ? 0.f : CA_MASKED * CA_LOG2E;
    uint32_t ka[HD / 16][4], va[HD / 16][4];
    float dka[HD / 8][4], dva[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
        dka[n][0] = dka[n][1] = dka[n][2] = dka[n][3] = 0.f;
        dva[n][0] = dva[n][1] = dva[n][2] = dva[n][3] = 0.f;
    }
    const bool hi = g & 1;                                   // halfword of the random word that belongs to this key column
    const bool active = k0 + warp * 16 < P.Lk;
    for (int q0 = 0; q0 < P.Lq; q0 += 64) {
        if (q0) __syncthreads();
        ca_load_rows<HD>(Qs, qb, P.ldq, q0, P.Lq, tid);
        ca_load_rows<HD>(Gs, gb, ldo, q0, P.Lq, tid);
        if (tid < 64) {
            const int qr = q0 + tid;
            lse_s[tid] = qr < P.Lq ? lse2[(long long)bh * P.Lq + qr] : INFINITY;
            dl_s[tid] = qr < P.Lq ? delta[(long long)bh * P.Lq + qr] : 0.f;
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        if (!active) continue;
        if (q0 == 0) {
            ca_a_frags<HD>(ka, Ks, warp * 16, lane);
            ca_a_frags<HD>(va, Vs, warp * 16, lane);
        }
        const int nt = min(8, (P.Lq - q0 + 7) >> 3);             // 8-query tiles of this block that hold valid queries
        float pd[8][4], ds[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            pd[j][0] = pd[j][1] = pd[j][2] = pd[j][3] = 0.f;
            ds[j][0] = ds[j][1] = ds[j][2] = ds[j][3] = 0.f;
            if (j >= nt) continue;
            float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
            ca_mma_nt<HD>(s, ka, Qs, 8 * j, lane);             // S^T: rows = keys, cols = queries
            ca_mma_nt<HD>(dp, va, Gs, 8 * j, lane);            // dP^T
            const int c = 8 * j + 2 * q4;
            const float l0 = lse_s[c], l1 = lse_s[c + 1], d0 = dl_s[c], d1 = dl_s[c + 1];
            const float p0 = fast_ex2(fmaf(s[0], P.scale_log2, add0) - l0), p1 = fast_ex2(fmaf(s[1], P.scale_log2, add0) - l1);
            const float p2 = fast_ex2(fmaf(s[2], P.scale_log2, add1) - l0), p3 = fast_ex2(fmaf(s[3], P.scale_log2, add1) - l1);
            // element (query, key): bits of rng_quad_bits(key, query, key column); key_g % 16 < 8 -> word .x, key_g + 8 -> word .y
            const uint2 w0 = rng_quad_bits(key, (uint32_t)(q0 + c), (uint32_t)key_g), w1 = rng_quad_bits(key, (uint32_t)(q0 + c + 1), (uint32_t)key_g);
            const bool k0_ = (hi ? (w0.x >> 16) : (w0.x & 0xffffu)) >= P.thresh16, k1_ = (hi ? (w1.x >> 16) : (w1.x & 0xffffu)) >= P.thresh16;
            const bool k2_ = (hi ? (w0.y >> 16) : (w0.y & 0xffffu)) >= P.thresh16, k3_ = (hi ? (w1.y >> 16) : (w1.y & 0xffffu)) >= P.thresh16;
            const float ks = P.keep_scale;
            pd[j][0] = k0_ ? p0 : 0.f;
            pd[j][1] = k1_ ? p1 : 0.f;
            pd[j][2] = k2_ ? p2 : 0.f;
            pd[j][3] = k3_ ? p3 : 0.f;
            ds[j][0] = p0 * ((k0_ ? dp[0] * ks : 0.f) - d0);
            ds[j][1] = p1 * ((k1_ ? dp[1] * ks : 0.f) - d1);
            ds[j][2] = p2 * ((k2_ ? dp[2] * ks : 0.f) - d0);
            ds[j][3] = p3 * ((k3_ ? dp[3] * ks : 0.f) - d1);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (2 * t >= nt) continue;
            ca_mma_nn<HD>(dva, pack_bf16(pd[2 * t][0], pd[2 * t][1]), pack_bf16(pd[2 * t][2], pd[2 * t][3]),
                          pack_bf16(pd[2 * t + 1][0], pd[2 * t + 1][1]), pack_bf16(pd[2 * t + 1][2], pd[2 * t + 1][3]), Gs, 16 * t, lane);
            ca_mma_nn<HD>(dka, pack_bf16(ds[2 * t][0], ds[2 * t][1]), pack_bf16(ds[2 * t][2], ds[2 * t][3]),
                          pack_bf16(ds[2 * t + 1][0], ds[2 * t + 1][1]), pack_bf16(ds[2 * t + 1][2], ds[2 * t + 1][3]), Qs, 16 * t, lane);
        }
    }
    const float ks = P.keep_scale, sc = P.scale;
    if (key_g < P.Lk) {
        const long long off = ((long long)b * P.Lk + key_g) * lddkv + h * HD + 2 * q4;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            *reinterpret_cast<uint32_t*>(dk + off + 8 * n) = pack_bf16(dka[n][0] * sc, dka[n][1] * sc);
            *reinterpret_cast<uint32_t*>(dv + off + 8 * n) = pack_bf16(dva[n][0] * ks, dva[n][1] * ks);
        }
    }
    if (key_g + 8 < P.Lk) {
        const long long off = ((long long)b * P.Lk + key_g + 8) * lddkv + h * HD + 2 * q4;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            *reinterpret_cast<uint32_t*>(dk + off + 8 * n) = pack_bf16(dka[n][2] * sc, dka[n][3] * sc);
            *reinterpret_cast<uint32_t*>(dv + off + 8 * n) = pack_bf16(dva[n][2] * ks, dva[n][3] * ks);
        }
    }
}

// ------------------------------------------------------------------------------------------------ fp32 validation path
// One warp per (b, h, query row) / (b, h, key row); probabilities of the row are staged in shared memory.  Natural-log LSE.
struct CAParamsF {
    const float *q, *k, *v;
    long long ldq, ldkv;
    const uint8_t* kmask;
    int H, Lq, Lk, HD;
    float scale, keep_scale;
    uint32_t key, thresh16;
    const unsigned long long* seed_off;
};
constexpr int CAF_WARPS = 4;

__device__ __forceinline__ float caf_dot(const float* a, const float* b, int n) {
    float s = 0.f;
    for (int d = 0; d < n; ++d) s = fmaf(a[d], b[d], s);
    return s;
}

__global__ void __launch_bounds__(CAF_WARPS * 32) cross_attn_f32_fwd_kernel(CAParamsF P, float* __restrict__ o, long long ldo, float* __restrict__ lse,
                                                                            long long nrows) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * CAF_WARPS + warp;          // (bh, q)
    if (r >= nrows) return;
    float* pr = sm + (size_t)warp * P.Lk;
    const int bh = (int)(r / P.Lq), qi = (int)(r % P.Lq), b = bh / P.H, h = bh % P.H;
    const uint32_t key = ca_bh_key(P.key, P.seed_off, (uint32_t)bh);
    const float* qrow = P.q + ((long long)b * P.Lq + qi) * P.ldq + h * P.HD;
    const float* kb = P.k + (long long)b * P.Lk * P.ldkv + h * P.HD;
    const float* vb = P.v + (long long)b * P.Lk * P.ldkv + h * P.HD;
    float mx = -INFINITY;
    for (int kc = lane; kc < P.Lk; kc += 32) {
        const float s = caf_dot(qrow, kb + (long long)kc * P.ldkv, P.HD) * P.scale + (P.kmask[(long long)b * P.Lk + kc] ? 0.f : CA_MASKED);
        pr[kc] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int kc = lane; kc < P.Lk; kc += 32) {
        const float e = expf(pr[kc] - mx);
        sum += e;
        pr[kc] = ca_keep(key, (uint32_t)qi, (uint32_t)kc, P.thresh16) ? e : 0.f;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = P.keep_scale / sum;
    for (int d = lane; d < P.HD; d += 32) {
        float acc = 0.f;
        for (int kc = 0; kc < P.Lk; ++kc) acc = fmaf(pr[kc], vb[(long long)kc * P.ldkv + d], acc);
        o[((long long)b * P.Lq + qi) * ldo + h * P.HD + d] = acc * inv;
    }
    if (lane == 0) lse[r] = mx + logf(sum);
}

__global__ void __launch_bounds__(CAF_WARPS * 32) cross_attn_f32_bwd_q_kernel(CAParamsF P, const float* __restrict__ o, const float* __restrict__ d_o,
                                                                              long long ldo, const float* __restrict__ lse, float* __restrict__ delta,
                                                                              float* __restrict__ dq, long long lddq, long long nrows) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * CAF_WARPS + warp;
    if (r >= nrows) return;
    float* pr = sm + (size_t)warp * P.Lk;
    const int bh = (int)(r / P.Lq), qi = (int)(r % P.Lq), b = bh / P.H, h = bh % P.H;
    const uint32_t key = ca_bh_key(P.key, P.seed_off, (uint32_t)bh);
    const float* qrow = P.q + ((long long)b * P.Lq + qi) * P.ldq + h * P.HD;
    const float* grow = d_o + ((long long)b * P.Lq + qi) * ldo + h * P.HD;
    const float* orow = o + ((long long)b * P.Lq + qi) * ldo + h * P.HD;
    const float* kb = P.k + (long long)b * P.Lk * P.ldkv + h * P.HD;
    const float* vb = P.v + (long long)b * P.Lk * P.ldkv + h * P.HD;
    float dl = 0.f;
    for (int d = lane; d < P.HD; d += 32) dl = fmaf(grow[d], orow[d], dl);
    dl = warp_sum(dl);
    if (lane == 0) delta[r] = dl;
    const float ls = lse[r];
    for (int kc = lane; kc < P.Lk; kc += 32) {
        const float s = caf_dot(qrow, kb + (long long)kc * P.ldkv, P.HD) * P.scale + (P.kmask[(long long)b * P.Lk + kc] ? 0.f : CA_MASKED);
        const float p = expf(s - ls);
        const float dp = ca_keep(key, (uint32_t)qi, (uint32_t)kc, P.thresh16) ? caf_dot(grow, vb + (long long)kc * P.ldkv, P.HD) * P.keep_scale : 0.f;
        pr[kc] = p * (dp - dl);
    }
    __syncwarp();
    for (int d = lane; d < P.HD; d += 32) {
        float acc = 0.f;
        for (int kc = 0; kc < P.Lk; ++kc) acc = fmaf(pr[kc], kb[(long long)kc * P.ldkv + d], acc);
        dq[((long long)b * P.Lq + qi) * lddq + h * P.HD + d] = acc * P.scale;
    }
}

__global__ void __launch_bounds__(CAF_WARPS * 32) cross_attn_f32_bwd_kv_kernel(CAParamsF P, const float* __restrict__ d_o, long long ldo,
                                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                                               float* __restrict__ dk, float* __restrict__ dv, long long lddkv,
                                                                               long long nrows) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * CAF_WARPS + warp;          // (bh, key)
    if (r >= nrows) return;
    float* pd = sm + (size_t)warp * 2 * P.Lq;
    float* ds = pd + P.Lq;
    const int bh = (int)(r / P.Lk), kc = (int)(r % P.Lk), b = bh / P.H, h = bh % P.H;
    const uint32_t key = ca_bh_key(P.key, P.seed_off, (uint32_t)bh);
    const float* qb = P.q + (long long)b * P.Lq * P.ldq + h * P.HD;
    const float* gb = d_o + (long long)b * P.Lq * ldo + h * P.HD;
    const float* krow = P.k + ((long long)b * P.Lk + kc) * P.ldkv + h * P.HD;
    const float* vrow = P.v + ((long long)b * P.Lk + kc) * P.ldkv + h * P.HD;
    const float add = P.kmask[(long long)b * P.Lk + kc] ? 0.f : CA_MASKED;
    for (int qi = lane; qi < P.Lq; qi += 32) {
        const float s = caf_dot(qb + (long long)qi * P.ldq, krow, P.HD) * P.scale + add;
        const float p = expf(s - lse[(long long)bh * P.Lq + qi]);
        const bool keep = ca_keep(key, (uint32_t)qi, (uint32_t)kc, P.thresh16);
        const float dp = keep ? caf_dot(gb + (long long)qi * ldo, vrow, P.HD) * P.keep_scale : 0.f;
        pd[qi] = keep ? p * P.keep_scale : 0.f;
        ds[qi] = p * (dp - delta[(long long)bh * P.Lq + qi]);
    }
    __syncwarp();
    for (int d = lane; d < P.HD; d += 32) {
        float ak = 0.f, av = 0.f;
        for (int qi = 0; qi < P.Lq; ++qi) {
            ak = fmaf(ds[qi], qb[(long long)qi * P.ldq + d], ak);
            av = fmaf(pd[qi], gb[(long long)qi * ldo + d], av);
        }
        const long long off = ((long long)b * P.Lk + kc) * lddkv + h * P.HD + d;
        dk[off] = ak * P.scale;
        dv[off] = av;
    }
}

__global__ void cross_attn_mask_kernel(uint8_t* keep, int H, int Lq, int Lk, long long n, uint32_t key0, const unsigned long long* seed_off,
                                       uint32_t thresh16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int kc = (int)(i % Lk), qi = (int)((i / Lk) % Lq);
        const uint32_t bh = (uint32_t)(i / ((long long)Lk * Lq));
        keep[i] = ca_keep(ca_bh_key(key0, seed_off, bh), (uint32_t)qi, (uint32_t)kc, thresh16) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------ masked mean pooling
// pooled[b] = (sum_{valid rows} x1[b] + sum_{valid rows} x2[b]) / (n1 + n2)   (models/mm_model.py:572-576)
constexpr int POOL_RG = 8;          // row groups per CTA (threadIdx.y)
template <typename T>
__global__ void __launch_bounds__(32 * POOL_RG) masked_pool_fwd_kernel(const T* __restrict__ x1, const uint8_t* __restrict__ m1, int L1,
                                                                      const T* __restrict__ x2, const uint8_t* __restrict__ m2, int L2,
                                                                      float* __restrict__ out, float* __restrict__ inv_cnt, int D) {
    // CTA = (molecule b, 128 columns): thread (x, y) owns columns 4x .. 4x+3 of the rows y, y + POOL_RG, ...
    __shared__ float part[POOL_RG][128];
    __shared__ int cnt_s[POOL_RG];
    const int b = blockIdx.x, tx = threadIdx.x, ty = threadIdx.y, d = blockIdx.y * 128 + tx * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    for (int r = ty; r < L1 + L2; r += POOL_RG) {
        const bool first = r < L1;
        const bool on = first ? m1[(long long)b * L1 + r] : m2[(long long)b * L2 + (r - L1)];
        if (!on) continue;
        ++cnt;
        if (d < D) {
            const T* row = first ? x1 + ((long long)b * L1 + r) * D + d : x2 + ((long long)b * L2 + (r - L1)) * D + d;
            if constexpr (sizeof(T) == 4) {
                const float4 v = *reinterpret_cast<const float4*>(row);
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            } else {
                const uint2 u = *reinterpret_cast<const uint2*>(row);
                const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y);
                acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) part[ty][tx * 4 + e] = acc[e];
    if (tx == 0) cnt_s[ty] = cnt;
    __syncthreads();
    if (ty == 0) {
        int c = 0;
#pragma unroll
        for (int k = 0; k < POOL_RG; ++k) c += cnt_s[k];
        const float inv = 1.f / (float)c;
        if (tx == 0 && blockIdx.y == 0) inv_cnt[b] = inv;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < POOL_RG; ++k) t += part[k][tx * 4 + e];
            if (d + e < D) out[(long long)b * D + d + e] = t * inv;
        }
    }
}
__global__ void __launch_bounds__(128) masked_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ inv_cnt,
                                                              const uint8_t* __restrict__ m1, int L1, const uint8_t* __restrict__ m2, int L2,
                                                              float* __restrict__ dx1, float* __restrict__ dx2, int D) {
    const int b = blockIdx.x;
    const float inv = inv_cnt[b];
    const int r = blockIdx.y;                                 // row of the concatenation
    const bool first = r < L1;
    const bool on = first ? m1[(long long)b * L1 + r] : m2[(long long)b * L2 + (r - L1)];
    float* dst = first ? dx1 + ((long long)b * L1 + r) * D : dx2 + ((long long)b * L2 + (r - L1)) * D;
    for (int d = threadIdx.x * 4; d < D; d += 512) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) {
            v = *reinterpret_cast<const float4*>(dout + (long long)b * D + d);
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        }
        *reinterpret_cast<float4*>(dst + d) = v;
    }
}

int ca_check(const void* q, const void* k, const void* v, const uint8_t* kmask, int B, int H, int Lq, int Lk, int hd, int act_dtype, long long ldq,
             long long ldkv) {
    MMDTI_REQUIRE(q && k && v && kmask, "cross_attn: null pointer");
    MMDTI_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0, "cross_attn: bad shape B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
    MMDTI_REQUIRE(act_dtype == MMDTI_F32 || act_dtype == MMDTI_BF16, "cross_attn: act_dtype must be f32 or bf16");
    MMDTI_REQUIRE(Lq < (1 << 21) && Lk < (1 << 11), "cross_attn: Lq < 2^21 and Lk < 2^11 (dropout counter layout)");
    if (act_dtype == MMDTI_BF16) {
        MMDTI_REQUIRE(hd == 32 || hd == 64, "cross_attn: bf16 path supports head_dim 32 or 64 (got %d)", hd);
        MMDTI_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && mmdti_aligned(q, 16) && mmdti_aligned(k, 16) && mmdti_aligned(v, 16),
                      "cross_attn: bf16 operands need 16-byte aligned rows");
    } else {
        MMDTI_REQUIRE(hd > 0 && hd <= 256, "cross_attn: head_dim out of range");
        MMDTI_REQUIRE(Lq <= 1536, "cross_attn: the f32 validation path stages 2 * Lq floats per warp in 48 KB of shared memory (Lq = %d)", Lq);
    }
    return MMDTI_OK;
}
}  // namespace

extern "C" int mmdti_cross_attn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* key_mask, void* o,
                                    int64_t ldo, float* lse, int B, int H, int Lq, int Lk, int head_dim, float scale, float dropout_p,
                                    uint64_t seed, int act_dtype, void* stream) {
    if (int rc = ca_check(q, k, v, key_mask, B, H, Lq, Lk, head_dim, act_dtype, ldq, ldkv)) return rc;
    MMDTI_REQUIRE(o && lse, "cross_attn_fwd: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t th;
    float ks;
    ca_drop_params(dropout_p, th, ks);
    if (act_dtype == MMDTI_BF16) {
        MMDTI_REQUIRE(ldo % 8 == 0 && mmdti_aligned(o, 16), "cross_attn_fwd: output rows must be 16-byte aligned");
        CAParams P{static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v), ldq, ldkv, key_mask, H, Lq, Lk,
                   scale, scale * CA_LOG2E, ks, ca_seed_key(seed), th, mmdti_seed_offset_ptr()};
        const dim3 grid(B * H, (Lq + 63) / 64);          // (b, h) on x: B * H may exceed the 65535 limit of y
        if (head_dim == 32) cross_attn_fwd_kernel<32><<<grid, 128, 0, st>>>(P, static_cast<bf16*>(o), ldo, lse);
        else cross_attn_fwd_kernel<64><<<grid, 128, 0, st>>>(P, static_cast<bf16*>(o), ldo, lse);
    } else {
        CAParamsF P{static_cast<const float*>(q), static_cast<const float*>(k), static_cast<const float*>(v), ldq, ldkv, key_mask, H, Lq, Lk,
                    head_dim, scale, ks, ca_seed_key(seed), th, mmdti_seed_offset_ptr()};
        const long long nrows = (long long)B * H * Lq;
        const size_t smem = (size_t)CAF_WARPS * Lk * sizeof(float);
        cross_attn_f32_fwd_kernel<<<(unsigned)((nrows + CAF_WARPS - 1) / CAF_WARPS), CAF_WARPS * 32, smem, st>>>(P, static_cast<float*>(o), ldo, lse, nrows);
    }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_cross_attn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* key_mask, const void* o,
                                    const void* d_o, int64_t ldo, const float* lse, float* delta, void* dq, int64_t lddq, void* dk, void* dv,
                                    int64_t lddkv, int B, int H, int Lq, int Lk, int head_dim, float scale, float dropout_p, uint64_t seed,
                                    int act_dtype, void* stream) {
    if (int rc = ca_check(q, k, v, key_mask, B, H, Lq, Lk, head_dim, act_dtype, ldq, ldkv)) return rc;
    MMDTI_REQUIRE(o && d_o && lse && delta && dq && dk && dv, "cross_attn_bwd: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t th;
    float ks;
    ca_drop_params(dropout_p, th, ks);
    if (act_dtype == MMDTI_BF16) {
        MMDTI_REQUIRE(ldo % 8 == 0 && lddq % 8 == 0 && lddkv % 8 == 0 && mmdti_aligned(o, 16) && mmdti_aligned(d_o, 16),
                      "cross_attn_bwd: rows must be 16-byte aligned");
        CAParams P{static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v), ldq, ldkv, key_mask, H, Lq, Lk,
                   scale, scale * CA_LOG2E, ks, ca_seed_key(seed), th, mmdti_seed_offset_ptr()};
        const dim3 gq(B * H, (Lq + 63) / 64), gk(B * H, (Lk + 63) / 64);
        if (head_dim == 32) {
            cross_attn_bwd_q_kernel<32><<<gq, 128, 0, st>>>(P, static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), ldo, lse, delta,
                                                            static_cast<bf16*>(dq), lddq);
            cross_attn_bwd_kv_kernel<32><<<gk, 128, 0, st>>>(P, static_cast<const bf16*>(d_o), ldo, lse, delta, static_cast<bf16*>(dk),
                                                             static_cast<bf16*>(dv), lddkv);
        } else {
            cross_attn_bwd_q_kernel<64><<<gq, 128, 0, st>>>(P, static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), ldo, lse, delta,
                                                            static_cast<bf16*>(dq), lddq);
            cross_attn_bwd_kv_kernel<64><<<gk, 128, 0, st>>>(P, static_cast<const bf16*>(d_o), ldo, lse, delta, static_cast<bf16*>(dk),
                                                             static_cast<bf16*>(dv), lddkv);
        }
    } else {
        CAParamsF P{static_cast<const float*>(q), static_cast<const float*>(k), static_cast<const float*>(v), ldq, ldkv, key_mask, H, Lq, Lk,
                    head_dim, scale, ks, ca_seed_key(seed), th, mmdti_seed_offset_ptr()};
        const long long nq = (long long)B * H * Lq, nk = (long long)B * H * Lk;
        cross_attn_f32_bwd_q_kernel<<<(unsigned)((nq + CAF_WARPS - 1) / CAF_WARPS), CAF_WARPS * 32, (size_t)CAF_WARPS * Lk * sizeof(float), st>>>(
            P, static_cast<const float*>(o), static_cast<const float*>(d_o), ldo, lse, delta, static_cast<float*>(dq), lddq, nq);
        cross_attn_f32_bwd_kv_kernel<<<(unsigned)((nk + CAF_WARPS - 1) / CAF_WARPS), CAF_WARPS * 32, (size_t)CAF_WARPS * 2 * Lq * sizeof(float), st>>>(
            P, static_cast<const float*>(d_o), ldo, lse, delta, static_cast<float*>(dk), static_cast<float*>(dv), lddkv, nk);
    }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_cross_attn_dropout_mask(uint8_t* keep, int B, int H, int Lq, int Lk, float dropout_p, uint64_t seed, void* stream) {
    MMDTI_REQUIRE(keep && B > 0 && H > 0 && Lq > 0 && Lk > 0, "cross_attn_dropout_mask: bad arguments");
    uint32_t th;
    float ks;
    ca_drop_params(dropout_p, th, ks);
    const long long n = (long long)B * H * Lq * Lk;
    const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    cross_attn_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(keep, H, Lq, Lk, n, ca_seed_key(seed), mmdti_seed_offset_ptr(), th);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_masked_pool_fwd(const void* x1, const uint8_t* mask1, int L1, const void* x2, const uint8_t* mask2, int L2, float* out,
                                     float* inv_count, int B, int D, int x_dtype, void* stream) {
    MMDTI_REQUIRE(x1 && x2 && mask1 && mask2 && out && inv_count && B > 0 && D > 0 && L1 > 0 && L2 > 0, "masked_pool_fwd: bad arguments");
    MMDTI_REQUIRE(x_dtype == MMDTI_F32 || x_dtype == MMDTI_BF16, "masked_pool_fwd: x_dtype must be f32 or bf16");
    MMDTI_REQUIRE(D % 4 == 0 && mmdti_aligned(x1, 16) && mmdti_aligned(x2, 16), "masked_pool_fwd: D %% 4 == 0 and 16-byte aligned inputs");
    const dim3 grid(B, (D + 127) / 128), block(32, POOL_RG);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (x_dtype == MMDTI_F32)
        masked_pool_fwd_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(x1), mask1, L1, static_cast<const float*>(x2), mask2, L2, out,
                                                              inv_count, D);
    else
        masked_pool_fwd_kernel<bf16><<<grid, block, 0, st>>>(static_cast<const bf16*>(x1), mask1, L1, static_cast<const bf16*>(x2), mask2, L2, out,
                                                             inv_count, D);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_masked_pool_bwd(const float* dout, const float* inv_count, const uint8_t* mask1, int L1, const uint8_t* mask2, int L2,
                                     float* dx1, float* dx2, int B, int D, void* stream) {
    MMDTI_REQUIRE(dout && inv_count && mask1 && mask2 && dx1 && dx2 && B > 0 && D > 0 && D % 4 == 0 && L1 > 0 && L2 > 0,
                  "masked_pool_bwd: bad arguments");
    const dim3 grid(B, L1 + L2);
    masked_pool_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(dout, inv_count, mask1, L1, mask2, L2, dx1, dx2, D);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
