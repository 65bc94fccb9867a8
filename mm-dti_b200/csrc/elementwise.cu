// Fused elementwise / row-reduction kernels around K2 (all HBM-bound, vectorised, one pass):
//   LayerNorm forward (+cast) and backward (+residual-gradient add, +dw/db reduction),
//   dropout + residual add (forward) and its backward (+bias-gradient column sums),
//   exact-erf GELU forward / backward (+bias-gradient column sums), column sums.
// Reference: Uni-Core TransformerEncoderLayer (pre-LN) as used by models/transformers.py:82-91,
// 136-139: x = r + dropout(out_proj(attn(LN(x)))); x = r + dropout(fc2(gelu(fc1(LN(x))))).
#include "common.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace {

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

// keep decision of flat element idx (pairs of consecutive elements share one hash)
__device__ __forceinline__ uint32_t ew_bits(uint32_t key, unsigned long long idx) {
    const unsigned long long pr = idx >> 1;
    return mix32(key ^ (uint32_t)pr ^ mix32((uint32_t)(pr >> 32) + 0x27d4eb2fU));
}
__device__ __forceinline__ bool ew_keep(uint32_t key, unsigned long long idx, uint32_t thresh16) {
    return rng_keep(ew_bits(key, idx), (uint32_t)(idx & 1), thresh16);
}

inline void drop_params(float p, uint32_t& thresh16, float& keep_scale) {
    double t = floor((double)p * 65536.0 + 0.5);
    if (t < 0) t = 0;
    if (t > 65535) t = 65535;
    thresh16 = (uint32_t)t;
    keep_scale = (float)(65536.0 / (65536.0 - t));
}

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        const float4 x = *reinterpret_cast<const float4*>(p);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        uint2 u;
        u.x = pack_bf16(v[0], v[1]);
        u.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = u;
    }
}

// ------------------------------------------------------------------ LayerNorm forward
// one warp per row; NV float4 per lane (D <= 128*NV); x f32 -> y TO; mean/rstd saved
template <typename TO, int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, TO* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                            int D, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
        const float* xr = x + row * D;
        float v[NV][4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) load4(xr + c, v[i]);
            else v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
            s += v[i][0] + v[i][1] + v[i][2] + v[i][3];
        }
        const float mu = warp_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mu; q += d * d; }
            }
        }
        const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float wv[4], bv[4], o[4];
                load4(w + c, wv);
                load4(b + c, bv);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = (v[i][e] - mu) * rs * wv[e] + bv[e];
                store4(y + row * D + c, o);
            }
        }
        if (lane == 0) {
            if (mean) mean[row] = mu;
            if (rstd) rstd[row] = rs;
        }
    }
}

// ------------------------------------------------------------------ dropout + residual + LayerNorm forward (fused)
// xo = res + dropout(a) (f32, stored: it is the next residual);  y = LN(xo) (TO);  mean/rstd saved.
// The dropout mask is the one of dropout_residual_fwd (same key, same flat index), so dropout_bwd agrees.
template <typename TA, typename TO, int NV>
__global__ void __launch_bounds__(256) dropres_layernorm_fwd_kernel(const float* __restrict__ res, const TA* __restrict__ a,
                                                                    float* __restrict__ xo, const float* __restrict__ w,
                                                                    const float* __restrict__ b, TO* __restrict__ y,
                                                                    float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                                    int D, float eps, uint32_t key, const unsigned long long* seed_off,
                                                                    uint32_t thresh16, float keep_scale) {
    key = rng_effective_key(key, seed_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
        float v[NV][4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float av[4];
                const long long idx = row * D + c;
                load4(res + idx, v[i]);
                load4(a + idx, av);
                if (thresh16) {
                    const uint32_t b0 = ew_bits(key, (unsigned long long)idx), b1 = ew_bits(key, (unsigned long long)idx + 2);
                    av[0] = rng_keep(b0, 0, thresh16) ? av[0] * keep_scale : 0.f;
                    av[1] = rng_keep(b0, 1, thresh16) ? av[1] * keep_scale : 0.f;
                    av[2] = rng_keep(b1, 0, thresh16) ? av[2] * keep_scale : 0.f;
                    av[3] = rng_keep(b1, 1, thresh16) ? av[3] * keep_scale : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) v[i][e] += av[e];
                store4(xo + idx, v[i]);
            } else v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
            s += v[i][0] + v[i][1] + v[i][2] + v[i][3];
        }
        const float mu = warp_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mu; q += d * d; }
            }
        }
        const float rs = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float wv[4], bv[4], o[4];
                load4(w + c, wv);
                load4(b + c, bv);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = (v[i][e] - mu) * rs * wv[e] + bv[e];
                store4(y + row * D + c, o);
            }
        }
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    }
}

// ------------------------------------------------------------------ LayerNorm backward + dropout backward (fused)
// dx = dx_add + dLN/dx (f32, stored: the residual gradient);  da = dropout'(dx) (TDA);  dbias += colsum(da);
// dw, db += LayerNorm parameter gradients.  Same mask as dropout_bwd (key, flat index).
template <typename TI, typename TDA, int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_dropout_kernel(const TI* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ w, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, const float* __restrict__ dx_add,
                                                                    float* __restrict__ dx, float* __restrict__ dw,
                                                                    float* __restrict__ db, TDA* __restrict__ da,
                                                                    float* __restrict__ dbias, int rows, int D, uint32_t key,
                                                                    const unsigned long long* seed_off, uint32_t thresh16,
                                                                    float keep_scale) {
    key = rng_effective_key(key, seed_off);
    extern __shared__ float red[];       // [wpb][3][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float aw[NV][4], ab[NV][4], ad[NV][4], wv[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) aw[i][e] = ab[i][e] = ad[i][e] = wv[i][e] = 0.f;
        if (c < D) load4(w + c, wv[i]);
    }
    for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
        const float mu = mean[row], rs = rstd[row];
        float g[NV][4], xh[NV][4], addv[NV][4];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D && dx_add) load4(dx_add + row * D + c, addv[i]);
            else addv[i][0] = addv[i][1] = addv[i][2] = addv[i][3] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float dyv[4], xv[4];
                load4(dy + row * D + c, dyv);
                load4(x + row * D + c, xv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    xh[i][e] = (xv[e] - mu) * rs;
                    g[i][e] = dyv[e] * wv[i][e];
                    c1 += g[i][e];
                    c2 += g[i][e] * xh[i][e];
                    aw[i][e] += dyv[e] * xh[i][e];
                    ab[i][e] += dyv[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) g[i][e] = xh[i][e] = 0.f;
            }
        }
        c1 = warp_sum(c1) / D;
        c2 = warp_sum(c2) / D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float o[4];
                const long long idx = row * D + c;
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = addv[i][e] + rs * (g[i][e] - c1 - xh[i][e] * c2);
                store4(dx + idx, o);
                if (thresh16) {
                    const uint32_t b0 = ew_bits(key, (unsigned long long)idx), b1 = ew_bits(key, (unsigned long long)idx + 2);
                    o[0] = rng_keep(b0, 0, thresh16) ? o[0] * keep_scale : 0.f;
                    o[1] = rng_keep(b0, 1, thresh16) ? o[1] * keep_scale : 0.f;
                    o[2] = rng_keep(b1, 0, thresh16) ? o[2] * keep_scale : 0.f;
                    o[3] = rng_keep(b1, 1, thresh16) ? o[3] * keep_scale : 0.f;
                }
                store4(da + idx, o);
                if constexpr (sizeof(TDA) == 2) {           // sum what was stored (rounded), like dropout_bwd
                    const float2 f0 = unpack_bf16(pack_bf16(o[0], o[1])), f1 = unpack_bf16(pack_bf16(o[2], o[3]));
                    ad[i][0] += f0.x; ad[i][1] += f0.y; ad[i][2] += f1.x; ad[i][3] += f1.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) ad[i][e] += o[e];
                }
            }
        }
    }
    float* rw = red + (size_t)warp * 3 * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < D) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { rw[c + e] = aw[i][e]; rw[D + c + e] = ab[i][e]; rw[2 * D + c + e] = ad[i][e]; }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x * 4; c < 3 * D; c += blockDim.x * 4) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int k = 0; k < wpb; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(red + (size_t)k * 3 * D + c);
            s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
        }
        red_add_v4(c < D ? dw + c : (c < 2 * D ? db + (c - D) : dbias + (c - 2 * D)), s0, s1, s2, s3);
    }
}

// ------------------------------------------------------------------ LayerNorm backward
// dx_out = (add ? dx_add : 0) + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*w
// dw += sum_rows dy*xhat, db += sum_rows dy   (warp partials -> smem -> atomics)
template <typename TI, int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const TI* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ dx_add,
                                                            float* __restrict__ dx, float* __restrict__ dw,
                                                            float* __restrict__ db, int rows, int D) {
    extern __shared__ float red[];       // [wpb][2][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float aw[NV][4], ab[NV][4], wv[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) aw[i][e] = ab[i][e] = wv[i][e] = 0.f;
        if (c < D) load4(w + c, wv[i]);
    }
    for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
        const float mu = mean[row], rs = rstd[row];
        float g[NV][4], xh[NV][4], addv[NV][4];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D && dx_add) load4(dx_add + row * D + c, addv[i]);
            else addv[i][0] = addv[i][1] = addv[i][2] = addv[i][3] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float dyv[4], xv[4];
                load4(dy + row * D + c, dyv);
                load4(x + row * D + c, xv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    xh[i][e] = (xv[e] - mu) * rs;
                    g[i][e] = dyv[e] * wv[i][e];
                    c1 += g[i][e];
                    c2 += g[i][e] * xh[i][e];
                    aw[i][e] += dyv[e] * xh[i][e];
                    ab[i][e] += dyv[e];
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) g[i][e] = xh[i][e] = 0.f;
            }
        }
        c1 = warp_sum(c1) / D;
        c2 = warp_sum(c2) / D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = addv[i][e] + rs * (g[i][e] - c1 - xh[i][e] * c2);
                store4(dx + row * D + c, o);
            }
        }
    }
    float* rw = red + (size_t)warp * 2 * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < D) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { rw[c + e] = aw[i][e]; rw[D + c + e] = ab[i][e]; }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x * 4; c < 2 * D; c += blockDim.x * 4) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int k = 0; k < wpb; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(red + (size_t)k * 2 * D + c);
            s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
        }
        red_add_v4(c < D ? dw + c : db + (c - D), s0, s1, s2, s3);
    }
}

// ------------------------------------------------------------------ dropout + residual
// out(f32) = res(f32) + dropout(a)   (a: TI)
template <typename TI>
__global__ void __launch_bounds__(256) dropout_residual_fwd_kernel(const float* __restrict__ res, const TI* __restrict__ a,
                                                                   float* __restrict__ out, long long n4, uint32_t key, const unsigned long long* seed_off,
                                                                   uint32_t thresh16, float keep_scale) {
    key = rng_effective_key(key, seed_off);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float r[4], v[4];
        load4(res + i * 4, r);
        load4(a + i * 4, v);
        if (thresh16) {
            const uint32_t b0 = ew_bits(key, (unsigned long long)i * 4), b1 = ew_bits(key, (unsigned long long)i * 4 + 2);
            v[0] = rng_keep(b0, 0, thresh16) ? v[0] * keep_scale : 0.f;
            v[1] = rng_keep(b0, 1, thresh16) ? v[1] * keep_scale : 0.f;
            v[2] = rng_keep(b1, 0, thresh16) ? v[2] * keep_scale : 0.f;
            v[3] = rng_keep(b1, 1, thresh16) ? v[3] * keep_scale : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] += v[e];
        store4(out + i * 4, r);
    }
}

// ------------------------------------------------------------------ GELU (exact erf)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}
// bf16 tensors: exact-erf GELU with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below the bf16
// rounding of the result): 2 MUFU + ~10 FMA-class instructions instead of erff + expf (~50); exp(-x^2/2) is
// shared between erf and the normal density of the gradient.
__device__ __forceinline__ void gelu_fast_pair(float x, float& val, float& grad) { gelu_fast_both(x, val, grad); }
template <typename T> __device__ __forceinline__ float gelu_t(float x) {
    if constexpr (sizeof(T) == 4) return gelu_f(x);
    float v, g;
    gelu_fast_pair(x, v, g);
    return v;
}
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) {
    if constexpr (sizeof(T) == 4) return gelu_grad_f(x);
    float v, g;
    gelu_fast_pair(x, v, g);
    return g;
}
template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T* __restrict__ z, T* __restrict__ u, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        if constexpr (sizeof(T) == 4) { load4(z + i * 8, *reinterpret_cast<float(*)[4]>(&v[0])); load4(z + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&v[4])); }
        else {
            const uint4 q = *reinterpret_cast<const uint4*>(z + i * 8);
            float2 f;
            f = unpack_bf16(q.x); v[0] = f.x; v[1] = f.y;
            f = unpack_bf16(q.y); v[2] = f.x; v[3] = f.y;
            f = unpack_bf16(q.z); v[4] = f.x; v[5] = f.y;
            f = unpack_bf16(q.w); v[6] = f.x; v[7] = f.y;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gelu_t<T>(v[e]);
        if constexpr (sizeof(T) == 4) { store4(u + i * 8, *reinterpret_cast<float(*)[4]>(&v[0])); store4(u + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&v[4])); }
        else {
            uint4 q;
            q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
            *reinterpret_cast<uint4*>(u + i * 8) = q;
        }
    }
}
// ------------------------------------------------------------------ row map + column sums
// One traversal for the three "produce a (rows,C) tensor and its column sums" kernels:
//   OP_COLSUM      : acc += in0
//   OP_GELU_BWD    : out = in0 * gelu'(in1);           acc += out
//   OP_DROPOUT_BWD : out = mask * keep_scale * in0;    acc += out      (in0 is f32)
// A thread owns 8 consecutive columns (16-byte accesses for bf16), a CTA owns a contiguous
// block of rows and keeps UNROLL row passes in flight; per-CTA partial sums leave by one
// atomicAdd per column.
enum { OP_COLSUM = 0, OP_GELU_BWD = 1, OP_DROPOUT_BWD = 2 };

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]) {
    if constexpr (sizeof(T) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        float2 f;
        f = unpack_bf16(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16(u.w); v[6] = f.x; v[7] = f.y;
    }
}
// stores 8 values and returns them as stored (rounded to T)
template <typename T> __device__ __forceinline__ void store8(T* p, float (&v)[8]) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 u;
        u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
        float2 f;
        f = unpack_bf16(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16(u.w); v[6] = f.x; v[7] = f.y;
    }
}

template <typename TIN0, typename TIO, int OP>
__global__ void __launch_bounds__(256, 4) rowmap_colsum_kernel(const TIN0* __restrict__ in0, const TIO* __restrict__ in1,
                                                            TIO* __restrict__ out, float* __restrict__ colsum, int rows, int C,
                                                            uint32_t key, const unsigned long long* seed_off, uint32_t thresh16, float keep_scale) {
    key = rng_effective_key(key, seed_off);
    constexpr int UNROLL = 2;         // 4 CTAs/SM x 8 warps with 2 row passes in flight each beat 2 CTAs x 4 passes (latency-bound)
    extern __shared__ float part[];                   // [rpp][C] when rpp > 1
    const int tpr = C >> 3;                           // threads per row (8 columns each); host guarantees tpr <= 256
    const int rpp = blockDim.x / tpr;                 // rows per pass
    const int tcol = threadIdx.x % tpr, trow = threadIdx.x / tpr;
    const bool active = trow < rpp;
    const int rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (active) {
        for (int r0 = r_begin + trow; r0 < r_end; r0 += rpp * UNROLL) {
            float a[UNROLL][8], b[UNROLL][8];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = r0 + u * rpp;
                if (r < r_end) {
                    const long long idx = (long long)r * C + tcol * 8;
                    load8(in0 + idx, a[u]);
                    if constexpr (OP == OP_GELU_BWD) load8(in1 + idx, b[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int r = r0 + u * rpp;
                if (r < r_end) {
                    const long long idx = (long long)r * C + tcol * 8;
                    if constexpr (OP == OP_GELU_BWD) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) a[u][e] *= gelu_grad_t<TIO>(b[u][e]);
                        store8(out + idx, a[u]);
                    } else if constexpr (OP == OP_DROPOUT_BWD) {
                        if (thresh16) {
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                const uint32_t bits = ew_bits(key, (unsigned long long)idx + e);
                                a[u][e] = rng_keep(bits, 0, thresh16) ? a[u][e] * keep_scale : 0.f;
                                a[u][e + 1] = rng_keep(bits, 1, thresh16) ? a[u][e + 1] * keep_scale : 0.f;
                            }
                        }
                        store8(out + idx, a[u]);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += a[u][e];
                }
            }
        }
    }
    if (!colsum) return;
    if (rpp > 1) {
        if (active) {
#pragma unroll
            for (int e = 0; e < 8; ++e) part[trow * C + tcol * 8 + e] = acc[e];
        }
        __syncthreads();
        for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int k = 0; k < rpp; ++k) {
                const float4 v = *reinterpret_cast<const float4*>(part + k * C + c);
                s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
            }
            red_add_v4(colsum + c, s0, s1, s2, s3);
        }
    } else if (active) {
        red_add_v4(colsum + tcol * 8, acc[0], acc[1], acc[2], acc[3]);
        red_add_v4(colsum + tcol * 8 + 4, acc[4], acc[5], acc[6], acc[7]);
    }
}

__global__ void ew_mask_kernel(uint8_t* keep, long long n, uint32_t key, const unsigned long long* seed_off, uint32_t thresh16) {
    key = rng_effective_key(key, seed_off);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        keep[i] = ew_keep(key, (unsigned long long)i, thresh16) ? 1 : 0;
}

// CTAs per SM of the row-reduction kernels (LayerNorm backward): every CTA ends with D..3D vector atomics, so
// fewer, longer-running CTAs win over occupancy.  MMDTI_LN_CTAS_PER_SM overrides for tuning.
inline int ln_ctas_per_sm() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("MMDTI_LN_CTAS_PER_SM");
        v = e ? atoi(e) : 2;
        if (v < 1) v = 1;
    }
    return v;
}

inline int rowmap_ctas_per_sm() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("MMDTI_ROWMAP_CTAS_PER_SM");
        v = e ? atoi(e) : 4;
        if (v < 1) v = 1;
    }
    return v;
}

// dst = (float)src (+ add): the dtype conversions / residual adds between fused GEMMs that have no epilogue to live in
// (ops_cross.py: bf16 LayerNorm output -> fp32 residual, bf16 dgrad + fp32 residual gradient).  8 elements per thread.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) convert_add_kernel(const TS* __restrict__ src, const float* __restrict__ add, TD* __restrict__ dst, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        load8(src + i * 8, v);
        if (add) {
            const float4 a = *reinterpret_cast<const float4*>(add + i * 8), b = *reinterpret_cast<const float4*>(add + i * 8 + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        store8(dst + i * 8, v);
    }
}

inline int ew_grid(long long work_items, int per_block) {
    return (int)std::max<long long>(1, std::min<long long>((work_items + per_block - 1) / per_block, (long long)num_sms() * 8));
}

}  // namespace

#define DISPATCH_NV(D, CALL)                                                        \
    if ((D) <= 128) { CALL(1); }                                                    \
    else if ((D) <= 256) { CALL(2); }                                               \
    else if ((D) <= 512) { CALL(4); }                                               \
    else { CALL(8); }

extern "C" int mmdti_layernorm_fwd(const float* x, const float* w, const float* b, void* y, float* mean, float* rstd,
                                   int rows, int D, float eps, int out_dtype, void* stream) {
    MMDTI_REQUIRE(x && w && b && y && rows > 0 && D > 0 && D % 4 == 0 && D <= 1024, "layernorm_fwd: need D %% 4 == 0 and D <= 1024 (D=%d)", D);
    MMDTI_REQUIRE(out_dtype == MMDTI_F32 || out_dtype == MMDTI_BF16, "layernorm_fwd: out_dtype must be f32 or bf16");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ew_grid(rows, 8);
#define CALL(NV)                                                                                                            \
    if (out_dtype == MMDTI_F32) layernorm_fwd_kernel<float, NV><<<grid, 256, 0, st>>>(x, w, b, static_cast<float*>(y), mean, rstd, rows, D, eps); \
    else layernorm_fwd_kernel<bf16, NV><<<grid, 256, 0, st>>>(x, w, b, static_cast<bf16*>(y), mean, rstd, rows, D, eps)
    DISPATCH_NV(D, CALL)
#undef CALL
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_layernorm_bwd(const void* dy, const float* x, const float* w, const float* mean, const float* rstd,
                                   const float* dx_add, float* dx, float* dw, float* db, int rows, int D, int dy_dtype,
                                   void* stream) {
    MMDTI_REQUIRE(dy && x && w && mean && rstd && dx && dw && db && rows > 0 && D % 4 == 0 && D <= 1024,
                  "layernorm_bwd: bad arguments (D=%d)", D);
    MMDTI_REQUIRE(dy_dtype == MMDTI_F32 || dy_dtype == MMDTI_BF16, "layernorm_bwd: dy_dtype must be f32 or bf16");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = std::min(ew_grid(rows, 8), num_sms() * ln_ctas_per_sm());
    const size_t smem = (size_t)8 * 2 * D * sizeof(float);
#define CALL(NV)                                                                                                         \
    if (dy_dtype == MMDTI_F32) {                                                                                         \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<float, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        layernorm_bwd_kernel<float, NV><<<grid, 256, smem, st>>>(static_cast<const float*>(dy), x, w, mean, rstd, dx_add, dx, dw, db, rows, D); \
    } else {                                                                                                             \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<bf16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        layernorm_bwd_kernel<bf16, NV><<<grid, 256, smem, st>>>(static_cast<const bf16*>(dy), x, w, mean, rstd, dx_add, dx, dw, db, rows, D); \
    }
    DISPATCH_NV(D, CALL)
#undef CALL
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_dropres_layernorm_fwd(const float* res, const void* a, float* xo, const float* w, const float* b, void* y,
                                          float* mean, float* rstd, int rows, int D, float eps, float p, uint64_t seed, int a_dtype,
                                          int out_dtype, void* stream) {
    MMDTI_REQUIRE(res && a && xo && w && b && y && mean && rstd && rows > 0 && D > 0 && D % 4 == 0 && D <= 1024,
                  "dropres_layernorm_fwd: need D %% 4 == 0 and D <= 1024 (D=%d)", D);
    MMDTI_REQUIRE(p >= 0.f && p < 1.f, "dropres_layernorm_fwd: p out of range");
    MMDTI_REQUIRE((a_dtype == MMDTI_F32 && out_dtype == MMDTI_F32) || (a_dtype == MMDTI_BF16 && out_dtype == MMDTI_BF16),
                  "dropres_layernorm_fwd: a / y must both be f32 or both bf16");
    uint32_t th;
    float ks;
    drop_params(p, th, ks);
    const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x165667B1U));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ew_grid(rows, 8);
#define CALL(NV)                                                                                                             \
    if (out_dtype == MMDTI_F32)                                                                                               \
        dropres_layernorm_fwd_kernel<float, float, NV><<<grid, 256, 0, st>>>(res, static_cast<const float*>(a), xo, w, b, static_cast<float*>(y), \
                                                                             mean, rstd, rows, D, eps, key, mmdti_seed_offset_ptr(), th, ks); \
    else                                                                                                                      \
        dropres_layernorm_fwd_kernel<bf16, bf16, NV><<<grid, 256, 0, st>>>(res, static_cast<const bf16*>(a), xo, w, b, static_cast<bf16*>(y), \
                                                                           mean, rstd, rows, D, eps, key, mmdti_seed_offset_ptr(), th, ks)
    DISPATCH_NV(D, CALL)
#undef CALL
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_layernorm_bwd_dropout(const void* dy, const float* x, const float* w, const float* mean, const float* rstd,
                                           const float* dx_add, float* dx, float* dw, float* db, void* da, float* dbias, int rows,
                                           int D, float p, uint64_t seed, int dy_dtype, void* stream) {
    MMDTI_REQUIRE(dy && x && w && mean && rstd && dx && dw && db && da && dbias && rows > 0 && D % 4 == 0 && D <= 1024,
                  "layernorm_bwd_dropout: bad arguments (D=%d)", D);
    MMDTI_REQUIRE(dy_dtype == MMDTI_F32 || dy_dtype == MMDTI_BF16, "layernorm_bwd_dropout: dy_dtype must be f32 or bf16");
    MMDTI_REQUIRE(p >= 0.f && p < 1.f, "layernorm_bwd_dropout: p out of range");
    uint32_t th;
    float ks;
    drop_params(p, th, ks);
    const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x165667B1U));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = std::min(ew_grid(rows, 8), num_sms() * ln_ctas_per_sm());
    const size_t smem = (size_t)8 * 3 * D * sizeof(float);
#define CALL(NV)                                                                                                         \
    if (dy_dtype == MMDTI_F32) {                                                                                         \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_dropout_kernel<float, float, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        layernorm_bwd_dropout_kernel<float, float, NV><<<grid, 256, smem, st>>>(static_cast<const float*>(dy), x, w, mean, rstd, dx_add, dx, dw, db, \
                                                                            static_cast<float*>(da), dbias, rows, D, key, mmdti_seed_offset_ptr(), th, ks); \
    } else {                                                                                                             \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_dropout_kernel<bf16, bf16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        layernorm_bwd_dropout_kernel<bf16, bf16, NV><<<grid, 256, smem, st>>>(static_cast<const bf16*>(dy), x, w, mean, rstd, dx_add, dx, dw, db, \
                                                                          static_cast<bf16*>(da), dbias, rows, D, key, mmdti_seed_offset_ptr(), th, ks); \
    }
    DISPATCH_NV(D, CALL)
#undef CALL
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_dropout_residual_fwd(const float* res, const void* a, float* out, int64_t n, float p, uint64_t seed,
                                          int a_dtype, void* stream) {
    MMDTI_REQUIRE(res && a && out && n > 0 && n % 4 == 0, "dropout_residual_fwd: n must be a positive multiple of 4");
    MMDTI_REQUIRE(p >= 0.f && p < 1.f, "dropout_residual_fwd: p out of range");
    uint32_t th;
    float ks;
    drop_params(p, th, ks);
    const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x165667B1U));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ew_grid(n / 4, 256);
    if (a_dtype == MMDTI_F32) dropout_residual_fwd_kernel<float><<<grid, 256, 0, st>>>(res, static_cast<const float*>(a), out, n / 4, key, mmdti_seed_offset_ptr(), th, ks);
    else if (a_dtype == MMDTI_BF16) dropout_residual_fwd_kernel<bf16><<<grid, 256, 0, st>>>(res, static_cast<const bf16*>(a), out, n / 4, key, mmdti_seed_offset_ptr(), th, ks);
    else { mmdti_set_error("dropout_residual_fwd: a_dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

template <typename TIN0, typename TIO, int OP>
static int launch_rowmap(const void* in0, const void* in1, void* out, float* colsum, int rows, int C, uint32_t key, uint32_t th,
                         float ks, cudaStream_t st) {
    const int tpr = C / 8;
    const int rpp = 256 / tpr;
    const size_t smem = rpp > 1 ? (size_t)rpp * C * sizeof(float) : 0;
    const int grid = (int)std::max<long long>(1, std::min<long long>((rows + 4 * rpp - 1) / (4 * rpp), (long long)num_sms() * (OP == OP_COLSUM ? std::min(2, rowmap_ctas_per_sm()) : rowmap_ctas_per_sm())));
    rowmap_colsum_kernel<TIN0, TIO, OP><<<grid, 256, smem, st>>>(static_cast<const TIN0*>(in0), static_cast<const TIO*>(in1),
                                                                  static_cast<TIO*>(out), colsum, rows, C, key, mmdti_seed_offset_ptr(), th, ks);
    return 0;
}

extern "C" int mmdti_dropout_bwd(const float* dx, void* da, float* dbias, int rows, int C, float p, uint64_t seed,
                                 int da_dtype, void* stream) {
    MMDTI_REQUIRE(dx && da && rows > 0 && C > 0 && C % 8 == 0 && C <= 2048, "dropout_bwd: need C %% 8 == 0 and C <= 2048 (C=%d)", C);
    uint32_t th;
    float ks;
    drop_params(p, th, ks);
    const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x165667B1U));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (da_dtype == MMDTI_F32) launch_rowmap<float, float, OP_DROPOUT_BWD>(dx, nullptr, da, dbias, rows, C, key, th, ks, st);
    else if (da_dtype == MMDTI_BF16) launch_rowmap<float, bf16, OP_DROPOUT_BWD>(dx, nullptr, da, dbias, rows, C, key, th, ks, st);
    else { mmdti_set_error("dropout_bwd: da_dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_convert_add(const void* src, int src_dtype, const float* add, void* dst, int dst_dtype, int64_t n, void* stream) {
    MMDTI_REQUIRE(src && dst && n > 0 && n % 8 == 0, "convert_add: n must be a positive multiple of 8");
    MMDTI_REQUIRE(mmdti_aligned(src, 16) && mmdti_aligned(dst, 16) && (!add || mmdti_aligned(add, 16)), "convert_add: 16-byte aligned buffers");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ew_grid(n / 8, 256);
    if (src_dtype == MMDTI_BF16 && dst_dtype == MMDTI_F32)
        convert_add_kernel<bf16, float><<<grid, 256, 0, st>>>(static_cast<const bf16*>(src), add, static_cast<float*>(dst), n / 8);
    else if (src_dtype == MMDTI_F32 && dst_dtype == MMDTI_BF16)
        convert_add_kernel<float, bf16><<<grid, 256, 0, st>>>(static_cast<const float*>(src), add, static_cast<bf16*>(dst), n / 8);
    else if (src_dtype == MMDTI_F32 && dst_dtype == MMDTI_F32)
        convert_add_kernel<float, float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), add, static_cast<float*>(dst), n / 8);
    else if (src_dtype == MMDTI_BF16 && dst_dtype == MMDTI_BF16)
        convert_add_kernel<bf16, bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(src), add, static_cast<bf16*>(dst), n / 8);
    else { mmdti_set_error("convert_add: dtypes must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_gelu_fwd(const void* z, void* u, int64_t n, int dtype, void* stream) {
    MMDTI_REQUIRE(z && u && n > 0 && n % 8 == 0, "gelu_fwd: n must be a positive multiple of 8");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ew_grid(n / 8, 256);
    if (dtype == MMDTI_F32) gelu_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(z), static_cast<float*>(u), n / 8);
    else if (dtype == MMDTI_BF16) gelu_fwd_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(z), static_cast<bf16*>(u), n / 8);
    else { mmdti_set_error("gelu_fwd: dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_gelu_bwd(const void* du, const void* z, void* dz, float* dbias, int rows, int C, int dtype,
                              void* stream) {
    MMDTI_REQUIRE(du && z && dz && rows > 0 && C > 0 && C % 8 == 0 && C <= 2048, "gelu_bwd: need C %% 8 == 0 and C <= 2048 (C=%d)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == MMDTI_F32) launch_rowmap<float, float, OP_GELU_BWD>(du, z, dz, dbias, rows, C, 0, 0, 1.f, st);
    else if (dtype == MMDTI_BF16) launch_rowmap<bf16, bf16, OP_GELU_BWD>(du, z, dz, dbias, rows, C, 0, 0, 1.f, st);
    else { mmdti_set_error("gelu_bwd: dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_colsum(const void* x, float* out, int rows, int C, int dtype, void* stream) {
    MMDTI_REQUIRE(x && out && rows > 0 && C > 0 && C % 8 == 0 && C <= 2048, "colsum: need C %% 8 == 0 and C <= 2048 (C=%d)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == MMDTI_F32) launch_rowmap<float, float, OP_COLSUM>(x, nullptr, nullptr, out, rows, C, 0, 0, 1.f, st);
    else if (dtype == MMDTI_BF16) launch_rowmap<bf16, bf16, OP_COLSUM>(x, nullptr, nullptr, out, rows, C, 0, 0, 1.f, st);
    else { mmdti_set_error("colsum: dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, void* stream) {
    MMDTI_REQUIRE(keep && n > 0, "dropout_mask: bad arguments");
    uint32_t th;
    float ks;
    drop_params(p, th, ks);
    const uint32_t key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x165667B1U));
    ew_mask_kernel<<<ew_grid(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(keep, n, key, mmdti_seed_offset_ptr(), th);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

// ------------------------------------------------------------------ fused multi-tensor Adam
// torch.optim.Adam semantics (tasks/trainer.py:160-162: Adam(lr, eps=1e-6), no weight decay, no amsgrad):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// ONE launch over every parameter: `table` holds per tensor {p, g, m, v, lowp, numel} (device pointers as int64),
// `chunks` maps a CTA to (tensor, first element).  The step count t lives on the device (CUDA-graph replays).
// When lowp != 0 the bf16 shadow of the updated weight is written in the same pass (the encoder's GEMM operands),
// which removes the separate fp32 -> bf16 cast pass of every step.
namespace {
constexpr int ADAM_CHUNK = 16384;        // elements per CTA
__global__ void __launch_bounds__(256) adam_multi_kernel(const long long* __restrict__ table, const int* __restrict__ chunks,
                                                         const long long* __restrict__ step_ptr, float lr, float beta1,
                                                         float beta2, float omb1, float omb2, float eps, float grad_scale) {
    const int ti = chunks[2 * blockIdx.x], start = chunks[2 * blockIdx.x + 1];
    const long long* row = table + 6LL * ti;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* m = reinterpret_cast<float*>(row[2]);
    float* v = reinterpret_cast<float*>(row[3]);
    bf16* lp = reinterpret_cast<bf16*>(row[4]);
    const long long n = row[5];
    const float t = (float)(*step_ptr);
    const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const long long end = min(n, (long long)start + ADAM_CHUNK);
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!lp || (reinterpret_cast<uintptr_t>(lp) & 7) == 0);
    if (vec) {
        for (long long i = start + threadIdx.x * 4LL; i + 3 < end; i += 256 * 4) {
            float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            const float4 gv = *reinterpret_cast<const float4*>(g + i);
            float pe[4] = {pv.x, pv.y, pv.z, pv.w}, me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
            const float ge[4] = {gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                me[e] = fmaf(beta1, me[e], omb1 * ge[e]);
                ve[e] = fmaf(beta2, ve[e], omb2 * ge[e] * ge[e]);
                pe[e] -= step_size * me[e] / (sqrtf(ve[e]) * inv_sqrt_bc2 + eps);
            }
            *reinterpret_cast<float4*>(p + i) = make_float4(pe[0], pe[1], pe[2], pe[3]);
            *reinterpret_cast<float4*>(m + i) = make_float4(me[0], me[1], me[2], me[3]);
            *reinterpret_cast<float4*>(v + i) = make_float4(ve[0], ve[1], ve[2], ve[3]);
            if (lp) {
                uint2 u;
                u.x = pack_bf16(pe[0], pe[1]);
                u.y = pack_bf16(pe[2], pe[3]);
                *reinterpret_cast<uint2*>(lp + i) = u;
            }
        }
    }
    // scalar path: unaligned tensors, and the (< 4 element) tail of an aligned chunk
    const long long tail0 = vec ? start + ((end - start) / 4) * 4 : start;
    for (long long i = tail0 + threadIdx.x; i < end; i += 256) {
        const float ge = g[i] * grad_scale;
        const float me = fmaf(beta1, m[i], omb1 * ge), ve = fmaf(beta2, v[i], omb2 * ge * ge);
        const float pe = p[i] - step_size * me / (sqrtf(ve) * inv_sqrt_bc2 + eps);
        p[i] = pe; m[i] = me; v[i] = ve;
        if (lp) lp[i] = __float2bfloat16_rn(pe);
    }
}
}  // namespace

extern "C" int mmdti_adam_chunk(void) { return ADAM_CHUNK; }

extern "C" int mmdti_adam_step(const int64_t* table, const int32_t* chunks, int nchunks, const int64_t* step, double lr, double beta1,
                               double beta2, double eps, double grad_scale, void* stream) {
    MMDTI_REQUIRE(table && chunks && step && nchunks > 0, "adam_step: bad arguments");
    MMDTI_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps > 0., "adam_step: bad hyper-parameters");
    adam_multi_kernel<<<nchunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(table), chunks,
                                                                           reinterpret_cast<const long long*>(step), (float)lr, (float)beta1,
                                                                           (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                                           (float)eps, (float)grad_scale);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
