// K1 — Gaussian pair-distance basis -> 2-layer MLP -> per-head pair bias, (B,H,L,L) layout.
//
// Reference: GaussianLayer.forward + gaussian() (models/mm_model.py:211-224,254-269),
// NonLinearHead.forward as gbf_proj (:117-128) and the permute/contiguous (:553-556).
//   u = mul[et]*dist + bias[et];   g_k = exp(-0.5((u-mu_k)/sigma_k)^2) / (sqrt(2*3.14159) sigma_k)
//   sigma_k = |std_k| + 1e-5;      out[b,h,i,j] = (W2 gelu(W1 g + b1) + b2)_h
// The (B,L,L,128) basis tensor and the (B,L,L,64) projection are never materialised: a
// persistent CTA keeps W1/W2 in shared memory (bf16), builds the basis directly in the
// mma A-fragment layout, chains both GEMMs through registers and writes the head-major
// output through a transposing smem tile so that global stores are contiguous per head.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int KB = 128;          // Gaussian kernels
constexpr int NH = 64;           // heads
constexpr int WS = KB + 8;       // smem row stride of W1/W2 (bf16): conflict-free ldmatrix
constexpr int TM = 64;           // pairs per CTA tile (16 per warp, 4 warps)
constexpr int OT_STRIDE = TM + 2;

struct BiasParams {
    const float* dist;
    const long long* et;
    const float *means, *stds, *mul, *bias, *w1, *b1, *w2, *b2;
    const unsigned char* key_pad;
    void* out;
    int B, L, Lp, E;       // Lp = row stride of the padded (B,H,L,Lp) output
    long long npairs;     // B*L*L
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// exact-erf GELU for the bf16 tensor-core path: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below
// the bf16 rounding of the result), 2 MUFU + 9 FMA-class instructions instead of erff's ~30
__device__ __forceinline__ float gelu_fast(float x) { return gelu_fast_val(x); }

// ------------------------------------------------------------------ tensor-core forward
template <typename TP>
__global__ void __launch_bounds__(128) pair_bias_fwd_tc_kernel(const BiasParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* W1s = reinterpret_cast<bf16*>(smem_raw);                  // [128][WS]
    bf16* W2s = W1s + KB * WS;                                       // [64][WS]
    float* b1s = reinterpret_cast<float*>(W2s + NH * WS);           // [128]
    float* b2s = b1s + KB;                                           // [64]
    float* mus = b2s + NH;                                           // [128]
    float* isg = mus + KB;                                           // [128] 1/sigma
    float* cof = isg + KB;                                           // [128] 1/(a sigma)
    float* muls = cof + KB;                                          // [E]
    float* biass = muls + p.E;                                       // [E]
    int* s_mol = reinterpret_cast<int*>(biass + p.E);                // [TM] molecule index (-1: no pair)
    int* s_off = s_mol + TM;                                         // [TM] irow*Lp + j inside the (b,h) tile
    TP* Ot = reinterpret_cast<TP*>(s_off + TM);                      // [64][OT_STRIDE]
    unsigned char* negf = reinterpret_cast<unsigned char*>(Ot + NH * OT_STRIDE);   // [TM] bit0: -inf key, bit1: row end
    long long* s_base = reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(negf + TM) + 7) & ~uintptr_t(7));   // [TM] offset of (b, h=0, i, j)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;

    for (int i = tid; i < KB * KB; i += blockDim.x) W1s[(i >> 7) * WS + (i & 127)] = __float2bfloat16_rn(p.w1[i]);
    for (int i = tid; i < NH * KB; i += blockDim.x) W2s[(i >> 7) * WS + (i & 127)] = __float2bfloat16_rn(p.w2[i]);
    for (int i = tid; i < KB; i += blockDim.x) {
        const float sg = fabsf(p.stds[i]) + 1e-5f;
        b1s[i] = p.b1[i];
        mus[i] = p.means[i];
        isg[i] = 1.f / sg;
        cof[i] = 1.f / (sqrtf(2.f * 3.14159f) * sg);
    }
    for (int i = tid; i < NH; i += blockDim.x) b2s[i] = p.b2[i];
    for (int i = tid; i < p.E; i += blockDim.x) { muls[i] = p.mul[i]; biass[i] = p.bias[i]; }
    __syncthreads();

    const long long LL = (long long)p.L * p.L;
    const long long ntiles = (p.npairs + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long P0 = tile * TM;
        // ---- u of this warp's 16 pairs (lanes 0..15), broadcast by shuffle
        float u = 0.f;
        __syncthreads();       // previous tile's readers are done with Ot / s_mol / s_off / negf
        if (lane < 16) {
            const long long P = P0 + warp * 16 + lane;
            int mol = -1, off = 0;
            unsigned char fl = 0;
            if (P < p.npairs) {
                long long e = p.et[P];
                if (e < 0) e = 0;
                if (e >= p.E) e = p.E - 1;
                u = fmaf(muls[e], p.dist[P], biass[e]);
                const long long bidx = P / LL;
                const int pp = (int)(P - bidx * LL), irow = pp / p.L, j = pp - irow * p.L;
                mol = (int)bidx;
                off = irow * p.Lp + j;
                if (p.key_pad && p.key_pad[bidx * p.L + j]) fl |= 1;
                if (j == p.L - 1) fl |= 2;
            }
            s_mol[warp * 16 + lane] = mol;
            s_off[warp * 16 + lane] = off;
            negf[warp * 16 + lane] = fl;
            s_base[warp * 16 + lane] = mol >= 0 ? (long long)mol * NH * ((long long)p.L * p.Lp) + off : -1;
        }
        __syncwarp();
        const float ua = __shfl_sync(0xffffffffu, u, g), ub = __shfl_sync(0xffffffffu, u, g + 8);

        // ---- GEMM1: z(16x128) = G(16x128) W1^T, basis built on the fly as A fragments
        float z[16][4];
#pragma unroll
        for (int nb = 0; nb < 16; ++nb) z[nb][0] = z[nb][1] = z[nb][2] = z[nb][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t a[4];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int k0 = kk * 16 + hf * 8 + 2 * q4;
                float ga[2], gb[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float m = mus[k0 + e], is = isg[k0 + e], c = cof[k0 + e];
                    const float ra = (ua - m) * is, rb = (ub - m) * is;
                    ga[e] = fast_ex2(ra * ra * -0.72134752044448170368f) * c;
                    gb[e] = fast_ex2(rb * rb * -0.72134752044448170368f) * c;
                }
                a[hf * 2 + 0] = pack_bf16(ga[0], ga[1]);
                a[hf * 2 + 1] = pack_bf16(gb[0], gb[1]);
            }
            // ldmatrix.x4 (non-trans): m0 = (n-block nb, k lo), m1 = (nb, k hi), m2 = (nb+1, k lo), m3 = (nb+1, k hi)
            const int mrow = lane & 7, msel = lane >> 3;
#pragma unroll
            for (int nb = 0; nb < 16; nb += 2) {
                uint32_t b0, b1, b2, b3;
                const bf16* addr = W1s + ((nb + (msel >> 1)) * 8 + mrow) * WS + kk * 16 + (msel & 1) * 8;
                ldmatrix_x4(b0, b1, b2, b3, addr);
                mma_bf16_16816(z[nb], a[0], a[1], a[2], a[3], b0, b1);
                mma_bf16_16816(z[nb + 1], a[0], a[1], a[2], a[3], b2, b3);
            }
        }
        // ---- h = gelu(z + b1) as A fragments of GEMM2; GEMM2: o(16x64) = h W2^T
        float o[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t a[4];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int nb = kk * 2 + hf, c0 = nb * 8 + 2 * q4;
                const float bb0 = b1s[c0], bb1 = b1s[c0 + 1];
                a[hf * 2 + 0] = pack_bf16(gelu_fast(z[nb][0] + bb0), gelu_fast(z[nb][1] + bb1));
                a[hf * 2 + 1] = pack_bf16(gelu_fast(z[nb][2] + bb0), gelu_fast(z[nb][3] + bb1));
            }
            const int mrow = lane & 7, msel = lane >> 3;
#pragma unroll
            for (int nb = 0; nb < 8; nb += 2) {
                uint32_t b0, b1, b2, b3;
                const bf16* addr = W2s + ((nb + (msel >> 1)) * 8 + mrow) * WS + kk * 16 + (msel & 1) * 8;
                ldmatrix_x4(b0, b1, b2, b3, addr);
                mma_bf16_16816(o[nb], a[0], a[1], a[2], a[3], b0, b1);
                mma_bf16_16816(o[nb + 1], a[0], a[1], a[2], a[3], b2, b3);
            }
        }
        // ---- transpose through smem: Ot[h][pair]
        const bool na = negf[warp * 16 + g] & 1, nbm = negf[warp * 16 + g + 8] & 1;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int h0 = nb * 8 + 2 * q4;
            const float c0 = b2s[h0], c1 = b2s[h0 + 1];
            const int pa = warp * 16 + g, pb = pa + 8;
            Ot[h0 * OT_STRIDE + pa] = from_f<TP>(na ? -INFINITY : o[nb][0] + c0);
            Ot[(h0 + 1) * OT_STRIDE + pa] = from_f<TP>(na ? -INFINITY : o[nb][1] + c1);
            Ot[h0 * OT_STRIDE + pb] = from_f<TP>(nbm ? -INFINITY : o[nb][2] + c0);
            Ot[(h0 + 1) * OT_STRIDE + pb] = from_f<TP>(nbm ? -INFINITY : o[nb][3] + c1);
        }
        __syncthreads();
        TP* out = static_cast<TP*>(p.out);
        const long long tile_elems = (long long)p.L * p.Lp;
        const TP ninf = from_f<TP>(-INFINITY);
        if (sizeof(TP) == 2 && (p.L & 1) == 0) {
            // even L, 16-bit output: pairs (P even, P+1) are adjacent keys of the same row and 4-byte aligned ->
            // one 32-bit store per two pairs; the row's -inf padding follows the odd pair of a row end
            const int npad2 = (p.Lp - p.L) >> 1;
            uint32_t ninf2;
            {
                const TP t2[2] = {ninf, ninf};
                ninf2 = *reinterpret_cast<const uint32_t*>(t2);
            }
            for (int idx = tid; idx < NH * (TM / 2); idx += blockDim.x) {
                const int h = idx >> 5, i = (idx & (TM / 2 - 1)) * 2;
                const long long b0 = s_base[i];
                if (b0 < 0) continue;
                TP* dst = out + b0 + (long long)h * tile_elems;
                if (s_base[i + 1] >= 0) {
                    *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(Ot + h * OT_STRIDE + i);
                    if (negf[i + 1] & 2)
                        for (int c = 1; c <= npad2; ++c) reinterpret_cast<uint32_t*>(dst)[c] = ninf2;
                } else {
                    *dst = Ot[h * OT_STRIDE + i];
                }
            }
        } else {
            for (int idx = tid; idx < NH * TM; idx += blockDim.x) {
                const int h = idx >> 6, i = idx & (TM - 1);
                const long long b0 = s_base[i];
                if (b0 >= 0) {
                    TP* dst = out + b0 + (long long)h * tile_elems;
                    *dst = Ot[h * OT_STRIDE + i];
                    if (negf[i] & 2) {                       // last key of its row: write the -inf padding columns
                        for (int c = 1; c <= p.Lp - p.L; ++c) dst[c] = ninf;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ fp32 validation forward
// one thread per pair, plain FMA in the reference's operation order
template <typename TP>
__global__ void __launch_bounds__(128) pair_bias_fwd_f32_kernel(const BiasParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* W1s = reinterpret_cast<float*>(smem_raw);    // [128][128]
    float* W2s = W1s + KB * KB;                          // [64][128]
    for (int i = threadIdx.x; i < KB * KB; i += blockDim.x) W1s[i] = p.w1[i];
    for (int i = threadIdx.x; i < NH * KB; i += blockDim.x) W2s[i] = p.w2[i];
    __syncthreads();
    const long long LL = (long long)p.L * p.L;
    const float a = sqrtf(2.f * 3.14159f);
    for (long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x; P < p.npairs; P += (long long)gridDim.x * blockDim.x) {
        long long e = p.et[P];
        if (e < 0) e = 0;
        if (e >= p.E) e = p.E - 1;
        const float u = p.mul[e] * p.dist[P] + p.bias[e];
        float gk[KB], hk[KB];
        for (int k = 0; k < KB; ++k) {
            const float sg = fabsf(p.stds[k]) + 1e-5f;
            const float r = (u - p.means[k]) / sg;
            gk[k] = expf(-0.5f * (r * r)) / (a * sg);
        }
        for (int n = 0; n < KB; ++n) {
            float acc = 0.f;
            for (int k = 0; k < KB; ++k) acc = fmaf(gk[k], W1s[n * KB + k], acc);
            hk[n] = gelu_erf(acc + p.b1[n]);
        }
        const long long bidx = P / LL;
        const int pp = (int)(P - bidx * LL), irow = pp / p.L, j = pp - irow * p.L;
        const bool neg = p.key_pad && p.key_pad[bidx * p.L + j];
        TP* out = static_cast<TP*>(p.out);
        const long long tile_elems = (long long)p.L * p.Lp;
        for (int h = 0; h < NH; ++h) {
            float acc = 0.f;
            for (int k = 0; k < KB; ++k) acc = fmaf(hk[k], W2s[h * KB + k], acc);
            TP* dst = out + (bidx * NH + h) * tile_elems + (long long)irow * p.Lp + j;
            *dst = from_f<TP>(neg ? -INFINITY : acc + p.b2[h]);
            if (j == p.L - 1)
                for (int c = 1; c <= p.Lp - p.L; ++c) dst[c] = from_f<TP>(-INFINITY);
        }
    }
}

// ------------------------------------------------------------------ helpers for the backward
// basis (npairs,128) in out_dtype, row-major (input of the library GEMMs of the interim backward)
template <typename TO>
__global__ void gauss_basis_kernel(const float* __restrict__ dist, const long long* __restrict__ et,
                                   const float* __restrict__ means, const float* __restrict__ stds,
                                   const float* __restrict__ mul, const float* __restrict__ bias, TO* __restrict__ out,
                                   long long npairs, int E) {
    __shared__ float mus[KB], sgs[KB];
    for (int i = threadIdx.x; i < KB; i += blockDim.x) { mus[i] = means[i]; sgs[i] = fabsf(stds[i]) + 1e-5f; }
    __syncthreads();
    const float a = sqrtf(2.f * 3.14159f);
    const int k = threadIdx.x & (KB - 1), sub = threadIdx.x >> 7;
    const int rows_per_block = blockDim.x >> 7;
    for (long long P = (long long)blockIdx.x * rows_per_block + sub; P < npairs; P += (long long)gridDim.x * rows_per_block) {
        long long e = et[P];
        if (e < 0) e = 0;
        if (e >= E) e = E - 1;
        const float u = mul[e] * dist[P] + bias[e];
        const float r = (u - mus[k]) / sgs[k];
        out[P * KB + k] = from_f<TO>(expf(-0.5f * (r * r)) / (a * sgs[k]));
    }
}

// d_out (B,H,L,L) TG -> (npairs, H) TO   [transpose of the head axis]
template <typename TG, typename TO>
__global__ void bhll_to_pairs_kernel(const TG* __restrict__ in, TO* __restrict__ out, int B, int H, int L, int Lp) {
    const long long LL = (long long)L * L;
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int h0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long pp = p0 + threadIdx.x;
        const int h = h0 + r;
        float v = 0.f;
        if (pp < LL && h < H) {
            const int irow = (int)(pp / L), j = (int)(pp - (long long)irow * L);
            v = to_f(in[(((long long)b * H + h) * L + irow) * Lp + j]);
        }
        if (!(v == v) || fabsf(v) == INFINITY) v = 0.f;
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long pp = p0 + r;
        const int h = h0 + threadIdx.x;
        if (pp < LL && h < H) out[((long long)b * LL + pp) * H + h] = from_f<TO>(tile[threadIdx.x][r]);
    }
}

// Gaussian-parameter gradients from dG (npairs,128):
//   t_k = dg_k g_k;  d mu_k += t_k r_k / sigma_k;  d std_k += sign(std_k) t_k (r_k^2 - 1)/sigma_k
//   du = -sum_k t_k r_k / sigma_k;  dmul[et] += du*dist;  dbias[et] += du
// one warp per pair (lanes over k), block-level partial sums, then atomics.
template <typename TI>
__global__ void __launch_bounds__(256) gauss_param_grad_kernel(const TI* __restrict__ dG, const float* __restrict__ dist,
                                                               const long long* __restrict__ et,
                                                               const float* __restrict__ means, const float* __restrict__ stds,
                                                               const float* __restrict__ mul, const float* __restrict__ bias,
                                                               float* __restrict__ d_means, float* __restrict__ d_stds,
                                                               float* __restrict__ d_mul, float* __restrict__ d_bias,
                                                               long long npairs, int E) {
    __shared__ float s_dmu[8][KB], s_dsd[8][KB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float a = sqrtf(2.f * 3.14159f);
    float mu[4], sg[4], sgn[4], dmu[4] = {0, 0, 0, 0}, dsd[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int k = lane + 32 * c;
        mu[c] = means[k];
        sg[c] = fabsf(stds[k]) + 1e-5f;
        sgn[c] = stds[k] > 0.f ? 1.f : (stds[k] < 0.f ? -1.f : 0.f);
    }
    for (long long P = (long long)blockIdx.x * nw + warp; P < npairs; P += (long long)gridDim.x * nw) {
        long long e = et[P];
        if (e < 0) e = 0;
        if (e >= E) e = E - 1;
        const float d = dist[P];
        const float u = mul[e] * d + bias[e];
        float du = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = lane + 32 * c;
            const float r = (u - mu[c]) / sg[c];
            const float gk = expf(-0.5f * (r * r)) / (a * sg[c]);
            const float t = to_f(dG[P * KB + k]) * gk;
            const float tr = t * r / sg[c];
            dmu[c] += tr;
            dsd[c] += sgn[c] * t * (r * r - 1.f) / sg[c];
            du -= tr;
        }
        du = warp_sum(du);
        if (lane == 0 && du != 0.f) {
            atomicAdd(d_mul + e, du * d);
            atomicAdd(d_bias + e, du);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { s_dmu[warp][lane + 32 * c] = dmu[c]; s_dsd[warp][lane + 32 * c] = dsd[c]; }
    __syncthreads();
    for (int k = threadIdx.x; k < KB; k += blockDim.x) {
        float x = 0.f, y = 0.f;
        for (int w = 0; w < nw; ++w) { x += s_dmu[w][k]; y += s_dsd[w][k]; }
        atomicAdd(d_means + k, x);
        atomicAdd(d_stds + k, y);
    }
}

// ------------------------------------------------------------------ mask fill / pair outputs
template <typename TP>
__global__ void pair_mask_fill_kernel(TP* __restrict__ pair, const unsigned char* __restrict__ key_pad, int H, int L, int ld,
                                      float fill) {
    // one CTA per (b,h); rows have stride ld (L for a dense tensor, Lp for the padded layout);
    // only masked key columns are written
    extern __shared__ unsigned char smask[];
    const int bh = blockIdx.x, b = bh / H;
    for (int j = threadIdx.x; j < L; j += blockDim.x) smask[j] = key_pad[(long long)b * L + j];
    __syncthreads();
    TP* tile = pair + (long long)bh * L * ld;
    const TP v = from_f<TP>(fill);
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        const int i = e / L, j = e - i * L;
        if (smask[j]) tile[i * ld + j] = v;
    }
}

// dense (BH,L,L) TI -> padded (BH,L,Lp) TP, padding columns = -inf
template <typename TI, typename TP>
__global__ void pair_pad_kernel(const TI* __restrict__ in, TP* __restrict__ out, int L, int Lp) {
    const long long bh = blockIdx.x;
    const TI* src = in + bh * L * L;
    TP* dst = out + bh * L * Lp;
    for (int e = threadIdx.x; e < L * Lp; e += blockDim.x) {
        const int i = e / Lp, j = e - i * Lp;
        dst[e] = j < L ? from_f<TP>(to_f(src[i * L + j])) : from_f<TP>(-INFINITY);
    }
}
// padded (BH,L,Lp) TG -> dense (BH,L,L) TO
template <typename TG, typename TO>
__global__ void pair_unpad_kernel(const TG* __restrict__ in, TO* __restrict__ out, int L, int Lp) {
    const long long bh = blockIdx.x;
    const TG* src = in + bh * L * Lp;
    TO* dst = out + bh * L * L;
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        const int i = e / L, j = e - i * L;
        dst[e] = from_f<TO>(to_f(src[i * Lp + j]));
    }
}

template <typename TP>
__global__ void pair_outputs_kernel(const TP* __restrict__ first, const TP* __restrict__ last, float* __restrict__ pair_out,
                                    float* __restrict__ delta_out, int H, int L, int Lp) {
    const long long LL = (long long)L * L;
    __shared__ float ta[32][33], tb[32][33];
    const int b = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int h0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long pp = p0 + threadIdx.x;
        const int h = h0 + r;
        float x = 0.f, d = 0.f;
        if (pp < LL && h < H) {
            const int irow = (int)(pp / L), j = (int)(pp - (long long)irow * L);
            const long long idx = (((long long)b * H + h) * L + irow) * Lp + j;
            x = to_f(last[idx]);
            // attn_mask - input_attn_mask is NaN at -inf columns and then filled with 0
            // (models/transformers.py:163-164); the input already carries the -inf there.
            d = (x == -INFINITY) ? 0.f : x - to_f(first[idx]);
        }
        ta[r][threadIdx.x] = x;
        tb[r][threadIdx.x] = d;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long pp = p0 + r;
        const int h = h0 + threadIdx.x;
        if (pp < LL && h < H) {
            const long long o = ((long long)b * LL + pp) * H + h;
            if (pair_out) pair_out[o] = ta[threadIdx.x][r];
            if (delta_out) delta_out[o] = tb[threadIdx.x][r];
        }
    }
}

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

}  // namespace

// tcgen05 forward (pair_bias_tc5.cu)
int mmdti_pair_bias_fwd_tc5(const float* dist, const long long* et, const float* means, const float* stds, const float* mul,
                            const float* bias, const float* w1, const float* b1, const float* w2, const float* b2,
                            const unsigned char* key_pad, void* out, int B, int L, int Lp, int E, int pair_dtype, cudaStream_t st);

extern "C" int mmdti_pair_bias_fwd(const float* dist, const int64_t* edge_type, const float* means, const float* stds,
                                   const float* mul, const float* bias, const float* w1, const float* b1,
                                   const float* w2, const float* b2, const uint8_t* key_pad, void* out, int B, int L,
                                   int K, int H, int E, int pair_dtype, int fp32_math, void* stream) {
    MMDTI_REQUIRE(K == KB && H == NH, "pair_bias_fwd: K must be 128 and H must be 64 (got %d, %d)", K, H);
    MMDTI_REQUIRE(B > 0 && L > 0 && E > 0, "pair_bias_fwd: empty problem");
    MMDTI_REQUIRE(dist && edge_type && means && stds && mul && bias && w1 && b1 && w2 && b2 && out, "pair_bias_fwd: null buffer");
    BiasParams p;
    p.dist = dist; p.et = reinterpret_cast<const long long*>(edge_type); p.means = means; p.stds = stds;
    p.mul = mul; p.bias = bias; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.key_pad = key_pad; p.out = out;
    p.B = B; p.L = L; p.Lp = mmdti_pair_nkb(L) * 8; p.E = E; p.npairs = (long long)B * L * L;
    MMDTI_REQUIRE(p.Lp > 0, "pair_bias_fwd: L=%d exceeds the supported maximum 264", L);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fp32_math) {
        const size_t smem = (size_t)(KB * KB + NH * KB) * sizeof(float);
        const int grid = (int)std::min<long long>((p.npairs + 127) / 128, (long long)num_sms() * 8);
#define LAUNCH_F32(TP)                                                                                              \
    {                                                                                                               \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(pair_bias_fwd_f32_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                                             \
        pair_bias_fwd_f32_kernel<TP><<<grid, 128, smem, st>>>(p);                                                   \
    }
        if (pair_dtype == MMDTI_F32) LAUNCH_F32(float)
        else if (pair_dtype == MMDTI_BF16) LAUNCH_F32(bf16)
        else if (pair_dtype == MMDTI_F16) LAUNCH_F32(__half)
        else { mmdti_set_error("pair_bias_fwd: bad pair_dtype %d", pair_dtype); return MMDTI_ERR_ARG; }
#undef LAUNCH_F32
    } else {
        // production path: tcgen05 kernel (MMDTI_K1_TC5=0 selects the previous mma.sync kernel, kept for A/B comparison)
        const char* e5 = getenv("MMDTI_K1_TC5");
        if (!e5 || atoi(e5) != 0)
            return mmdti_pair_bias_fwd_tc5(dist, p.et, means, stds, mul, bias, w1, b1, w2, b2, key_pad, out, B, L, p.Lp, E, pair_dtype, st);
        const size_t esz = pair_dtype == MMDTI_F32 ? 4 : 2;
        const size_t smem = (size_t)(KB + NH) * WS * sizeof(bf16) + (size_t)(KB * 4 + NH + 2 * E) * sizeof(float) +
                            (size_t)2 * TM * sizeof(int) + (size_t)NH * OT_STRIDE * esz + TM + 16 + (size_t)TM * sizeof(long long) + 8;
        const long long ntiles = (p.npairs + TM - 1) / TM;
        const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * 3);
#define LAUNCH_TC(TP)                                                                                              \
    {                                                                                                              \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(pair_bias_fwd_tc_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                                            \
        pair_bias_fwd_tc_kernel<TP><<<grid, 128, smem, st>>>(p);                                                   \
    }
        if (pair_dtype == MMDTI_F32) LAUNCH_TC(float)
        else if (pair_dtype == MMDTI_BF16) LAUNCH_TC(bf16)
        else if (pair_dtype == MMDTI_F16) LAUNCH_TC(__half)
        else { mmdti_set_error("pair_bias_fwd: bad pair_dtype %d", pair_dtype); return MMDTI_ERR_ARG; }
#undef LAUNCH_TC
    }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_gauss_basis(const float* dist, const int64_t* edge_type, const float* means, const float* stds,
                                 const float* mul, const float* bias, void* out, int64_t npairs, int K, int E,
                                 int out_dtype, void* stream) {
    MMDTI_REQUIRE(K == KB, "gauss_basis: K must be 128");
    MMDTI_REQUIRE(npairs > 0 && dist && edge_type && out, "gauss_basis: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (int)std::min<long long>((npairs + 1) / 2, (long long)num_sms() * 16);
    const long long* et = reinterpret_cast<const long long*>(edge_type);
    if (out_dtype == MMDTI_F32) gauss_basis_kernel<float><<<grid, 256, 0, st>>>(dist, et, means, stds, mul, bias, static_cast<float*>(out), npairs, E);
    else if (out_dtype == MMDTI_BF16) gauss_basis_kernel<bf16><<<grid, 256, 0, st>>>(dist, et, means, stds, mul, bias, static_cast<bf16*>(out), npairs, E);
    else { mmdti_set_error("gauss_basis: out_dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_pair_to_rows(const void* in, void* out, int B, int H, int L, int in_dtype, int out_dtype, void* stream) {
    MMDTI_REQUIRE(in && out && B > 0 && H > 0 && L > 0, "pair_to_rows: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long LL = (long long)L * L;
    const int Lp = mmdti_pair_nkb(L) * 8;
    MMDTI_REQUIRE(Lp > 0, "pair_to_rows: L=%d exceeds the supported maximum 264", L);
    dim3 grid((unsigned)((LL + 31) / 32), (unsigned)((H + 31) / 32), (unsigned)B), block(32, 8);
#define GO(TG, TO) bhll_to_pairs_kernel<TG, TO><<<grid, block, 0, st>>>(static_cast<const TG*>(in), static_cast<TO*>(out), B, H, L, Lp)
    if (in_dtype == MMDTI_F32 && out_dtype == MMDTI_F32) GO(float, float);
    else if (in_dtype == MMDTI_BF16 && out_dtype == MMDTI_BF16) GO(bf16, bf16);
    else if (in_dtype == MMDTI_F32 && out_dtype == MMDTI_BF16) GO(float, bf16);
    else if (in_dtype == MMDTI_BF16 && out_dtype == MMDTI_F32) GO(bf16, float);
    else if (in_dtype == MMDTI_F16 && out_dtype == MMDTI_BF16) GO(__half, bf16);
    else { mmdti_set_error("pair_to_rows: unsupported dtype combination"); return MMDTI_ERR_ARG; }
#undef GO
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_gauss_param_grad(const void* dG, const float* dist, const int64_t* edge_type, const float* means,
                                      const float* stds, const float* mul, const float* bias, float* d_means,
                                      float* d_stds, float* d_mul, float* d_bias, int64_t npairs, int K, int E,
                                      int dg_dtype, void* stream) {
    MMDTI_REQUIRE(K == KB, "gauss_param_grad: K must be 128");
    MMDTI_REQUIRE(npairs > 0 && dG && d_means && d_stds && d_mul && d_bias, "gauss_param_grad: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (int)std::min<long long>((npairs + 7) / 8, (long long)num_sms() * 8);
    const long long* et = reinterpret_cast<const long long*>(edge_type);
    if (dg_dtype == MMDTI_F32)
        gauss_param_grad_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dG), dist, et, means, stds, mul, bias, d_means, d_stds, d_mul, d_bias, npairs, E);
    else if (dg_dtype == MMDTI_BF16)
        gauss_param_grad_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(dG), dist, et, means, stds, mul, bias, d_means, d_stds, d_mul, d_bias, npairs, E);
    else { mmdti_set_error("gauss_param_grad: dg_dtype must be f32 or bf16"); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_pair_mask_fill(void* pair, const uint8_t* key_pad, int B, int H, int L, int ld, int pair_dtype,
                                    float fill, void* stream) {
    MMDTI_REQUIRE(pair && key_pad && B > 0 && H > 0 && L > 0 && ld >= L, "pair_mask_fill: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)L;
    if (pair_dtype == MMDTI_F32) pair_mask_fill_kernel<float><<<B * H, 256, smem, st>>>(static_cast<float*>(pair), key_pad, H, L, ld, fill);
    else if (pair_dtype == MMDTI_BF16) pair_mask_fill_kernel<bf16><<<B * H, 256, smem, st>>>(static_cast<bf16*>(pair), key_pad, H, L, ld, fill);
    else if (pair_dtype == MMDTI_F16) pair_mask_fill_kernel<__half><<<B * H, 256, smem, st>>>(static_cast<__half*>(pair), key_pad, H, L, ld, fill);
    else { mmdti_set_error("pair_mask_fill: bad pair_dtype %d", pair_dtype); return MMDTI_ERR_ARG; }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_pair_outputs(const void* pair_first, const void* pair_last, float* pair_out, float* delta_out, int B,
                                  int H, int L, int pair_dtype, void* stream) {
    MMDTI_REQUIRE(pair_first && pair_last && B > 0 && H > 0 && L > 0, "pair_outputs: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long LL = (long long)L * L;
    const int Lp = mmdti_pair_nkb(L) * 8;
    MMDTI_REQUIRE(Lp > 0, "pair_outputs: L=%d exceeds the supported maximum 264", L);
    dim3 grid((unsigned)((LL + 31) / 32), (unsigned)((H + 31) / 32), (unsigned)B), block(32, 8);
#define GO(TP) pair_outputs_kernel<TP><<<grid, block, 0, st>>>(static_cast<const TP*>(pair_first), static_cast<const TP*>(pair_last), pair_out, delta_out, H, L, Lp)
    if (pair_dtype == MMDTI_F32) GO(float);
    else if (pair_dtype == MMDTI_BF16) GO(bf16);
    else if (pair_dtype == MMDTI_F16) GO(__half);
    else { mmdti_set_error("pair_outputs: bad pair_dtype %d", pair_dtype); return MMDTI_ERR_ARG; }
#undef GO
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_pair_pad(const void* dense, void* padded, int BH, int L, int in_dtype, int pair_dtype, void* stream) {
    MMDTI_REQUIRE(dense && padded && BH > 0 && L > 0, "pair_pad: bad arguments");
    const int Lp = mmdti_pair_nkb(L) * 8;
    MMDTI_REQUIRE(Lp > 0, "pair_pad: L=%d exceeds the supported maximum 264", L);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GO(TI, TP) pair_pad_kernel<TI, TP><<<BH, 256, 0, st>>>(static_cast<const TI*>(dense), static_cast<TP*>(padded), L, Lp)
    if (in_dtype == MMDTI_F32 && pair_dtype == MMDTI_F32) GO(float, float);
    else if (in_dtype == MMDTI_F32 && pair_dtype == MMDTI_BF16) GO(float, bf16);
    else if (in_dtype == MMDTI_F32 && pair_dtype == MMDTI_F16) GO(float, __half);
    else if (in_dtype == MMDTI_BF16 && pair_dtype == MMDTI_BF16) GO(bf16, bf16);
    else if (in_dtype == MMDTI_BF16 && pair_dtype == MMDTI_F32) GO(bf16, float);
    else if (in_dtype == MMDTI_BF16 && pair_dtype == MMDTI_F16) GO(bf16, __half);
    else if (in_dtype == MMDTI_F16 && pair_dtype == MMDTI_F16) GO(__half, __half);
    else if (in_dtype == MMDTI_F16 && pair_dtype == MMDTI_F32) GO(__half, float);
    else if (in_dtype == MMDTI_F16 && pair_dtype == MMDTI_BF16) GO(__half, bf16);
    else { mmdti_set_error("pair_pad: unsupported dtype combination"); return MMDTI_ERR_ARG; }
#undef GO
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_pair_unpad(const void* padded, void* dense, int BH, int L, int pair_dtype, int out_dtype, void* stream) {
    MMDTI_REQUIRE(dense && padded && BH > 0 && L > 0, "pair_unpad: bad arguments");
    const int Lp = mmdti_pair_nkb(L) * 8;
    MMDTI_REQUIRE(Lp > 0, "pair_unpad: L=%d exceeds the supported maximum 264", L);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GO(TG, TO) pair_unpad_kernel<TG, TO><<<BH, 256, 0, st>>>(static_cast<const TG*>(padded), static_cast<TO*>(dense), L, Lp)
    if (pair_dtype == MMDTI_F32 && out_dtype == MMDTI_F32) GO(float, float);
    else if (pair_dtype == MMDTI_BF16 && out_dtype == MMDTI_F32) GO(bf16, float);
    else if (pair_dtype == MMDTI_F16 && out_dtype == MMDTI_F32) GO(__half, float);
    else if (pair_dtype == MMDTI_BF16 && out_dtype == MMDTI_BF16) GO(bf16, bf16);
    else if (pair_dtype == MMDTI_F16 && out_dtype == MMDTI_F16) GO(__half, __half);
    else if (pair_dtype == MMDTI_F32 && out_dtype == MMDTI_BF16) GO(float, bf16);
    else if (pair_dtype == MMDTI_F32 && out_dtype == MMDTI_F16) GO(float, __half);
    else { mmdti_set_error("pair_unpad: unsupported dtype combination"); return MMDTI_ERR_ARG; }
#undef GO
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
