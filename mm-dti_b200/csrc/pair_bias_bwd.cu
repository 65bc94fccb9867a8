// K1 backward, fused: from the gradient of the pair bias dO (B,H,L,Lp) to the gradients of every
// parameter of GaussianLayer + NonLinearHead in ONE kernel (bf16 tensor-core math, fp32 accumulation).
//
// Reference forward (models/mm_model.py:211-224,254-269,117-128,553-556):
//   u = mul[et] dist + bias[et];  g_k = N(u; mu_k, sigma_k);  z = W1 g + b1;  h = gelu(z);  o = W2 h + b2
// Backward per pair p:
//   dW2 += dO_p h_p^T, db2 += dO_p, dh = W2^T dO_p, dz = dh o gelu'(z), dW1 += dz g^T, db1 += dz,
//   dg = W1^T dz;  t_k = dg_k g_k, r_k = (u - mu_k)/sigma_k:
//   dmu_k += t_k r_k / sigma_k;  dstd_k += sign(std_k) t_k (r_k^2 - 1)/sigma_k;  du = -sum_k t_k r_k/sigma_k;
//   dmul[et] += du dist;  dbias[et] += du.
// Nothing of size (pairs x 128) ever reaches HBM: the basis, z, h, dz live in registers / shared
// memory of a persistent CTA (8 warps, 128 pairs per tile); HBM traffic is the dO read (B*H*L*Lp*2 B)
// plus dist / edge_type.
//
//   phase A  warp w owns pairs [16w, 16w+16) of the tile:
//            g (A fragments, on the fly) -> z = g W1^T -> h, gelu'(z) -> smem;  dh = dO W2 -> dz -> smem;
//            dg = dz W1 -> Gaussian-parameter terms (quad/warp shuffles -> shared-memory bins)
//   phase B  warp w owns columns [16w, 16w+16) of the weight gradients, all 128 pairs of the tile:
//            dW2[:, cols] += dO^T h,  dW1[:, cols] += dz^T g   (register accumulators across all tiles)
// Weight gradients leave through fp32 atomics once per CTA.
#include "common.cuh"

#include <math.h>
#include <algorithm>

namespace {

constexpr int KB = 128;          // Gaussian kernels / hidden width
constexpr int NH = 64;           // heads
constexpr int WS = KB + 8;       // bf16 row stride of every [*][128] smem matrix (conflict-free ldmatrix)
constexpr int TPR = 128;         // pairs per CTA tile
constexpr int NWARP = 8;

struct BwdBiasParams {
    const float* dist;
    const long long* et;
    const float *means, *stds, *mul, *bias, *w1, *b1, *w2;
    const void* d_out;               // (B,H,L,Lp) TG
    float *d_means, *d_stds, *d_mul, *d_bias, *d_w1, *d_b1, *d_w2, *d_b2;
    int B, L, Lp, E;
    long long npairs;
};

// gelu(x) = x Phi(x) and gelu'(x) = Phi(x) + x phi(x); erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7),
// sharing exp(-x^2/2) between erf and phi.
__device__ __forceinline__ void gelu_and_grad(float x, float& h, float& gp) { gelu_fast_both(x, h, gp); }

struct BwdSmem {
    bf16* W1s;      // [128][WS]   W1[hidden][k]
    bf16* W2s;      // [64][WS]    W2[head][hidden]
    bf16* dOt;      // [64][WS]    dO^T  [head][pair]
    bf16* Gs;       // [128][WS]   g     [pair][k]
    bf16* Hs;       // [128][WS]   h     [pair][hidden]   (phase A2: also holds nothing else)
    bf16* Zs;       // [128][WS]   gelu'(z) then dz  [pair][hidden]
    float *b1s, *mus, *isg, *cof, *sgn;          // [128] each
    float *muls, *biass, *dmul, *dbias;          // [E] each
    float *dmu, *dsd, *db1;                      // [128] each
    float* db2;                                  // [64]
    float *s_u, *s_d;                            // [128] per-pair u, dist
    int* s_e;                                    // [128] per-pair edge type (-1: no pair)
    long long* s_off;                            // [128] element offset of (b, h=0, i, j) in d_out (-1: no pair)
    float* wacc;                                 // [NWARP][2][128] per-warp private d_means / d_stds partials (no atomics)
    static size_t bytes(int E) {
        return sizeof(bf16) * WS * (size_t)(128 + 64 + 64 + 3 * 128) + sizeof(float) * (size_t)(5 * 128 + 4 * E + 3 * 128 + 64 + 2 * 128) +
               sizeof(int) * 128 + sizeof(long long) * 128 + sizeof(float) * NWARP * 256 + 64;
    }
    __device__ void carve(unsigned char* base, int E) {
        W1s = reinterpret_cast<bf16*>(base);
        W2s = W1s + 128 * WS;
        dOt = W2s + 64 * WS;
        Gs = dOt + 64 * WS;
        Hs = Gs + 128 * WS;
        Zs = Hs + 128 * WS;
        b1s = reinterpret_cast<float*>(Zs + 128 * WS);
        mus = b1s + 128; isg = mus + 128; cof = isg + 128; sgn = cof + 128;
        muls = sgn + 128; biass = muls + E; dmul = biass + E; dbias = dmul + E;
        dmu = dbias + E; dsd = dmu + 128; db1 = dsd + 128; db2 = db1 + 128;
        s_u = db2 + 64; s_d = s_u + 128;
        s_e = reinterpret_cast<int*>(s_d + 128);
        // 8-byte aligned: everything before is a multiple of 8 bytes when E is odd or even? keep it safe:
        s_off = reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(s_e + 128) + 7) & ~uintptr_t(7));
        wacc = reinterpret_cast<float*>(s_off + 128);
    }
};

template <typename TG>
__global__ void __launch_bounds__(NWARP * 32, 1) pair_bias_bwd_kernel(const BwdBiasParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem S;
    S.carve(smem_raw, p.E);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const int mrow = lane & 7, msel = lane >> 3;

    for (int i = tid; i < KB * KB; i += blockDim.x) S.W1s[(i >> 7) * WS + (i & 127)] = __float2bfloat16_rn(p.w1[i]);
    for (int i = tid; i < NH * KB; i += blockDim.x) S.W2s[(i >> 7) * WS + (i & 127)] = __float2bfloat16_rn(p.w2[i]);
    for (int i = tid; i < KB; i += blockDim.x) {
        const float sd = p.stds[i], sg = fabsf(sd) + 1e-5f;
        S.b1s[i] = p.b1[i];
        S.mus[i] = p.means[i];
        S.isg[i] = 1.f / sg;
        S.cof[i] = 1.f / (sqrtf(2.f * 3.14159f) * sg);
        S.sgn[i] = sd > 0.f ? 1.f : (sd < 0.f ? -1.f : 0.f);
        S.dmu[i] = S.dsd[i] = S.db1[i] = 0.f;
    }
    for (int i = tid; i < NH; i += blockDim.x) S.db2[i] = 0.f;
    for (int i = tid; i < NWARP * 256; i += blockDim.x) S.wacc[i] = 0.f;
    for (int i = tid; i < p.E; i += blockDim.x) {
        S.muls[i] = p.mul[i];
        S.biass[i] = p.bias[i];
        S.dmul[i] = S.dbias[i] = 0.f;
    }

    // persistent weight-gradient accumulators of this warp's 16-column slice
    float aw2[4][2][4], aw1[8][2][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n) aw2[m][n][0] = aw2[m][n][1] = aw2[m][n][2] = aw2[m][n][3] = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n) aw1[m][n][0] = aw1[m][n][1] = aw1[m][n][2] = aw1[m][n][3] = 0.f;
    float colacc = 0.f;                     // db1 (threads 0..127) / db2 (threads 128..191) partial

    const long long LL = (long long)p.L * p.L;
    const long long ntiles = (p.npairs + TPR - 1) / TPR;
    const TG* dout = static_cast<const TG*>(p.d_out);
    const long long tile_elems = (long long)p.L * p.Lp;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long P0 = tile * TPR;
        __syncthreads();                    // previous tile's phase B readers are done
        // ---- per-pair scalars
        if (tid < TPR) {
            const long long P = P0 + tid;
            int e = -1;
            float u = 0.f, d = 0.f;
            long long off = -1;
            if (P < p.npairs) {
                long long ee = p.et[P];
                if (ee < 0) ee = 0;
                if (ee >= p.E) ee = p.E - 1;
                e = (int)ee;
                d = p.dist[P];
                u = fmaf(S.muls[e], d, S.biass[e]);
                const long long bidx = P / LL;
                const int pp = (int)(P - bidx * LL), irow = pp / p.L, j = pp - irow * p.L;
                off = bidx * NH * tile_elems + (long long)irow * p.Lp + j;
            }
            S.s_e[tid] = e;
            S.s_u[tid] = u;
            S.s_d[tid] = d;
            S.s_off[tid] = off;
        }
        __syncthreads();
        // ---- dO^T tile [head][pair] (non-finite entries -> 0: they sit at masked keys / padding).  Thread t owns
        //      pair t & 127 and every other head: 32 independent 2-byte loads in flight per thread, no index math
        {
            const int i = tid & (TPR - 1), h0 = tid >> 7;
            const long long off = S.s_off[i];
            float v[NH / 2];
#pragma unroll
            for (int k = 0; k < NH / 2; ++k) v[k] = off >= 0 ? to_f(dout[off + (long long)(h0 + 2 * k) * tile_elems]) : 0.f;
#pragma unroll
            for (int k = 0; k < NH / 2; ++k) {
                float x = v[k];
                if (!(fabsf(x) <= 3.0e38f)) x = 0.f;
                S.dOt[(h0 + 2 * k) * WS + i] = __float2bfloat16_rn(x);
            }
        }
        __syncthreads();

        // ======================================================= phase A (16 pairs per warp)
        {
            const int pa = warp * 16 + g, pb = pa + 8;
            const float ua = S.s_u[pa], ub = S.s_u[pb];
            float z[16][4];
#pragma unroll
            for (int nb = 0; nb < 16; ++nb) z[nb][0] = z[nb][1] = z[nb][2] = z[nb][3] = 0.f;
            // A1: basis fragments + z = g W1^T; g also to smem
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                uint32_t a[4];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int k0 = kk * 16 + hf * 8 + 2 * q4;
                    float ga[2], gb[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float m = S.mus[k0 + e], is = S.isg[k0 + e], c = S.cof[k0 + e];
                        const float ra = (ua - m) * is, rb = (ub - m) * is;
                        ga[e] = fast_ex2(ra * ra * -0.72134752044448170368f) * c;
                        gb[e] = fast_ex2(rb * rb * -0.72134752044448170368f) * c;
                    }
                    a[hf * 2 + 0] = pack_bf16(ga[0], ga[1]);
                    a[hf * 2 + 1] = pack_bf16(gb[0], gb[1]);
                    *reinterpret_cast<uint32_t*>(S.Gs + pa * WS + k0) = a[hf * 2 + 0];
                    *reinterpret_cast<uint32_t*>(S.Gs + pb * WS + k0) = a[hf * 2 + 1];
                }
#pragma unroll
                for (int nb = 0; nb < 16; nb += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(b0, b1, b2, b3, S.W1s + ((nb + (msel >> 1)) * 8 + mrow) * WS + kk * 16 + (msel & 1) * 8);
                    mma_bf16_16816(z[nb], a[0], a[1], a[2], a[3], b0, b1);
                    mma_bf16_16816(z[nb + 1], a[0], a[1], a[2], a[3], b2, b3);
                }
            }
            // A2: h -> Hs, gelu'(z) -> Zs
#pragma unroll
            for (int nb = 0; nb < 16; ++nb) {
                const int c0 = nb * 8 + 2 * q4;
                const float bb0 = S.b1s[c0], bb1 = S.b1s[c0 + 1];
                float h0, h1, h2, h3, g0, g1, g2, g3;
                gelu_and_grad(z[nb][0] + bb0, h0, g0);
                gelu_and_grad(z[nb][1] + bb1, h1, g1);
                gelu_and_grad(z[nb][2] + bb0, h2, g2);
                gelu_and_grad(z[nb][3] + bb1, h3, g3);
                *reinterpret_cast<uint32_t*>(S.Hs + pa * WS + c0) = pack_bf16(h0, h1);
                *reinterpret_cast<uint32_t*>(S.Hs + pb * WS + c0) = pack_bf16(h2, h3);
                *reinterpret_cast<uint32_t*>(S.Zs + pa * WS + c0) = pack_bf16(g0, g1);
                *reinterpret_cast<uint32_t*>(S.Zs + pb * WS + c0) = pack_bf16(g2, g3);
            }
            // A3: dh = dO W2   (A = dO^T read transposed, B = W2 [head][hidden] read transposed)
#pragma unroll
            for (int nb = 0; nb < 16; ++nb) z[nb][0] = z[nb][1] = z[nb][2] = z[nb][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {               // 16 heads per step
                uint32_t a0, a1, a2, a3;
                ldmatrix_x4_trans(a0, a1, a2, a3, S.dOt + (kk * 16 + (msel >> 1) * 8 + mrow) * WS + warp * 16 + (msel & 1) * 8);
#pragma unroll
                for (int nb = 0; nb < 16; nb += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(b0, b1, b2, b3, S.W2s + (kk * 16 + (msel & 1) * 8 + mrow) * WS + (nb + (msel >> 1)) * 8);
                    mma_bf16_16816(z[nb], a0, a1, a2, a3, b0, b1);
                    mma_bf16_16816(z[nb + 1], a0, a1, a2, a3, b2, b3);
                }
            }
            // A4: dz = dh o gelu'(z) -> Zs (own entries) and packed A fragments for A5
            uint32_t dzf[8][4];
            __syncwarp();
#pragma unroll
            for (int nb = 0; nb < 16; ++nb) {
                const int c0 = nb * 8 + 2 * q4;
                const float2 ga = unpack_bf16(*reinterpret_cast<const uint32_t*>(S.Zs + pa * WS + c0));
                const float2 gb = unpack_bf16(*reinterpret_cast<const uint32_t*>(S.Zs + pb * WS + c0));
                const uint32_t wa = pack_bf16(z[nb][0] * ga.x, z[nb][1] * ga.y);
                const uint32_t wb = pack_bf16(z[nb][2] * gb.x, z[nb][3] * gb.y);
                *reinterpret_cast<uint32_t*>(S.Zs + pa * WS + c0) = wa;
                *reinterpret_cast<uint32_t*>(S.Zs + pb * WS + c0) = wb;
                dzf[nb >> 1][(nb & 1) * 2 + 0] = wa;
                dzf[nb >> 1][(nb & 1) * 2 + 1] = wb;
            }
            // A5: dg = dz W1   (B = W1 [hidden][k] read transposed)
#pragma unroll
            for (int nb = 0; nb < 16; ++nb) z[nb][0] = z[nb][1] = z[nb][2] = z[nb][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {               // 16 hidden units per step
#pragma unroll
                for (int nb = 0; nb < 16; nb += 2) {
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(b0, b1, b2, b3, S.W1s + (kk * 16 + (msel & 1) * 8 + mrow) * WS + (nb + (msel >> 1)) * 8);
                    mma_bf16_16816(z[nb], dzf[kk][0], dzf[kk][1], dzf[kk][2], dzf[kk][3], b0, b1);
                    mma_bf16_16816(z[nb + 1], dzf[kk][0], dzf[kk][1], dzf[kk][2], dzf[kk][3], b2, b3);
                }
            }
            // A6: Gaussian-parameter terms
            float dua = 0.f, dub = 0.f;
#pragma unroll
            for (int nb = 0; nb < 16; ++nb) {
                const int c0 = nb * 8 + 2 * q4;
                const float2 ga = unpack_bf16(*reinterpret_cast<const uint32_t*>(S.Gs + pa * WS + c0));
                const float2 gb = unpack_bf16(*reinterpret_cast<const uint32_t*>(S.Gs + pb * WS + c0));
                float dm[2], ds[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float m = S.mus[c0 + e], is = S.isg[c0 + e], sg = S.sgn[c0 + e];
                    const float ra = (ua - m) * is, rb = (ub - m) * is;
                    const float ta = z[nb][e] * (e ? ga.y : ga.x) * is, tb = z[nb][2 + e] * (e ? gb.y : gb.x) * is;
                    const float tra = ta * ra, trb = tb * rb;
                    dm[e] = tra + trb;
                    ds[e] = sg * (ta * fmaf(ra, ra, -1.f) + tb * fmaf(rb, rb, -1.f));
                    dua -= tra;
                    dub -= trb;
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        dm[e] += __shfl_xor_sync(0xffffffffu, dm[e], o);
                        ds[e] += __shfl_xor_sync(0xffffffffu, ds[e], o);
                    }
                }
                if (g == 0) {                 // lanes 0..3 own distinct columns of this warp's private row: plain read-modify-write
                    float* pm = S.wacc + warp * 256;
                    float2 a = *reinterpret_cast<float2*>(pm + c0), b = *reinterpret_cast<float2*>(pm + 128 + c0);
                    a.x += dm[0]; a.y += dm[1]; b.x += ds[0]; b.y += ds[1];
                    *reinterpret_cast<float2*>(pm + c0) = a;
                    *reinterpret_cast<float2*>(pm + 128 + c0) = b;
                }
            }
            dua = quad_sum(dua);
            dub = quad_sum(dub);
            if (q4 == 0) {
                const int ea = S.s_e[pa], eb = S.s_e[pb];
                if (ea >= 0 && dua != 0.f) { atomicAdd(S.dmul + ea, dua * S.s_d[pa]); atomicAdd(S.dbias + ea, dua); }
                if (eb >= 0 && dub != 0.f) { atomicAdd(S.dmul + eb, dub * S.s_d[pb]); atomicAdd(S.dbias + eb, dub); }
            }
        }
        __syncthreads();

        // ======================================================= phase B (16 weight-gradient columns per warp)
        {
            const int n0 = warp * 16;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {               // 16 pairs per step
                uint32_t bh0, bh1, bh2, bh3, bg0, bg1, bg2, bg3;
                // B operands: h[pair][n0..n0+16) and g[pair][n0..n0+16), storage [k][n] -> transposed load
                ldmatrix_x4_trans(bh0, bh1, bh2, bh3, S.Hs + (kk * 16 + (msel & 1) * 8 + mrow) * WS + n0 + (msel >> 1) * 8);
                ldmatrix_x4_trans(bg0, bg1, bg2, bg3, S.Gs + (kk * 16 + (msel & 1) * 8 + mrow) * WS + n0 + (msel >> 1) * 8);
#pragma unroll
                for (int m = 0; m < 4; ++m) {              // dW2: A = dO^T [head][pair], storage [m][k]
                    uint32_t a0, a1, a2, a3;
                    ldmatrix_x4(a0, a1, a2, a3, S.dOt + (m * 16 + (msel & 1) * 8 + mrow) * WS + kk * 16 + (msel >> 1) * 8);
                    mma_bf16_16816(aw2[m][0], a0, a1, a2, a3, bh0, bh1);
                    mma_bf16_16816(aw2[m][1], a0, a1, a2, a3, bh2, bh3);
                }
#pragma unroll
                for (int m = 0; m < 8; ++m) {              // dW1: A = dz^T [hidden][pair], storage [k][m] -> transposed
                    uint32_t a0, a1, a2, a3;
                    ldmatrix_x4_trans(a0, a1, a2, a3, S.Zs + (kk * 16 + (msel >> 1) * 8 + mrow) * WS + m * 16 + (msel & 1) * 8);
                    mma_bf16_16816(aw1[m][0], a0, a1, a2, a3, bg0, bg1);
                    mma_bf16_16816(aw1[m][1], a0, a1, a2, a3, bg2, bg3);
                }
            }
            // bias gradients: column sums of dz (threads 0..127) and of dO (threads 128..191)
            if (tid < KB) {
                float s = 0.f;
#pragma unroll 8
                for (int r = 0; r < TPR; ++r) s += __bfloat162float(S.Zs[r * WS + tid]);
                colacc += s;
            } else if (tid < KB + NH) {
                const bf16* row = S.dOt + (tid - KB) * WS;
                float s = 0.f;
#pragma unroll 8
                for (int r = 0; r < TPR; r += 2) {
                    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(row + r));
                    s += v.x + v.y;
                }
                colacc += s;
            }
        }
    }
    __syncthreads();

    // ---- flush: weight gradients (fragment layout -> global atomics), bins, vectors
    {
        const int n0 = warp * 16;
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int r0 = m * 16 + g, c = n0 + n * 8 + 2 * q4;
                atomicAdd(p.d_w2 + (long long)r0 * KB + c, aw2[m][n][0]);
                atomicAdd(p.d_w2 + (long long)r0 * KB + c + 1, aw2[m][n][1]);
                atomicAdd(p.d_w2 + (long long)(r0 + 8) * KB + c, aw2[m][n][2]);
                atomicAdd(p.d_w2 + (long long)(r0 + 8) * KB + c + 1, aw2[m][n][3]);
            }
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int r0 = m * 16 + g, c = n0 + n * 8 + 2 * q4;
                atomicAdd(p.d_w1 + (long long)r0 * KB + c, aw1[m][n][0]);
                atomicAdd(p.d_w1 + (long long)r0 * KB + c + 1, aw1[m][n][1]);
                atomicAdd(p.d_w1 + (long long)(r0 + 8) * KB + c, aw1[m][n][2]);
                atomicAdd(p.d_w1 + (long long)(r0 + 8) * KB + c + 1, aw1[m][n][3]);
            }
    }
    if (tid < KB) {
        atomicAdd(p.d_b1 + tid, colacc);
        float sm = 0.f, ss = 0.f;
        for (int w = 0; w < NWARP; ++w) { sm += S.wacc[w * 256 + tid]; ss += S.wacc[w * 256 + 128 + tid]; }
        atomicAdd(p.d_means + tid, sm);
        atomicAdd(p.d_stds + tid, ss);
    } else if (tid < KB + NH) {
        atomicAdd(p.d_b2 + (tid - KB), colacc);
    }
    for (int i = tid; i < p.E; i += blockDim.x) {
        const float a = S.dmul[i], b = S.dbias[i];
        if (a != 0.f) atomicAdd(p.d_mul + i, a);
        if (b != 0.f) atomicAdd(p.d_bias + i, b);
    }
}

int num_sms_b() {
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

extern "C" int mmdti_pair_bias_bwd(const void* d_out, const float* dist, const int64_t* edge_type, const float* means,
                                   const float* stds, const float* mul, const float* bias, const float* w1, const float* b1,
                                   const float* w2, float* d_means, float* d_stds, float* d_mul, float* d_bias, float* d_w1,
                                   float* d_b1, float* d_w2, float* d_b2, int B, int L, int K, int H, int E, int gpair_dtype,
                                   void* stream) {
    MMDTI_REQUIRE(K == KB && H == NH, "pair_bias_bwd: K must be 128 and H must be 64 (got %d, %d)", K, H);
    MMDTI_REQUIRE(B > 0 && L > 0 && E > 0, "pair_bias_bwd: empty problem");
    MMDTI_REQUIRE(d_out && dist && edge_type && means && stds && mul && bias && w1 && b1 && w2 && d_means && d_stds && d_mul &&
                      d_bias && d_w1 && d_b1 && d_w2 && d_b2,
                  "pair_bias_bwd: null buffer");
    BwdBiasParams p;
    p.dist = dist; p.et = reinterpret_cast<const long long*>(edge_type); p.means = means; p.stds = stds; p.mul = mul;
    p.bias = bias; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.d_out = d_out;
    p.d_means = d_means; p.d_stds = d_stds; p.d_mul = d_mul; p.d_bias = d_bias; p.d_w1 = d_w1; p.d_b1 = d_b1; p.d_w2 = d_w2;
    p.d_b2 = d_b2;
    p.B = B; p.L = L; p.Lp = mmdti_pair_nkb(L) * 8; p.E = E; p.npairs = (long long)B * L * L;
    MMDTI_REQUIRE(p.Lp > 0, "pair_bias_bwd: L=%d exceeds the supported maximum 264", L);
    const size_t smem = BwdSmem::bytes(E);
    MMDTI_REQUIRE(smem <= 227 * 1024, "pair_bias_bwd: %d edge types need %zu bytes of shared memory", E, smem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long ntiles = (p.npairs + TPR - 1) / TPR;
    const int grid = (int)std::min<long long>(ntiles, num_sms_b());
#define GO(TG)                                                                                                        \
    do {                                                                                                              \
        MMDTI_CUDA_OK(cudaFuncSetAttribute(pair_bias_bwd_kernel<TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pair_bias_bwd_kernel<TG><<<grid, NWARP * 32, smem, st>>>(p);                                                  \
    } while (0)
    if (gpair_dtype == MMDTI_BF16) GO(bf16);
    else if (gpair_dtype == MMDTI_F16) GO(__half);
    else if (gpair_dtype == MMDTI_F32) GO(float);
    else { mmdti_set_error("pair_bias_bwd: bad gpair_dtype %d", gpair_dtype); return MMDTI_ERR_ARG; }
#undef GO
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
