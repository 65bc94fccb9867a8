// K3 / K4 support kernels and the fp32 (validation-mode) similarity kernels.
//
//   rownorm_fwd / rownorm_bwd   F.normalize(dim=-1) and its backward (models/infonce.py:70-71,
//                               models/contrastive.py:21-22 / 79-80 / 131-132), also emitting the
//                               zero-padded bf16 operand copy the tensor-core path consumes
//   sim_stats  (fp32 FMA)       phase 1: one pass over the similarity tiles -> per-row statistics
//   sim_grad   (fp32 FMA)       phase 2: recompute the tile, form H_ij, accumulate dA_i = sum_j H_ij b_j
//   infonce_finalize / ct_finalize   per-row statistics -> loss and the per-row coefficients phase 2 needs
//   ct_masks                    debug/test export of the boolean masks (bit-exact check)
//
// The N x N similarity matrix is never written.  The tensor-core versions of sim_stats / sim_grad
// live in sim_tc.cu and share the functors in sim_common.cuh.
#define SIM_EXP(x) expf(x)
#include "sim_common.cuh"

#include <math.h>
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

// ------------------------------------------------------------------ row normalisation
// one warp per row: xhat = x / max(||x||, eps); optional outputs: f32 copy (ld = D), bf16 copy
// (ld = Dp >= D, columns [D, Dp) zero), 1/max(||x||, eps).
__global__ void __launch_bounds__(256) rownorm_fwd_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ xh,
                                                          bf16* __restrict__ xb, int Dp, float* __restrict__ inv_norm, int N,
                                                          int D, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long row = (long long)blockIdx.x * wpb + warp; row < N; row += (long long)gridDim.x * wpb) {
        const float* xr = x + row * ldx;
        float ss = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float v = xr[c];
            ss += v * v;
        }
        ss = warp_sum(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), eps);
        for (int c = lane; c < Dp; c += 32) {
            const float v = c < D ? xr[c] * inv : 0.f;
            if (xh && c < D) xh[row * D + c] = v;
            if (xb) xb[row * Dp + c] = __float2bfloat16_rn(v);
        }
        if (inv_norm && lane == 0) inv_norm[row] = inv;
    }
}

// dx = coef * (g - xhat (xhat . g)) * inv_norm,  coef = scale * (gscale ? *gscale : 1).
// (rows whose norm was clamped by eps have xhat = x/eps and no projection term)
__global__ void __launch_bounds__(256) rownorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xh,
                                                          const float* __restrict__ inv_norm, float* __restrict__ dx,
                                                          long long lddx, int N, int D, float scale,
                                                          const float* __restrict__ gscale, float eps, int accumulate) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float coef = scale * (gscale ? *gscale : 1.f);
    for (long long row = (long long)blockIdx.x * wpb + warp; row < N; row += (long long)gridDim.x * wpb) {
        const float* gr = g + row * D;
        const float* hr = xh + row * D;
        float dot = 0.f;
        for (int c = lane; c < D; c += 32) dot += gr[c] * hr[c];
        dot = warp_sum(dot);
        const float inv = inv_norm[row];
        if (inv >= 1.f / eps) dot = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float v = coef * (gr[c] - hr[c] * dot) * inv;
            if (accumulate) dx[row * lddx + c] += v;
            else dx[row * lddx + c] = v;
        }
    }
}

// ------------------------------------------------------------------ fp32 similarity kernels
// Block = 256 threads, anchor stripe of BM = 32 rows (resident in smem, full D), key tiles of
// BN = 64 columns streamed in k-chunks of 32.  S tile: thread t owns row t/8, columns (t%8)+8c.
// PHASE 1: functor -> 5 row accumulators -> 8-lane shuffle reduce -> stats (atomics iff jsplit>1).
// PHASE 2: functor -> H tile in smem -> dA (32 x D) += H (32 x 64) . B_J (64 x D), thread t owns
//          rows (t/32)*4+r, columns lane+32c (c < 16, so D <= 512).
constexpr int BM = 32, BN = 64, KC = 32, DMAX = 512;

template <int PHASE>
__global__ void __launch_bounds__(256) sim_simt_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int M, int D,
                                                       SimAux aux, float* __restrict__ out, int jtiles_per_split, int use_atomics) {
    extern __shared__ float smem[];
    float* As = smem;                       // [BM][D+1]
    float* Bs = As + BM * (D + 1);          // [BN][KC+1]
    float* Hs = Bs + BN * (KC + 1);         // [BM][BN+1]   (phase 2)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int i0 = blockIdx.x * BM;
    const int N = aux.N;
    const int ntiles = (N + BN - 1) / BN;
    const int jt0 = blockIdx.y * jtiles_per_split, jt1 = min(ntiles, jt0 + jtiles_per_split);

    for (int idx = t; idx < BM * D; idx += 256) {
        const int r = idx / D, c = idx - r * D;
        As[r * (D + 1) + c] = (i0 + r < M) ? A[(long long)(i0 + r) * D + c] : 0.f;
    }
    const int srow = t >> 3, scol = t & 7;
    const int gi = aux.row_offset + i0 + srow;          // global anchor index of this thread's S row
    const bool row_ok = (i0 + srow) < M;
    RowAcc racc;
    racc.clear();
    float ri0 = 0.f, ri1 = 0.f;
    if (PHASE == 2 && row_ok) {
        if (aux.mode == SIM_INFONCE) ri0 = aux.rs_row[i0 + srow];
        else { ri0 = aux.rs_row[2 * (i0 + srow)]; ri1 = aux.rs_row[2 * (i0 + srow) + 1]; }
    }
    float dacc[4][16];
    if (PHASE == 2) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 16; ++c) dacc[r][c] = 0.f;
    }
    __syncthreads();

    for (int jt = jt0; jt < jt1; ++jt) {
        const int j0 = jt * BN;
        float s[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) s[c] = 0.f;
        for (int k0 = 0; k0 < D; k0 += KC) {
            __syncthreads();
            for (int idx = t; idx < BN * KC; idx += 256) {
                const int r = idx / KC, c = idx - r * KC;
                Bs[r * (KC + 1) + c] = (j0 + r < N && k0 + c < D) ? Bm[(long long)(j0 + r) * D + k0 + c] : 0.f;
            }
            __syncthreads();
            const int kn = min(KC, D - k0);
            for (int k = 0; k < kn; ++k) {
                const float a = As[srow * (D + 1) + k0 + k];
#pragma unroll
                for (int c = 0; c < 8; ++c) s[c] = fmaf(a, Bs[(scol + 8 * c) * (KC + 1) + k], s[c]);
            }
        }
        if (PHASE == 1) {
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = j0 + scol + 8 * c;
                    if (j < N) sim_stats_accum(aux, racc, gi, j, s[c]);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = j0 + scol + 8 * c;
                Hs[srow * (BN + 1) + scol + 8 * c] = (row_ok && j < N) ? sim_grad_coeff(aux, gi, j, s[c], ri0, ri1) : 0.f;
            }
            __syncthreads();
            const int jn = min(BN, N - j0);
            for (int j = 0; j < jn; ++j) {
                float h[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) h[r] = Hs[(warp * 4 + r) * (BN + 1) + j];
                const float* brow = Bm + (long long)(j0 + j) * D;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int col = lane + 32 * c;
                    if (col < D) {
                        const float b = __ldg(brow + col);
#pragma unroll
                        for (int r = 0; r < 4; ++r) dacc[r][c] = fmaf(h[r], b, dacc[r][c]);
                    }
                }
            }
        }
    }

    if (PHASE == 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            float v = racc.v[k];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (scol == 0 && row_ok) {
                float* dst = out + (long long)(i0 + srow) * SIM_NSTAT + k;
                if (use_atomics) atomicAdd(dst, v);
                else *dst = v;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = i0 + warp * 4 + r;
            if (row >= M) continue;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int col = lane + 32 * c;
                if (col < D) {
                    float* dst = out + (long long)row * D + col;
                    if (use_atomics) atomicAdd(dst, dacc[r][c]);
                    else *dst = dacc[r][c];
                }
            }
        }
    }
}

// ------------------------------------------------------------------ finalize kernels
// InfoNCE: stats (M,8) {sum exp(z - 1/t), z_ii} -> lse (M) and loss += sum_i (lse_i - z_ii) * scale
__global__ void infonce_finalize_kernel(const float* __restrict__ stats, float* __restrict__ lse, float* __restrict__ loss,
                                        int M, float inv_t, float scale) {
    float part = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float l = inv_t + logf(stats[(long long)i * SIM_NSTAT]);
        lse[i] = l;
        part += (l - stats[(long long)i * SIM_NSTAT + 1]) * scale;
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(loss, part);
}

// CT: stats (M,8) {sumP, sumN, cntP, cntN, sumPs} -> rowstat (M,2) {c_i, alpha_i}, loss += sum_i loss_i / N
//   Z_i = sumP + (N - cntP) + sumN          (the exp(0)=1 quirk, contrastive.py:53 / 108 / 165)
//   loss_i = flag_i / denom_i * (cntP log Z_i - sumPs);  c_i = flag_i / (N denom_i);  alpha_i = c_i cntP / Z_i
//   denom_i = cntP + 1 (ConR, contrastive.py:51: l_dist.le(w) counts the diagonal) | max(cntP, 1) otherwise
__global__ void ct_finalize_kernel(const float* __restrict__ stats, float* __restrict__ rowstat, float* __restrict__ loss,
                                   int M, int N, int mode, int diag_in_denom) {
    float part = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float* st = stats + (long long)i * SIM_NSTAT;
        const float sumP = st[0], sumN = st[1], cntP = st[2], cntN = st[3], sumPs = st[4];
        const float Z = sumP + ((float)N - cntP) + sumN;
        const float flag = cntN > 0.f ? 1.f : 0.f;
        float denom = mode == SIM_REGRESS ? cntP + (diag_in_denom ? 1.f : 0.f) : fmaxf(cntP, 1.f);
        float li = 0.f, c = 0.f, al = 0.f;
        if (cntP > 0.f) {
            li = flag / denom * (cntP * logf(Z) - sumPs);
            c = flag / ((float)N * denom);
            al = c * cntP / Z;
        }
        rowstat[2 * i] = c;
        rowstat[2 * i + 1] = al;
        part += li / (float)N;
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(loss, part);
}

__global__ void ct_masks_kernel(SimAux aux, unsigned char* __restrict__ pos, unsigned char* __restrict__ neg) {
    const long long n2 = (long long)aux.N * aux.N;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n2; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / aux.N), j = (int)(idx - (long long)i * aux.N);
        const PairMask m = sim_pair_mask(aux, i, j);
        pos[idx] = m.pos;
        neg[idx] = m.neg;
    }
}

}  // namespace

// shared by sim_tc.cu
int sim_check_aux(const SimAux& a, int phase) {
    if (a.mode < SIM_INFONCE || a.mode > SIM_MULTI) { mmdti_set_error("sim: bad mode %d", a.mode); return MMDTI_ERR_ARG; }
    if (a.N <= 0 || !(a.inv_t > 0.f)) { mmdti_set_error("sim: N and temperature must be positive"); return MMDTI_ERR_ARG; }
    if (a.mode == SIM_REGRESS && !(a.y && a.yhat)) { mmdti_set_error("sim: ConR needs y and yhat"); return MMDTI_ERR_ARG; }
    if ((a.mode == SIM_SINGLE || a.mode == SIM_MULTI) && !(a.key && a.C >= 1)) { mmdti_set_error("sim: SupCon/multi need integer keys (N,C)"); return MMDTI_ERR_ARG; }
    if (phase == 2 && !(a.rs_row && a.rs_col)) { mmdti_set_error("sim: phase 2 needs the row statistics"); return MMDTI_ERR_ARG; }
    return MMDTI_OK;
}

SimAux sim_make_aux(int mode, int N, int row_offset, float temperature, const float* y, const float* yhat, float w_thr,
                    float e_push, const int64_t* key, int C, float coef_multi, const float* wrow, const float* wcol,
                    const float* rs_row, const float* rs_col) {
    SimAux a;
    a.mode = mode; a.N = N; a.row_offset = row_offset; a.inv_t = 1.f / temperature;
    a.y = y; a.yhat = yhat; a.w_thr = w_thr; a.e_push = e_push;
    a.key = reinterpret_cast<const long long*>(key); a.C = C;
    a.thr_multi = C > 0 ? (float)((double)coef_multi / (double)C) : 0.f;
    a.wrow = wrow; a.wcol = wcol; a.rs_row = rs_row; a.rs_col = rs_col;
    return a;
}

static int simt_launch(int phase, const float* A, const float* B, int M, int D, const SimAux& aux, float* out, cudaStream_t st) {
    if (D < 1 || D > DMAX) { mmdti_set_error("sim (fp32 path): feature dim %d not in [1,%d]", D, DMAX); return MMDTI_ERR_ARG; }
    const int stripes = (M + BM - 1) / BM;
    const int ntiles = (aux.N + BN - 1) / BN;
    int jsplit = std::max(1, std::min(ntiles, (2 * num_sms() + stripes - 1) / stripes));
    const int per = (ntiles + jsplit - 1) / jsplit;
    jsplit = (ntiles + per - 1) / per;
    const size_t smem = ((size_t)BM * (D + 1) + (size_t)BN * (KC + 1) + (size_t)BM * (BN + 1)) * sizeof(float);
    const size_t outbytes = phase == 1 ? (size_t)M * SIM_NSTAT * sizeof(float) : (size_t)M * D * sizeof(float);
    if (jsplit > 1) MMDTI_CUDA_OK(cudaMemsetAsync(out, 0, outbytes, st));
    dim3 grid(stripes, jsplit);
    if (phase == 1) {
        MMDTI_CUDA_OK(cudaFuncSetAttribute(sim_simt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (jsplit == 1) MMDTI_CUDA_OK(cudaMemsetAsync(out, 0, outbytes, st));      // unused stat slots stay 0
        sim_simt_kernel<1><<<grid, 256, smem, st>>>(A, B, M, D, aux, out, per, jsplit > 1);
    } else {
        MMDTI_CUDA_OK(cudaFuncSetAttribute(sim_simt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sim_simt_kernel<2><<<grid, 256, smem, st>>>(A, B, M, D, aux, out, per, jsplit > 1);
    }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_rownorm_fwd(const float* x, int64_t ldx, float* xhat_f32, void* xhat_bf16, int Dp, float* inv_norm, int N,
                                 int D, float eps, void* stream) {
    MMDTI_REQUIRE(x && N > 0 && D > 0 && ldx >= D, "rownorm_fwd: bad arguments");
    MMDTI_REQUIRE(!xhat_bf16 || Dp >= D, "rownorm_fwd: Dp must be >= D");
    if (!xhat_bf16) Dp = D;
    rownorm_fwd_kernel<<<std::min((N + 7) / 8, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, xhat_f32, static_cast<bf16*>(xhat_bf16), Dp, inv_norm, N, D, eps);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_rownorm_bwd(const float* g, const float* xhat, const float* inv_norm, float* dx, int64_t lddx, int N, int D,
                                 float scale, const float* gscale, float eps, int accumulate, void* stream) {
    MMDTI_REQUIRE(g && xhat && inv_norm && dx && N > 0 && D > 0 && lddx >= D, "rownorm_bwd: bad arguments");
    rownorm_bwd_kernel<<<std::min((N + 7) / 8, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g, xhat, inv_norm, dx, lddx, N, D, scale, gscale, eps, accumulate);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_sim_stats_f32(const float* A, const float* B, int M, int N, int D, int row_offset, int mode, float temperature,
                                   const float* y, const float* yhat, float w_thr, float e_push, const int64_t* key, int C,
                                   float coef_multi, const float* wrow, const float* wcol, float* stats, void* stream) {
    MMDTI_REQUIRE(A && B && stats && M > 0, "sim_stats_f32: bad arguments");
    const SimAux aux = sim_make_aux(mode, N, row_offset, temperature, y, yhat, w_thr, e_push, key, C, coef_multi, wrow, wcol, nullptr, nullptr);
    if (int rc = sim_check_aux(aux, 1)) return rc;
    return simt_launch(1, A, B, M, D, aux, stats, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_sim_grad_f32(const float* A, const float* B, int M, int N, int D, int row_offset, int mode, float temperature,
                                  const float* y, const float* yhat, float w_thr, float e_push, const int64_t* key, int C,
                                  float coef_multi, const float* wrow, const float* wcol, const float* rs_row, const float* rs_col,
                                  float* dA, void* stream) {
    MMDTI_REQUIRE(A && B && dA && M > 0, "sim_grad_f32: bad arguments");
    const SimAux aux = sim_make_aux(mode, N, row_offset, temperature, y, yhat, w_thr, e_push, key, C, coef_multi, wrow, wcol, rs_row, rs_col);
    if (int rc = sim_check_aux(aux, 2)) return rc;
    return simt_launch(2, A, B, M, D, aux, dA, static_cast<cudaStream_t>(stream));
}

extern "C" int mmdti_infonce_finalize(const float* stats, float* lse, float* loss, int M, float temperature, float scale, void* stream) {
    MMDTI_REQUIRE(stats && lse && loss && M > 0 && temperature > 0.f, "infonce_finalize: bad arguments");
    infonce_finalize_kernel<<<std::min((M + 255) / 256, num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(stats, lse, loss, M,
                                                                                                           1.f / temperature, scale);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_ct_finalize(const float* stats, float* rowstat, float* loss, int M, int N, int mode, float w_thr, void* stream) {
    MMDTI_REQUIRE(stats && rowstat && loss && M > 0 && N > 0, "ct_finalize: bad arguments");
    MMDTI_REQUIRE(mode >= SIM_REGRESS && mode <= SIM_MULTI, "ct_finalize: bad mode");
    ct_finalize_kernel<<<std::min((M + 255) / 256, num_sms()), 256, 0, static_cast<cudaStream_t>(stream)>>>(stats, rowstat, loss, M, N, mode,
                                                                                                      0.f <= w_thr ? 1 : 0);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_ct_masks(int mode, int N, const float* y, const float* yhat, float w_thr, const int64_t* key, int C,
                              float coef_multi, uint8_t* pos, uint8_t* neg, void* stream) {
    MMDTI_REQUIRE(pos && neg && N > 0, "ct_masks: bad arguments");
    const SimAux aux = sim_make_aux(mode, N, 0, 1.f, y, yhat, w_thr, 0.f, key, C, coef_multi, nullptr, nullptr, nullptr, nullptr);
    if (int rc = sim_check_aux(aux, 1)) return rc;
    MMDTI_REQUIRE(mode != SIM_INFONCE, "ct_masks: InfoNCE has no label masks");
    const long long n2 = (long long)N * N;
    ct_masks_kernel<<<(int)std::min<long long>((n2 + 255) / 256, (long long)num_sms() * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(aux, pos, neg);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
