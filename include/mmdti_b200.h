/* mmdti_b200 — C ABI of the B200 (sm_100a) kernels behind the MM-DTI training hot path.
 *
 * The reference (ndlongvn/MM-DTI) is pure Python/PyTorch and has no FFI of its own; its
 * boundary for this path is a set of nn.Module / function signatures.
 * Each entry point below states which reference computation it replaces (file:line under
 * the reference tree).  The Python host side (mm-dti_b200/ops.py) binds these with ctypes;
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (MMDTI_ERR_*); the message is
 *     available from mmdti_last_error() (thread-local).  Nothing throws, nothing allocates:
 *     the caller owns every buffer (device pointers unless stated otherwise).
 *   - `stream` is a cudaStream_t passed as void*; kernels are asynchronous on it.
 *   - dtype codes: MMDTI_F32 / MMDTI_BF16 / MMDTI_F16.  "act" = activations (q,k,v,o,x),
 *     "pair" = the (B,H,L,L) pair tensor, "gpair" = its gradient.
 *   - all tensors are dense row-major; `ld*` arguments are row strides in ELEMENTS.
 *   - head_dim is fixed at 8 (Uni-Mol: 512-d / 64 heads, models/mm_model.py:325-343).
 *   - PAIR LAYOUT: inside the library the pair tensor is (B, H, L, Lp) with row stride
 *     Lp = mmdti_pair_ld(L) >= L (a multiple of 8 that is 8 mod 16): rows are 16-byte aligned and
 *     one (molecule, head) tile is one contiguous range, which is what lets a tile move with a
 *     single TMA bulk copy and sit bank-conflict-free in shared memory.  Padding columns
 *     [L, Lp) hold -inf (0 in gradients).  mmdti_pair_pad / mmdti_pair_unpad convert from / to
 *     the reference's dense (B*H, L, L) tensors.
 */
#ifndef MMDTI_B200_H
#define MMDTI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMDTI_F32 0
#define MMDTI_BF16 1
#define MMDTI_F16 2

#define MMDTI_OK 0
#define MMDTI_ERR_ARG 1
#define MMDTI_ERR_CUDA 2

#define MMDTI_HEAD_DIM 8

/* library version (major*10000 + minor*100 + patch) and last error text */
int mmdti_version(void);
const char* mmdti_last_error(void);

/* CUDA-graph support: register a device-resident uint64 counter (or NULL to clear).  Every dropout
 * kernel then derives its masks from (seed, *counter), so a captured training step that increments
 * the counter once per replay draws fresh masks each step; forward, backward and the debug mask
 * exports of one step (same counter value) still agree. */
int mmdti_set_seed_offset(const uint64_t* device_counter);

/* row stride Lp of the padded pair layout for sequence length L (-1 if L > 264) */
int mmdti_pair_ld(int L);
/* dense (BH, L, L) in_dtype -> padded (BH, L, Lp) pair_dtype (padding columns = -inf), and back */
int mmdti_pair_pad(const void* dense, void* padded, int BH, int L, int in_dtype, int pair_dtype,
                   void* stream);
int mmdti_pair_unpad(const void* padded, void* dense, int BH, int L, int pair_dtype, int out_dtype,
                     void* stream);

/* ---------------------------------------------------------------- K1: pair bias
 * Replaces GaussianLayer.forward + gaussian() + NonLinearHead.forward + permute/contiguous
 * (models/mm_model.py:211-224, 254-269, 117-128, 553-556; duplicates models/encoder.py:207-265,
 * 484-491):   u = mul[et]*dist + bias[et];  g_k = N(u; mean_k, |std_k|+1e-5) (pi=3.14159);
 *             out[b,h,i,j] = W2 gelu(W1 g + b1) + b2.
 * dist (B,L,L) f32, edge_type (B,L,L) int64, means/stds (K) f32, mul/bias (E) f32,
 * w1 (K,K) b1 (K) w2 (H,K) b2 (H) f32 (row-major, torch Linear layout), out (B,H,L,Lp) pair_dtype
 * (padded pair layout, padding columns written as -inf).
 * key_pad (B,L) uint8 or NULL: when given, padded KEY columns are written as -inf, i.e. the
 * merge of models/transformers.py:122-132 is fused into the producer.
 * K must be 128 and H must be 64. */
int mmdti_pair_bias_fwd(const float* dist, const int64_t* edge_type, const float* means,
                        const float* stds, const float* mul, const float* bias, const float* w1,
                        const float* b1, const float* w2, const float* b2, const uint8_t* key_pad,
                        void* out, int B, int L, int K, int H, int E, int pair_dtype, int fp32_math,
                        void* stream);

/* Fused backward of the above (bf16 tensor-core math, fp32 accumulation): from d_out (B,H,L,Lp)
 * gpair_dtype (non-finite entries are read as 0) ACCUMULATE (+=) the gradients of every parameter:
 * d_means,d_stds (K), d_mul,d_bias (E), d_w1 (K,K), d_b1 (K), d_w2 (H,K), d_b2 (H), all f32.
 * The basis / hidden activations are recomputed on chip; nothing of size pairs x 128 touches HBM
 * (SURVEY.md Appendix B, K1).  There is no gradient to dist / edge_type. */
int mmdti_pair_bias_bwd(const void* d_out, const float* dist, const int64_t* edge_type,
                        const float* means, const float* stds, const float* mul, const float* bias,
                        const float* w1, const float* b1, const float* w2, float* d_means,
                        float* d_stds, float* d_mul, float* d_bias, float* d_w1, float* d_b1,
                        float* d_w2, float* d_b2, int B, int L, int K, int H, int E, int gpair_dtype,
                        void* stream);

/* Pair featurisation on the device: src_distance (B,L,L) f32 and src_edge_type (B,L,L) int64 of
 * a padded batch from src_tokens (B,L) int64 and src_coord (B,L,3) f32; both outputs are 0 wherever either position
 * holds pad_idx.  Bit-exact with data/conformer.py:205-212,216-218 (float64 scipy distance_matrix cast to float32,
 * tok_i * n_dict + tok_j) followed by the zero padding of utils/util.py:41-105. */
int mmdti_featurise(const float* coord, const int64_t* tokens, int B, int L, int n_dict, int64_t pad_idx,
                    float* dist, int64_t* edge_type, void* stream);

/* Pieces of the fp32 validation-mode backward (the 128-wide GEMMs in between run as fp32 library
 * GEMMs on the host side, see mm-dti_b200/ops.py:PairBiasFn.backward):
 *  - mmdti_gauss_basis: the (npairs,128) basis g (out_dtype f32|bf16), i.e. GaussianLayer.forward
 *    alone (models/mm_model.py:254-269);
 *  - mmdti_pair_to_rows: padded (B,H,L,Lp) -> (B*L*L, H) transpose of the incoming gradient (non-finite
 *    entries are written as 0);
 *  - mmdti_gauss_param_grad: from dG (npairs,128) accumulate (+=) d_means,d_stds (K) and the
 *    961-bin scatter d_mul,d_bias (E). */
int mmdti_gauss_basis(const float* dist, const int64_t* edge_type, const float* means,
                      const float* stds, const float* mul, const float* bias, void* out,
                      int64_t npairs, int K, int E, int out_dtype, void* stream);
int mmdti_pair_to_rows(const void* in, void* out, int B, int H, int L, int in_dtype, int out_dtype,
                       void* stream);
int mmdti_gauss_param_grad(const void* dG, const float* dist, const int64_t* edge_type,
                           const float* means, const float* stds, const float* mul,
                           const float* bias, float* d_means, float* d_stds, float* d_mul,
                           float* d_bias, int64_t npairs, int K, int E, int dg_dtype, void* stream);

/* In-place merge of the key-padding mask into the pair bias: pair[b,h,i,j] = fill where
 * key_pad[b,j] != 0.  Replaces fill_attn_mask, models/transformers.py:122-132 (bit-exact).
 * ld = row stride of `pair`: L for the reference's dense tensor, Lp for the padded layout. */
int mmdti_pair_mask_fill(void* pair, const uint8_t* key_pad, int B, int H, int L, int ld,
                         int pair_dtype, float fill, void* stream);

/* ---------------------------------------------------------------- K2: pair-biased attention
 * Replaces the attention core of Uni-Core's SelfMultiheadAttention(return_attn=True) as driven
 * by models/transformers.py:136-139:
 *     S = scale * Q K^T + P_in   (P carries -inf at padded keys)        -> pair_out := S
 *     A = dropout(softmax_rows(S));  O = A V
 * q,k,v: (B*L, *) act_dtype with row stride ldqkv; head h of token t lives at
 * columns [h*8, h*8+8) of row t (so q/k/v may point INTO the in_proj output: no transposes).
 * o: (B*L, H*8) with row stride ldo.  pair_in/pair_out (B,H,L,Lp) pair_dtype (padded layout,
 * -inf in the padding columns); may alias.
 * Dropout: element (b,h,i,j) is kept iff its counter-based random number (a function of
 * seed,b,h,i,j only) is >= round(p*65536); kept values are scaled by 65536/(65536-thresh). */
int mmdti_pair_attn_fwd(const void* q, const void* k, const void* v, int64_t ldqkv,
                        const void* pair_in, void* pair_out, void* o, int64_t ldo, int B, int H,
                        int L, float scale, float dropout_p, uint64_t seed, int act_dtype,
                        int pair_dtype, void* stream);

/* Backward.  s = pair_out and o = the output of the forward call (same seed/dropout_p; o and d_o
 * share the row stride lddo; rowsum(dA o A) is taken as d_o . o).  d_pair_out may be NULL
 * (treated as 0) and must be 0 wherever s is -inf.  Pair tensors are in the padded layout.
 * Writes d_pair_in (gpair_dtype == pair_dtype; may alias d_pair_out) and dq,dk,dv
 * (act_dtype, row stride lddqkv, same head-column convention as q,k,v):
 *     dS = A o (dA - rowsum(dA o A)) + d_pair_out;  d_pair_in = dS;
 *     dQ = scale dS K;  dK = scale dS^T Q;  dV = A'^T dO. */
int mmdti_pair_attn_bwd(const void* q, const void* k, const void* v, int64_t ldqkv, const void* s,
                        const void* o, const void* d_o, int64_t lddo, const void* d_pair_out, void* d_pair_in,
                        void* dq, void* dk, void* dv, int64_t lddqkv, int B, int H, int L,
                        float scale, float dropout_p, uint64_t seed, int act_dtype, int pair_dtype,
                        int gpair_dtype, void* stream);

/* Debug/test export: the keep mask (uint8, (B,H,L,L)) mmdti_pair_attn_fwd uses for `seed`. */
int mmdti_pair_attn_dropout_mask(uint8_t* keep, int B, int H, int L, float dropout_p, uint64_t seed,
                                 void* stream);

/* ---------------------------------------------------------------- dense projections of the encoder layer (tcgen05)
 * The four linear layers of Uni-Core's TransformerEncoderLayer (reference call sites models/transformers.py:82-91,
 * 136-139: in_proj / out_proj inside SelfMultiheadAttention, fc1 / fc2) with the layer's elementwise work fused into
 * the GEMM epilogues.  All matrices bf16 row-major with row strides ld* (elements, multiples of 8), fp32 accumulation
 * in TMEM; M = tokens (B*L), N = out features, K = in features of the linear layer in EVERY call below.  W is the
 * nn.Linear weight (N, K).  Replaces torch.addmm / torch.mm + mmdti_gelu_fwd / mmdti_dropres_layernorm_fwd /
 * mmdti_gelu_bwd / mmdti_layernorm_bwd_dropout / mmdti_dropout_bwd around them. */

/* Y = X W^T + bias (bias (N) bf16, may be NULL).                                        [in_proj forward] */
int mmdti_gemm_bias(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* Y,
                    int64_t ldy, int M, int N, int K, void* stream);
/* Z = X W^T + bias;  U = gelu(Z) (exact-erf GELU).  store_grad == 0: Z is stored (bf16) and U is the GELU of the stored,
 * rounded Z (what a backward that re-evaluates gelu'(Z) needs).  store_grad != 0: the Z buffer receives gelu'(z) instead
 * (bf16), U = gelu(z) of the fp32 z: the backward then is one multiply per element (pass z_is_grad to
 * mmdti_gemm_dgrad_gelu).                                                                [fc1 forward] */
int mmdti_gemm_bias_gelu(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* Z,
                         int64_t ldz, void* U, int64_t ldu, int M, int N, int K, int store_grad, void* stream);
/* xo = res + dropout(X W^T + bias) (fp32, dense (M,N));  Y = LayerNorm(xo; ln_w, ln_b, eps) (bf16, dense (M,N)),
 * mean/rstd (M) saved.  ln_w == NULL: no LayerNorm (Y/mean/rstd unused).  N <= 512 (whole rows per CTA pair).
 * Dropout mask = the flat-tensor mask of mmdti_dropout_mask(seed) over (M*N).           [out_proj / fc2 forward] */
int mmdti_gemm_dropres_ln(const void* X, int64_t ldx, const void* W, int64_t ldw, const void* bias, const float* res,
                          float* xo, const float* ln_w, const float* ln_b, void* Y, float* mean, float* rstd,
                          int M, int N, int K, float eps, float dropout_p, uint64_t seed, void* stream);
/* dX (M,K) = dY (M,N) W (N,K).                                                          [out_proj backward] */
int mmdti_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dX, int64_t lddx, int M, int N,
                     int K, void* stream);
/* dZ (M,K) = (dY W) * gelu'(Z)  (z_is_grad != 0: the Z buffer already holds gelu'(z), see mmdti_gemm_bias_gelu);
 * dbias (K) += column sums of the stored dZ.                                             [fc2 backward -> fc1 output] */
int mmdti_gemm_dgrad_gelu(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* Z, int64_t ldz,
                          void* dZ, int64_t lddz, float* dbias, int M, int N, int K, int z_is_grad, void* stream);
/* dh = dY W (M,K), K <= 512: gradient at the output of LayerNorm(x; ln_w) whose statistics are mean/rstd;
 * dx = dx_add + LayerNorm'(dh) (fp32 dense (M,K); dx_add may be NULL);  da = dropout'(dx) (bf16 dense (M,K), mask of
 * `seed` over (M*K));  dw (K) += sum_rows dh*xhat, db (K) += sum_rows dh, dbias (K) += sum_rows da.
 *                                                                                        [fc1 / in_proj backward] */
int mmdti_gemm_dgrad_lnbwd(const void* dY, int64_t lddy, const void* W, int64_t ldw, const float* x, const float* mean,
                           const float* rstd, const float* ln_w, const float* dx_add, float* dx, float* dw, float* db,
                           void* da, float* dbias, int M, int N, int K, float dropout_p, uint64_t seed, void* stream);
/* dW (N,K) fp32 (row stride lddw) = dY^T X, or += when accumulate != 0.                 [weight gradients] */
int mmdti_gemm_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t lddw, int M, int N,
                     int K, int accumulate, void* stream);
/* C (M,N) fp32 = A (M,K) bf16 @ B (K,N) bf16, both row-major.  The second step of the two-step contrastive backward
 * (dA = H . B over the key dimension, models/infonce.py:70-98 / models/contrastive.py gradients), replacing a library GEMM. */
int mmdti_gemm_nn_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int M, int N, int K,
                      void* stream);

/* padded (B,H,L,Lp) pair_dtype -> dense (B,L,L,H) f32 "pair" and "delta pair" outputs of
 * TransformerEncoderWithPair.forward (models/transformers.py:163-172): pair_last permuted, and
 * delta = pair_last - pair_first with padded key columns (where pair_last is -inf) set to 0. */
int mmdti_pair_outputs(const void* pair_first, const void* pair_last, float* pair_out,
                       float* delta_out, int B, int H, int L, int pair_dtype, void* stream);

/* ---------------------------------------------------------------- fused elementwise around K2
 * The non-GEMM pieces of Uni-Core's pre-LN TransformerEncoderLayer (call
 * site models/transformers.py:82-91,136-139), one HBM pass each.  Dropout uses the same
 * counter-based generator as K2: element i of the flat tensor is kept iff hash(seed, i) >=
 * round(p*65536); kept values are scaled by 65536/(65536-thresh).
 *
 * LayerNorm (eps 1e-5): y = (x-mean)*rstd*w + b; x (rows,D) f32, y out_dtype (f32|bf16);
 * mean/rstd (rows) f32 are saved for the backward.  D % 4 == 0, D <= 1024. */
int mmdti_layernorm_fwd(const float* x, const float* w, const float* b, void* y, float* mean,
                        float* rstd, int rows, int D, float eps, int out_dtype, void* stream);
/* dx = (dx_add ? dx_add : 0) + dLN/dx (dx may alias dx_add); dw,db (D) f32 are ACCUMULATED. */
int mmdti_layernorm_bwd(const void* dy, const float* x, const float* w, const float* mean,
                        const float* rstd, const float* dx_add, float* dx, float* dw, float* db,
                        int rows, int D, int dy_dtype, void* stream);
/* Fused pairs used by the encoder layer (one HBM pass instead of two each):
 *   dropres_layernorm_fwd: xo = res + dropout(a) (f32, stored);  y = LN(xo) (f32|bf16 = dtype of a);
 *   layernorm_bwd_dropout: dx = dx_add + dLN/dx (f32, stored);  da = dropout'(dx) (dtype of dy);
 *                          dbias += colsum(da);  dw,db += LayerNorm parameter gradients.
 * Same masks as dropout_residual_fwd / dropout_bwd for the same (p, seed). */
int mmdti_dropres_layernorm_fwd(const float* res, const void* a, float* xo, const float* w,
                                const float* b, void* y, float* mean, float* rstd, int rows, int D,
                                float eps, float p, uint64_t seed, int a_dtype, int out_dtype,
                                void* stream);
int mmdti_layernorm_bwd_dropout(const void* dy, const float* x, const float* w, const float* mean,
                                const float* rstd, const float* dx_add, float* dx, float* dw,
                                float* db, void* da, float* dbias, int rows, int D, float p,
                                uint64_t seed, int dy_dtype, void* stream);
/* out = res + dropout(a): res,out (n) f32 (may alias), a (n) a_dtype.  n % 4 == 0. */
int mmdti_dropout_residual_fwd(const float* res, const void* a, float* out, int64_t n, float p,
                               uint64_t seed, int a_dtype, void* stream);
/* da = dropout'(dx) in da_dtype (rows,C); dbias (C) f32 += column sums of da (NULL to skip). */
int mmdti_dropout_bwd(const float* dx, void* da, float* dbias, int rows, int C, float p,
                      uint64_t seed, int da_dtype, void* stream);
/* dst (n) dst_dtype = (float)src (n) src_dtype + (add ? add (n) f32 : 0): dtype conversions and residual adds between fused
 * GEMMs (the post-LN layers of mm_module.py:525-536,578-589 keep their residual in fp32).  n % 8 == 0, f32 | bf16. */
int mmdti_convert_add(const void* src, int src_dtype, const float* add, void* dst, int dst_dtype, int64_t n, void* stream);
/* exact-erf GELU (unicore.utils.get_activation_fn("gelu") = F.gelu) and its backward:
 * dz = du * gelu'(z); dbias (C) f32 += column sums of dz (NULL to skip). */
int mmdti_gelu_fwd(const void* z, void* u, int64_t n, int dtype, void* stream);   /* n % 8 == 0 */
int mmdti_gelu_bwd(const void* du, const void* z, void* dz, float* dbias, int rows, int C, int dtype,
                   void* stream);
/* out (C) f32 += column sums of x (rows,C) — bias gradients. */
int mmdti_colsum(const void* x, float* out, int rows, int C, int dtype, void* stream);
/* Debug/test export: keep mask (uint8, n) of the flat-tensor dropout for `seed`. */
int mmdti_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, void* stream);

/* ---------------------------------------------------------------- K3 / K4: contrastive similarity
 * One two-phase engine serves InfoNCE (models/infonce.py:70-98), ConR = CT_Regress
 * (models/contrastive.py:3-59), SupCon-style CT_Single (:62-112) and CT_Multi (:114-169).  The
 * N x N similarity matrix is never written:
 *   phase 1 (sim_stats): s_ij = <a_i, b_j> / t for the M local anchor rows against all N keys
 *                        -> per-row statistics (M, 8) f32
 *   finalize:            statistics -> loss contribution and the per-row coefficients of phase 2
 *   phase 2 (sim_grad):  recompute s_ij, form the gradient coefficient H_ij,
 *                        dA_i = sum_j H_ij b_j   (M, D) f32
 * mode: MMDTI_SIM_INFONCE | _REGRESS | _SINGLE | _MULTI.  Anchor row r is GLOBAL sample
 * row_offset + r (data parallel: a rank owns rows [row_offset, row_offset+M) of the all-gathered
 * batch); label arrays are indexed by global sample.
 *   y, yhat   (N) f32: mean(depth,1), mean(output,1)      [REGRESS]   w_thr = w, e_push = e
 *   key       (N, C) int64 labels                           [SINGLE: C = 1; MULTI: C classes, coef]
 *   wrow/wcol (N) f32 or NULL: pushing weight w_ij = wrow[i] * wcol[j] (REGRESS: * |y_i-y_j| * e)
 * Masks are evaluated exactly as the reference does (fp32 subtract/abs/compare, integer equality):
 * bit-exact; mmdti_ct_masks exports them for the tests.
 * Statistics layout (row stride 8): InfoNCE {sum_j exp(z_ij - 1/t), z_ii}; CT {sum_P e^s, sum_N w e^s,
 * |P|, |N|, sum_P s}.
 *
 * *_f32: plain fp32 FMA (validation mode, 1e-5 parity), D <= 512.
 * *_tc : bf16 operands on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
 *        staged by TMA), fp32 statistics; A,B are the zero-padded bf16 copies written by
 *        mmdti_rownorm_fwd with row stride Dp (a multiple of 64, <= 512). */
#define MMDTI_SIM_INFONCE 0
#define MMDTI_SIM_REGRESS 1
#define MMDTI_SIM_SINGLE 2
#define MMDTI_SIM_MULTI 3
#define MMDTI_SIM_NSTAT 8

/* F.normalize(dim=-1, eps) (models/infonce.py:104-105, models/contrastive.py:21-22): x (N, D) f32 with
 * row stride ldx -> any of: xhat_f32 (N, D), xhat_bf16 (N, Dp) zero-padded columns [D, Dp),
 * inv_norm (N) = 1 / max(||x||, eps). */
int mmdti_rownorm_fwd(const float* x, int64_t ldx, float* xhat_f32, void* xhat_bf16, int Dp,
                      float* inv_norm, int N, int D, float eps, void* stream);
/* dx = coef * (g - xhat <xhat, g>) * inv_norm, coef = scale * (gscale ? *gscale : 1); gscale is a
 * device scalar (the upstream loss gradient).  accumulate != 0: dx += . */
int mmdti_rownorm_bwd(const float* g, const float* xhat, const float* inv_norm, float* dx,
                      int64_t lddx, int N, int D, float scale, const float* gscale, float eps,
                      int accumulate, void* stream);
int mmdti_sim_stats_f32(const float* A, const float* B, int M, int N, int D, int row_offset, int mode,
                        float temperature, const float* y, const float* yhat, float w_thr,
                        float e_push, const int64_t* key, int C, float coef_multi, const float* wrow,
                        const float* wcol, float* stats, void* stream);
int mmdti_sim_grad_f32(const float* A, const float* B, int M, int N, int D, int row_offset, int mode,
                       float temperature, const float* y, const float* yhat, float w_thr,
                       float e_push, const int64_t* key, int C, float coef_multi, const float* wrow,
                       const float* wcol, const float* rs_row, const float* rs_col, float* dA,
                       void* stream);
int mmdti_sim_stats_tc(const void* A, const void* B, int M, int N, int Dp, int row_offset, int mode,
                       float temperature, const float* y, const float* yhat, float w_thr,
                       float e_push, const int64_t* key, int C, float coef_multi, const float* wrow,
                       const float* wcol, float* stats, void* stream);
int mmdti_sim_grad_tc(const void* A, const void* B, int M, int N, int Dp, int row_offset, int mode,
                      float temperature, const float* y, const float* yhat, float w_thr, float e_push,
                      const int64_t* key, int C, float coef_multi, const float* wrow,
                      const float* wcol, const float* rs_row, const float* rs_col, float* dA,
                      int64_t lddA, void* stream);

/* Gradient coefficients alone (tensor cores): H (M, ldh) bf16, H_ij as in mmdti_sim_grad_*; columns [N, ldh) of a row are
 * written as zeros or left untouched.  With dA = H . B (a plain GEMM) this is the two-step form of phase 2 that large
 * N x 512-d problems use: the 512-column TMEM cannot hold the dA accumulators (512 columns at D = 512) next to a
 * similarity tile, so the fused kernel has to evaluate every H_ij once per 128/256-column slice of dA.
 * Same argument meaning as mmdti_sim_grad_tc; ldh >= N, ldh % 8 == 0.  Replaces the autograd backward of
 * models/infonce.py:91-98 and models/contrastive.py:55-60,110-115 up to the final H . B product. */
int mmdti_sim_coef_tc(const void* A, const void* B, int M, int N, int Dp, int row_offset, int mode, float temperature,
                      const float* y, const float* yhat, float w_thr, float e_push, const int64_t* key, int C,
                      float coef_multi, const float* wrow, const float* wcol, const float* rs_row, const float* rs_col,
                      void* H, int64_t ldh, void* stream);
/* InfoNCE: stats (M,8) -> lse (M) = log sum_j exp(z_ij); *loss += scale * sum_i (lse_i - z_ii). */
int mmdti_infonce_finalize(const float* stats, float* lse, float* loss, int M, float temperature,
                           float scale, void* stream);
/* CT: stats (M,8) -> rowstat (M,2) {c_i, alpha_i}; *loss += sum_i loss_i / N
 * (Z_i = sum_P e^s + (N - |P|) + sum_N w e^s: the exp(0)=1 quirk of contrastive.py:53). */
int mmdti_ct_finalize(const float* stats, float* rowstat, float* loss, int M, int N, int mode,
                      float w_thr, void* stream);
/* Debug/test export of the boolean masks (N,N) uint8 for mode REGRESS | SINGLE | MULTI. */
int mmdti_ct_masks(int mode, int N, const float* y, const float* yhat, float w_thr,
                   const int64_t* key, int C, float coef_multi, uint8_t* pos, uint8_t* neg,
                   void* stream);

/* ---------------------------------------------------------------- K5: FDS
 * models/fds.py + utils/util.py:159-169 without leaving the device.
 * mmdti_fds_bin: bins[i] = int((label_i - min_value) // bin_width) in fp32 exactly like
 *   models/fds.py:125,164 (torch floor_divide); present (bucket_num - bucket_start) int32 flags the
 *   in-range bins that occur in the batch (the reference loops over torch.unique(bins): the edge
 *   buckets absorb the tails only when the edge bin itself occurs). labels: element i at labels[i*ld]. */
int mmdti_fds_bin(const float* labels, int64_t ld, int N, float min_value, float bin_width,
                  int bucket_start, int bucket_num, int32_t* bins, int32_t* present, void* stream);
/* FDS.smooth (models/fds.py:157-190): x (N,D) f32 row stride ldx, IN PLACE:
 * x = (x - m1[b]) * sqrt(clamp(v2[b]/v1[b], .1, 10)) + m2[b] per calibrate_mean_var's three branches.
 * m1,v1 = running_{mean,var}_last_epoch, m2,v2 = smoothed_{mean,var}_last_epoch (nb, D).
 * work: optional caller-provided scratch of (3*nb*D + nb) floats, 16-byte aligned (nb = bucket_num - bucket_start): with it
 * (and D, ldx multiples of 4) the per-bucket factors are tabulated once and the per-sample pass is vectorised; NULL selects
 * the scalar kernel.  Same results either way. */
int mmdti_fds_smooth_fwd(float* x, int64_t ldx, const int32_t* bins, const int32_t* present, int N,
                         int D, int bucket_start, int bucket_num, const float* m1, const float* v1,
                         const float* m2, const float* v2, float* work, void* stream);
int mmdti_fds_smooth_bwd(const float* dy, float* dx, const int32_t* bins, const int32_t* present,
                         int N, int D, int bucket_start, int bucket_num, const float* v1,
                         const float* v2, float* work, void* stream);
/* FDS.update_running_stats (models/fds.py:116-155) in four steps so that data-parallel ranks can
 * all-reduce {count, sum1} and {m2} in between:
 *   group:       rows grouped by bucket -> seg (nb+1) offsets, order (N) row indices, count (nb) f32
 *   bucket_sums: sum1 (nb,D) = per-bucket column sums
 *   bucket_m2:   m2 (nb,D) = per-bucket sum of (x - sum1/count)^2   (count = global count)
 *   ema:         mean = sum1/count, var = m2/(count-1) (count==1: m2/1); tracked += count;
 *                running = (1-f) cur + f running, f = first_update ? 0 : (momentum >= 0 ? momentum
 *                : 1 - count/tracked)          -- only buckets with count > 0 are touched. */
int mmdti_fds_group(const int32_t* bins, const int32_t* present, int N, int bucket_start,
                    int bucket_num, int32_t* seg, int32_t* order, float* count, void* stream);
int mmdti_fds_bucket_sums(const float* x, int64_t ldx, const int32_t* seg, const int32_t* order,
                          float* sum1, int N, int D, int nb, void* stream);
int mmdti_fds_bucket_m2(const float* x, int64_t ldx, const int32_t* seg, const int32_t* order,
                        const float* sum1, const float* count, float* m2, int N, int D, int nb,
                        void* stream);
int mmdti_fds_ema(const float* count, const float* sum1, const float* m2, float* running_mean,
                  float* running_var, float* num_samples_tracked, int nb, int D, float momentum,
                  int first_update, void* stream);
/* FDS._update_last_epoch_stats (models/fds.py:86-99): out[b] = sum_k window[k] in[reflect(b+k-half)]
 * along the bucket axis (F.pad reflect + conv1d); in/out (nb, D) f32, in != out. */
int mmdti_fds_window(const float* in, const float* window, float* out, int nb, int D, int ks,
                     void* stream);

/* ---------------------------------------------------------------- fused multi-tensor Adam
 * torch.optim.Adam as the reference configures it (tasks/trainer.py:160-162: Adam(lr, eps=1e-6); no weight decay,
 * no amsgrad) in ONE launch over every parameter tensor:
 *   table  (ntensors, 6) int64 on the device: {p, g, m, v, lowp, numel}; p,g,m,v f32 device pointers, lowp a bf16
 *          device pointer or 0 — when given, the bf16 copy of the updated weight is written in the same pass;
 *   chunks (nchunks, 2) int32 on the device: {tensor index, first element}; one CTA per chunk of
 *          mmdti_adam_chunk() elements;
 *   step   device int64: the 1-based step count t (kept on the device so that CUDA-graph replays advance it);
 *   grad_scale multiplies every gradient first (1/world for summed data-parallel gradients). */
int mmdti_adam_chunk(void);
int mmdti_adam_step(const int64_t* table, const int32_t* chunks, int nchunks, const int64_t* step, double lr,
                    double beta1, double beta2, double eps, double grad_scale, void* stream);

/* ---------------------------------------------------------------- cross-modal fusion
 * BertCoAttention.forward (models/mm_module.py:493-522) as called by CrossAttentionModel.forward
 * (models/mm_model.py:386-406): queries of one modality attend to keys/values of the other.
 *   q (B*Lq, ldq), k / v (B*Lk, ldkv): head h occupies columns [h*head_dim, (h+1)*head_dim); act_dtype f32 | bf16
 *   key_mask (B, Lk) uint8, 1 = attend: the reference's additive mask (1 - mask) * -10000 on the scores
 *   o (B*Lq, ldo) = dropout(softmax(scale q k^T + mask)) v, act_dtype;  lse (B, H, Lq) f32 = opaque row statistics for
 *   the backward (log2 domain on the bf16 path, natural log on the f32 path).
 * bf16: head_dim 32 or 64, mma.sync flash-style kernels; f32: validation path.  Dropout keep decisions are a counter hash
 * of (seed, b*H+h, query, key): mmdti_cross_attn_dropout_mask exports them as (B, H, Lq, Lk) uint8. */
int mmdti_cross_attn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* key_mask,
                         void* o, int64_t ldo, float* lse, int B, int H, int Lq, int Lk, int head_dim, float scale,
                         float dropout_p, uint64_t seed, int act_dtype, void* stream);
/* Backward of the above: delta (B, H, Lq) f32 scratch (rowsum(dO * O), written here); dq (B*Lq, lddq), dk / dv
 * (B*Lk, lddkv) act_dtype, fully overwritten.  Deterministic (no atomics). */
int mmdti_cross_attn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* key_mask,
                         const void* o, const void* d_o, int64_t ldo, const float* lse, float* delta, void* dq,
                         int64_t lddq, void* dk, void* dv, int64_t lddkv, int B, int H, int Lq, int Lk, int head_dim,
                         float scale, float dropout_p, uint64_t seed, int act_dtype, void* stream);
int mmdti_cross_attn_dropout_mask(uint8_t* keep, int B, int H, int Lq, int Lk, float dropout_p, uint64_t seed,
                                  void* stream);
/* Masked mean pooling over the concatenation of the two fused sequences (models/mm_model.py:572-576):
 * out (B, D) f32 = (sum of the rows of x1 (B, L1, D) with mask1 + sum of the rows of x2 (B, L2, D) with mask2) /
 * (count1 + count2); x_dtype f32 | bf16; inv_count (B) f32 = 1 / (count1 + count2), kept for the backward; D % 4 == 0.
 * Backward: dx1 / dx2 f32, dense (zeros on masked rows). */
int mmdti_masked_pool_fwd(const void* x1, const uint8_t* mask1, int L1, const void* x2, const uint8_t* mask2, int L2,
                          float* out, float* inv_count, int B, int D, int x_dtype, void* stream);
int mmdti_masked_pool_bwd(const float* dout, const float* inv_count, const uint8_t* mask1, int L1, const uint8_t* mask2,
                          int L2, float* dx1, float* dx2, int B, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDTI_B200_H */
