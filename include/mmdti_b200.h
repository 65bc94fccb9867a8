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

/* Pieces of the backward of the above (the 128-wide GEMMs in between currently run as library
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
/* out = res + dropout(a): res,out (n) f32 (may alias), a (n) a_dtype.  n % 4 == 0. */
int mmdti_dropout_residual_fwd(const float* res, const void* a, float* out, int64_t n, float p,
                               uint64_t seed, int a_dtype, void* stream);
/* da = dropout'(dx) in da_dtype (rows,C); dbias (C) f32 += column sums of da (NULL to skip). */
int mmdti_dropout_bwd(const float* dx, void* da, float* dbias, int rows, int C, float p,
                      uint64_t seed, int da_dtype, void* stream);
/* exact-erf GELU (unicore.utils.get_activation_fn("gelu") = F.gelu) and its backward:
 * dz = du * gelu'(z); dbias (C) f32 += column sums of dz (NULL to skip). */
int mmdti_gelu_fwd(const void* z, void* u, int64_t n, int dtype, void* stream);
int mmdti_gelu_bwd(const void* du, const void* z, void* dz, float* dbias, int rows, int C, int dtype,
                   void* stream);
/* out (C) f32 += column sums of x (rows,C) — bias gradients. */
int mmdti_colsum(const void* x, float* out, int rows, int C, int dtype, void* stream);
/* Debug/test export: keep mask (uint8, n) of the flat-tensor dropout for `seed`. */
int mmdti_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDTI_B200_H */
