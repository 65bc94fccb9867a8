// Shared pieces of the contrastive-similarity kernels (K3 InfoNCE, K4 SupCon / ConR / multi-label):
// the per-pair functors that turn one similarity value s_ij into (a) row statistics and (b) the
// gradient coefficient H_ij.  Used unchanged by the fp32 validation kernels (contrastive.cu) and by
// the tcgen05 tensor-core kernels (sim_tc.cu), so both paths share one definition of every mask.
//
// Reference: models/infonce.py:70-98 (info_nce), models/contrastive.py:3-59 (CT_Regress = ConR),
// :62-112 (CT_Single = SupCon-style), :114-169 (CT_Multi); closed forms in SURVEY.md Appendix B.
#pragma once
#include "common.cuh"

#define SIM_INFONCE 0
#define SIM_REGRESS 1
#define SIM_SINGLE 2
#define SIM_MULTI 3

#define SIM_NSTAT 8   // floats of per-row statistics (row stride of the `stats` buffer)

// exp used by the functors: the including .cu may pre-define SIM_EXP (contrastive.cu: accurate expf
// for the fp32 validation path; sim_tc.cu: ex2.approx via __expf).
#ifndef SIM_EXP
#define SIM_EXP(x) __expf(x)
#endif

// Everything the functors need besides s_ij.  Label arrays are indexed by GLOBAL sample index
// (0..N-1); local anchor row r of this rank is global row row_offset + r.
struct SimAux {
    int mode;
    int N;               // number of keys (global batch)
    int row_offset;      // global index of local row 0
    float inv_t;         // 1 / temperature
    // ---- ConR (SIM_REGRESS): labels / predictions / threshold / pushing weight
    const float* y;      // (N) label = mean(depth, dim=1)            contrastive.py:9-11
    const float* yhat;   // (N) prediction = mean(output, dim=1)      contrastive.py:13-15
    float w_thr;         // w                                         contrastive.py:27-28
    float e_push;        // e                                         contrastive.py:41
    // ---- SupCon / multi: integer keys (N, C)
    const long long* key;
    int C;
    float thr_multi;     // (float)(coef / C)                         contrastive.py:136
    // ---- pushing-weight factors: w_ij = wrow[i] * wcol[j] (* l_dist * e for ConR); NULL = 1
    const float* wrow;
    const float* wcol;
    // ---- phase 2 inputs
    const float* rs_row; // InfoNCE: lse of the anchor rows (local, M).   CT: unused
    const float* rs_col; // InfoNCE: lse of the key "rows" (global, N).   CT: (N,2) {c_j, alpha_j}
};

struct PairMask {
    bool pos, neg;
    float l;             // l_dist (ConR) — feeds the pushing weight
};

// positive / negative masks of one (anchor i, key j) pair, global indices.  Bit-exact restatement of
// contrastive.py:17-31 (regress), :74-86 (single), :115-141 (multi): fp32 subtract, fabs, compare.
__device__ __forceinline__ PairMask sim_pair_mask(const SimAux& a, int i, int j) {
    PairMask m;
    m.l = 0.f;
    if (a.mode == SIM_REGRESS) {
        const float l = fabsf(__fsub_rn(a.y[i], a.y[j]));
        const float p = fabsf(__fsub_rn(a.yhat[i], a.yhat[j]));
        const bool close = l <= a.w_thr;
        m.pos = close && (i != j);
        m.neg = (!close) && (p <= a.w_thr);
        m.l = l;
    } else if (a.mode == SIM_SINGLE) {
        const bool same = a.key[i] == a.key[j];
        m.pos = same && (i != j);
        m.neg = !same;
    } else {
        int cnt = 0;
        for (int c = 0; c < a.C; ++c) cnt += (a.key[(long long)i * a.C + c] == a.key[(long long)j * a.C + c]) ? 1 : 0;
        const bool ge = __fdiv_rn((float)cnt, (float)a.C) >= a.thr_multi;
        m.pos = ge && (i != j);
        m.neg = !ge;
    }
    return m;
}

__device__ __forceinline__ float sim_push_w(const SimAux& a, const PairMask& m, int i, int j) {
    float w = 1.f;
    if (a.wrow) w *= a.wrow[i];
    if (a.wcol) w *= a.wcol[j];
    if (a.mode == SIM_REGRESS) w = m.l * w * a.e_push;
    return w;
}

// ---------------------------------------------------------------- phase 1: row statistics
// InfoNCE: st[0] += exp(z - 1/t)  (|z| <= 1/t, so no running max is needed), st[1] = z_ii.
// CT:      st[0] += [pos] e^s, st[1] += [neg] w e^s, st[2] += [pos], st[3] += [neg], st[4] += [pos] s.
struct RowAcc {
    float v[5];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = 0.f;
    }
};

__device__ __forceinline__ void sim_stats_accum(const SimAux& a, RowAcc& acc, int i, int j, float dot) {
    const float s = dot * a.inv_t;
    if (a.mode == SIM_INFONCE) {
        acc.v[0] += SIM_EXP(s - a.inv_t);
        if (i == j) acc.v[1] += s;
    } else {
        const PairMask m = sim_pair_mask(a, i, j);
        if (m.pos) {
            acc.v[0] += SIM_EXP(s);
            acc.v[2] += 1.f;
            acc.v[4] += s;
        } else if (m.neg) {
            acc.v[1] += sim_push_w(a, m, i, j) * SIM_EXP(s);
            acc.v[3] += 1.f;
        }
    }
}

// ---------------------------------------------------------------- phase 2: gradient coefficient
// InfoNCE: H_ij = softmax_row_i(z)_ij + softmax_col_j(z)_ij - 2 [i==j]       (x 1/(2N t) outside)
// CT:      H_ij = G_ij + G_ji,  G_ij = [pos](alpha_i e^s - c_i) + [neg] alpha_i w_ij e^s   (x 1/t outside)
// ri0/ri1: the anchor row's own statistics (InfoNCE: lse_i, unused; CT: c_i, alpha_i).
__device__ __forceinline__ float sim_grad_coeff(const SimAux& a, int i, int j, float dot, float ri0, float ri1) {
    const float s = dot * a.inv_t;
    if (a.mode == SIM_INFONCE) {
        float h = SIM_EXP(s - ri0) + SIM_EXP(s - a.rs_col[j]);
        if (i == j) h -= 2.f;
        return h;
    }
    const PairMask m = sim_pair_mask(a, i, j);
    if (!(m.pos || m.neg)) return 0.f;
    const float e = SIM_EXP(s);
    const float cj = a.rs_col[2 * j], aj = a.rs_col[2 * j + 1];
    if (m.pos) return (ri1 * e - ri0) + (aj * e - cj);
    const float wij = sim_push_w(a, m, i, j);
    PairMask mt = m;                       // l_dist, masks are symmetric in (i,j)
    const float wji = sim_push_w(a, mt, j, i);
    return (ri1 * wij + aj * wji) * e;
}
