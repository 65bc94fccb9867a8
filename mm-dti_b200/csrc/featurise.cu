// On-device pair featurisation: src_distance and src_edge_type of a padded batch from the
// tokens and the centred coordinates, so that a step uploads B*L*20 bytes instead of B*L*L*12.
//
// Replaces data/conformer.py:205-212,216-218 (scipy distance_matrix: float64 (sum |d|^2)^(1/2) of float32-valued
// coordinates, cast to float32; edge type = tok_i * len(dictionary) + tok_j) and the zero padding of
// utils/util.py:41-105.  Bit-exact by construction: the squares and the two additions are separate IEEE float64
// operations in numpy's order (no FMA contraction), the square root is correctly rounded, then one rounding to float32.
#include "common.cuh"

namespace {

__global__ void featurise_kernel(const float* __restrict__ coord, const long long* __restrict__ tok, int L, int n_dict, long long pad,
                                 float* __restrict__ dist, long long* __restrict__ et) {
    const int b = blockIdx.y, i = blockIdx.x;
    const long long ti = tok[(size_t)b * L + i];
    const float* ci = coord + ((size_t)b * L + i) * 3;
    const double xi = ci[0], yi = ci[1], zi = ci[2];
    const size_t row = ((size_t)b * L + i) * L;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const long long tj = tok[(size_t)b * L + j];
        const float* cj = coord + ((size_t)b * L + j) * 3;
        const double dx = xi - (double)cj[0], dy = yi - (double)cj[1], dz = zi - (double)cj[2];
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        const bool valid = ti != pad && tj != pad;
        dist[row + j] = valid ? (float)sqrt(s) : 0.f;
        et[row + j] = valid ? ti * n_dict + tj : 0ll;
    }
}

}  // namespace

extern "C" int mmdti_featurise(const float* coord, const int64_t* tokens, int B, int L, int n_dict, int64_t pad_idx, float* dist,
                               int64_t* edge_type, void* stream) {
    MMDTI_REQUIRE(coord && tokens && dist && edge_type && B > 0 && L > 0 && n_dict > 0, "featurise: bad arguments");
    MMDTI_REQUIRE(B <= 65535, "featurise: B must be <= 65535 (got %d)", B);
    const int threads = L >= 256 ? 256 : (L + 31) / 32 * 32;
    featurise_kernel<<<dim3(L, B), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        coord, reinterpret_cast<const long long*>(tokens), L, n_dict, (long long)pad_idx, dist, reinterpret_cast<long long*>(edge_type));
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
