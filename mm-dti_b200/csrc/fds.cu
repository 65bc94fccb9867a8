// K5 — FDS (feature distribution smoothing) on the device: label binning, per-bucket running
// statistics with warp-shuffle segmented reductions, bucket-axis window smoothing, and the
// per-sample feature re-calibration (forward in place + backward).
//
// Reference: models/fds.py:116-155 (update_running_stats), :157-190 (smooth), :86-99
// (_update_last_epoch_stats), utils/util.py:159-169 (calibrate_mean_var).  The reference bins
// every sample on the HOST (`int((value - min)//bin_width)` per molecule = one device sync per
// sample) and then loops over torch.unique(bins) with boolean-index kernels; here a batch is
// binned, grouped and reduced without leaving the GPU.
//
// All kernels are HBM-bound: smooth reads+writes N*D*4 B once, the statistics read N*D*4 B twice
// (mean pass, centred second-moment pass — the two-pass form torch.var uses).
#include "common.cuh"

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

// torch.floor_divide on fp32 (c10::div_floor_floating) — what `(value - min) // bin_width`
// evaluates to for a 0-dim fp32 tensor `value` and host scalars (models/fds.py:125,164).
__device__ __forceinline__ float floor_div_f32(float a, float b) {
    if (b == 0.f) return __fdiv_rn(a, b);
    const float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.f && ((b < 0.f) != (mod < 0.f))) div = __fsub_rn(div, 1.f);
    float fl;
    if (div != 0.f) {
        fl = floorf(div);
        if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.f);
    } else {
        fl = copysignf(0.f, __fdiv_rn(a, b));
    }
    return fl;
}

// bucket (0-based, relative to bucket_start) a sample with bin `bin` is grouped into by the loops at
// models/fds.py:133-141 / :166-189, or -1 when no loop iteration touches it.  Interior bins map to
// themselves; the two edge buckets absorb the tails, but ONLY when the edge bin itself occurs in the
// batch (`present`), because the reference iterates over torch.unique(bins).
__device__ __forceinline__ int fds_bucket(int bin, const int* __restrict__ present, int bucket_start, int bucket_num) {
    const int last = bucket_num - 1;
    if (bucket_start >= last) {           // degenerate: a single bucket
        if (bucket_start == last && bin == last) return 0;
        return -1;
    }
    if (bin > bucket_start && bin < last) return bin - bucket_start;
    if (bin <= bucket_start) return present[0] ? 0 : -1;
    return present[last - bucket_start] ? last - bucket_start : -1;
}

// ------------------------------------------------------------------ binning
// bins[i] = int(floor_divide(label_i - min, width)); present[b] = 1 for every in-range bin that occurs.
__global__ void fds_bin_kernel(const float* __restrict__ labels, long long ld, int N, float min_value, float bin_width,
                               int bucket_start, int bucket_num, int* __restrict__ bins, int* __restrict__ present) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float f = floor_div_f32(__fsub_rn(labels[(long long)i * ld], min_value), bin_width);
        int b;
        if (!(f == f)) b = INT_MIN;                       // NaN label: never grouped
        else if (f >= 2147483520.f) b = INT_MAX;
        else if (f <= -2147483520.f) b = INT_MIN;
        else b = (int)f;
        bins[i] = b;
        if (b >= bucket_start && b <= bucket_num - 1) present[b - bucket_start] = 1;     // benign race: all write 1
    }
}

// ------------------------------------------------------------------ grouping (counting sort by bucket)
// one CTA: histogram of buckets -> exclusive scan -> seg (nb+1) ; then scatter row indices into `order`.
__global__ void __launch_bounds__(1024) fds_group_kernel(const int* __restrict__ bins, const int* __restrict__ present, int N,
                                                         int bucket_start, int bucket_num, int nb, int* __restrict__ seg,
                                                         int* __restrict__ order, float* __restrict__ count) {
    extern __shared__ int sh[];           // cnt[nb] | cursor[nb]
    int* cnt = sh;
    int* cur = sh + nb;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int b = fds_bucket(bins[i], present, bucket_start, bucket_num);
        if (b >= 0) atomicAdd(&cnt[b], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nb; ++b) {
            seg[b] = run;
            cur[b] = run;
            count[b] = (float)cnt[b];
            run += cnt[b];
        }
        seg[nb] = run;
    }
    __syncthreads();
    // stable within a bucket: rows are visited in ascending chunks of blockDim and ranked with a
    // per-chunk ballot-free scheme (one thread per bucket walks its chunk) would serialise; instead keep
    // ascending order by letting thread 0..nb-1 own bucket b and scan the rows (N is an epoch's sample
    // count, nb <= 1024: N*nb/1024 comparisons per thread).
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        if (cnt[b] == 0) continue;
        int pos = cur[b];
        for (int i = 0; i < N; ++i)
            if (fds_bucket(bins[i], present, bucket_start, bucket_num) == b) order[pos++] = i;
    }
}

// Parallel, STABLE counting sort (nb <= 256): every warp owns a contiguous row range, counts its rows per bucket,
// a scan over (bucket, warp) gives each warp its write cursor per bucket, and the warp scatters its rows in ascending
// order -- lanes of a 32-row step that share a bucket rank themselves with __match_any_sync.  Same `order` as the serial
// form above (ascending rows inside a bucket), ~200x faster at epoch sizes.
constexpr int GROUP_WARPS = 32;
__global__ void __launch_bounds__(GROUP_WARPS * 32) fds_group_par_kernel(const int* __restrict__ bins, const int* __restrict__ present,
                                                                         int N, int bucket_start, int bucket_num, int nb,
                                                                         int* __restrict__ seg, int* __restrict__ order,
                                                                         float* __restrict__ count) {
    extern __shared__ int sh[];           // wcnt[GROUP_WARPS][nb] | tot[nb] | base[nb]
    int* wcnt = sh;
    int* tot = sh + GROUP_WARPS * nb;
    int* base = tot + nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < GROUP_WARPS * nb; i += blockDim.x) wcnt[i] = 0;
    __syncthreads();
    const int chunk = ((N + GROUP_WARPS - 1) / GROUP_WARPS + 31) & ~31;      // rows per warp, multiple of 32
    const int r0 = warp * chunk, r1 = min(N, r0 + chunk);
    for (int i = r0 + lane; i < r1; i += 32) {
        const int b = fds_bucket(bins[i], present, bucket_start, bucket_num);
        if (b >= 0) atomicAdd(&wcnt[warp * nb + b], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        int run = 0;
        for (int w = 0; w < GROUP_WARPS; ++w) {
            const int c = wcnt[w * nb + b];
            wcnt[w * nb + b] = run;            // cursor of warp w inside bucket b
            run += c;
        }
        tot[b] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nb; ++b) {
            seg[b] = run;
            base[b] = run;
            count[b] = (float)tot[b];
            run += tot[b];
        }
        seg[nb] = run;
    }
    __syncthreads();
    for (int i0 = r0; i0 < r1; i0 += 32) {
        const int i = i0 + lane;
        const int b = i < r1 ? fds_bucket(bins[i], present, bucket_start, bucket_num) : -1;
        const unsigned same = __match_any_sync(0xffffffffu, b);
        const int rank = __popc(same & ((1u << lane) - 1u));
        if (b >= 0) {
            const int cur = wcnt[warp * nb + b];
            order[base[b] + cur + rank] = i;
        }
        __syncwarp();
        if (b >= 0 && rank == 0) wcnt[warp * nb + b] += __popc(same);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ segmented sums over row ranges (D == 512 panels, D % 4 == 0)
// CTA = RPC consecutive entries of `order` (rows sorted by bucket, so a range spans one or two buckets); every warp sums its
// rows of the range into register accumulators (NV float4 per lane: one coalesced 2 KB row per warp load), warps of the CTA
// are folded through shared memory whenever the bucket changes, and the CTA adds its partial to the bucket's row with
// 16-byte reductions.  Work per CTA is uniform whatever the bucket populations are (the per-bucket grid above serialises
// on the fullest bucket).
constexpr int SEG_RPC = 128;
template <int PASS>
__global__ void __launch_bounds__(256) fds_segsum_range_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ seg,
                                                               const int* __restrict__ order, const float* __restrict__ sum1,
                                                               const float* __restrict__ count, float* __restrict__ out, int D, int nb) {
    extern __shared__ float red[];        // [8 warps][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = seg[nb];
    const int k0 = blockIdx.x * SEG_RPC, k1 = min(total, k0 + SEG_RPC);
    if (k0 >= k1) return;
    // bucket of the first entry of the range (seg is ascending: small linear search, nb <= 1024)
    int b = 0;
    while (seg[b + 1] <= k0) ++b;
    int k = k0;
    while (k < k1) {
        const int kend = min(k1, seg[b + 1]);             // entries [k, kend) belong to bucket b
        if (kend > k) {
            const float invn = PASS == 2 ? 1.f / count[b] : 0.f;
            for (int c0 = 0; c0 < D; c0 += 512) {
                float4 acc[4], mu[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int c = c0 + (v * 32 + lane) * 4;
                    mu[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (PASS == 2 && c < D) {
                        const float4 s4 = *reinterpret_cast<const float4*>(sum1 + (long long)b * D + c);
                        mu[v] = make_float4(s4.x * invn, s4.y * invn, s4.z * invn, s4.w * invn);
                    }
                }
                for (int kk = k + warp; kk < kend; kk += 8) {
                    const float* xr = x + (long long)order[kk] * ldx;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int c = c0 + (v * 32 + lane) * 4;
                        if (c < D) {
                            const float4 t = *reinterpret_cast<const float4*>(xr + c);
                            if (PASS == 1) { acc[v].x += t.x; acc[v].y += t.y; acc[v].z += t.z; acc[v].w += t.w; }
                            else {
                                const float dx_ = t.x - mu[v].x, dy_ = t.y - mu[v].y, dz_ = t.z - mu[v].z, dw_ = t.w - mu[v].w;
                                acc[v].x = fmaf(dx_, dx_, acc[v].x); acc[v].y = fmaf(dy_, dy_, acc[v].y);
                                acc[v].z = fmaf(dz_, dz_, acc[v].z); acc[v].w = fmaf(dw_, dw_, acc[v].w);
                            }
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) *reinterpret_cast<float4*>(red + warp * 512 + (v * 32 + lane) * 4) = acc[v];
                __syncthreads();
                for (int c = threadIdx.x * 4; c < 512 && c0 + c < D; c += blockDim.x * 4) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        const float4 u = *reinterpret_cast<const float4*>(red + w * 512 + c);
                        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                    }
                    red_add_v4(out + (long long)b * D + c0 + c, t.x, t.y, t.z, t.w);
                }
                __syncthreads();
            }
        }
        k = kend;
        ++b;
    }
}

// ------------------------------------------------------------------ segmented sums
// grid (nb, splits).  CTA (b, s) walks rows order[seg[b] + s :: splits]; each warp takes every
// (nwarps)-th of those rows and holds the row as NV float4 per lane (coalesced 512 B per warp load);
// warps combine through shared memory, CTAs through atomics (splits > 1).
//   PASS 1: out[b][c] += x[r][c]                    (then mean = sum / n)
//   PASS 2: out[b][c] += (x[r][c] - mean[b][c])^2   (centred second moment, the torch.var two-pass form)
template <int PASS>
__global__ void __launch_bounds__(256) fds_segsum_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ seg,
                                                         const int* __restrict__ order, const float* __restrict__ sum1,
                                                         const float* __restrict__ count, float* __restrict__ out, int D) {
    extern __shared__ float red[];        // [nwarps][D]
    const int b = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int r0 = seg[b], r1 = seg[b + 1];
    if (r1 == r0) return;
    const float invn = PASS == 2 ? 1.f / count[b] : 0.f;      // count = GLOBAL samples of the bucket (all ranks)
    for (int c0 = 0; c0 < D; c0 += 128 * 4) {            // column panel of 512
        float acc[4][4];
        float mu[4][4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[v][e] = 0.f;
                const int c = c0 + (v * 32 + lane) * 4 + e;
                mu[v][e] = (PASS == 2 && c < D) ? sum1[(long long)b * D + c] * invn : 0.f;
            }
        for (int k = r0 + split * nw + warp; k < r1; k += nsplit * nw) {
            const float* xr = x + (long long)order[k] * ldx;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int c = c0 + (v * 32 + lane) * 4;
                if (c + 3 < D && (ldx & 3) == 0) {
                    const float4 t = *reinterpret_cast<const float4*>(xr + c);
                    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float d = tv[e] - mu[v][e];
                        acc[v][e] += PASS == 1 ? tv[e] : d * d;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c + e < D) {
                            const float d = xr[c + e] - mu[v][e];
                            acc[v][e] += PASS == 1 ? xr[c + e] : d * d;
                        }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = (v * 32 + lane) * 4 + e;
                red[warp * 512 + c] = acc[v][e];
            }
        __syncthreads();
        for (int c = threadIdx.x; c < 512 && c0 + c < D; c += blockDim.x) {
            float s = 0.f;
            for (int w = 0; w < nw; ++w) s += red[w * 512 + c];
            float* dst = out + (long long)b * D + c0 + c;
            if (nsplit > 1) atomicAdd(dst, s);
            else *dst = s;
        }
        __syncthreads();
    }
}

// small-D variant (D <= 32): one warp per bucket, lanes = (row-in-group, column), the row groups are
// folded with warp shuffles.  Covers narrow feature dims where the float4 panel above wastes lanes.
template <int PASS>
__global__ void __launch_bounds__(128) fds_segsum_small_kernel(const float* __restrict__ x, long long ldx,
                                                               const int* __restrict__ seg, const int* __restrict__ order,
                                                               const float* __restrict__ sum1, const float* __restrict__ count,
                                                               float* __restrict__ out, int D, int DP, int nb) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nb) return;
    const int b = warp;
    const int r0 = seg[b], r1 = seg[b + 1];
    if (r1 == r0) return;
    const int rpw = 32 / DP;                       // rows per warp step (DP = D rounded up to a power of two)
    const int c = lane % DP, rr = lane / DP;
    const float mu = (PASS == 2 && c < D) ? sum1[(long long)b * D + c] / count[b] : 0.f;
    float acc = 0.f;
    for (int k = r0 + rr; k < r1; k += rpw) {
        if (c < D) {
            const float v = x[(long long)order[k] * ldx + c];
            const float d = v - mu;
            acc += PASS == 1 ? v : d * d;
        }
    }
    for (int o = 16; o >= DP; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (rr == 0 && c < D) out[(long long)b * D + c] = acc;
}

// ------------------------------------------------------------------ EMA update of the running statistics
// models/fds.py:142-153 for every bucket that received samples (n = seg[b+1]-seg[b] > 0):
//   mean = sum/n; var = m2/(n-1) (n == 1: biased, i.e. 0); tracked[b] += n;
//   factor = first ? 0 : (momentum >= 0 ? momentum : 1 - n/tracked[b]);  running = (1-f) cur + f running
__global__ void fds_ema_kernel(const float* __restrict__ sum1, const float* __restrict__ m2,
                               float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ tracked,
                               int nb, int D, float momentum, int first, const float* __restrict__ count) {
    const int b = blockIdx.x;
    const float n = count[b];
    if (n <= 0.f) return;
    __shared__ float s_factor;
    if (threadIdx.x == 0) {
        const float t = tracked[b] + n;
        tracked[b] = t;
        float f = momentum >= 0.f ? momentum : 1.f - n / t;
        if (first) f = 0.f;
        s_factor = f;
    }
    __syncthreads();
    const float f = s_factor;
    const float dn = n > 1.f ? n - 1.f : 1.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const long long k = (long long)b * D + c;
        const float mean = sum1[k] / n, var = m2[k] / dn;
        running_mean[k] = (1.f - f) * mean + f * running_mean[k];
        running_var[k] = (1.f - f) * var + f * running_var[k];
    }
}

// ------------------------------------------------------------------ bucket-axis window smoothing
// out[b][c] = sum_k win[k] * in[reflect(b + k - half)][c]   (F.pad mode='reflect' + conv1d, fds.py:90-99)
__global__ void fds_window_kernel(const float* __restrict__ in, const float* __restrict__ win, float* __restrict__ out, int nb,
                                  int D, int ks) {
    const int half = (ks - 1) / 2;
    const long long n = (long long)nb * D;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / D), c = (int)(idx - (long long)b * D);
        float s = 0.f;
        for (int k = 0; k < ks; ++k) {
            int j = b + k - half;
            if (j < 0) j = -j;
            if (j >= nb) j = 2 * (nb - 1) - j;
            s = fmaf(win[k], in[(long long)j * D + c], s);
        }
        out[idx] = s;
    }
}

// ------------------------------------------------------------------ smooth (calibrate) forward / backward
// One warp per sample.  With b = bucket of the sample (skip when none):
//   if sum_c v1[b][c] < 1e-10: unchanged                                   (util.py:160-161)
//   else for every column with v1 != 0:  x = (x - m1) * sqrt(clamp(v2/v1, 0.1, 10)) + m2   (:162-169)
// BWD = 1: dx = dy * sqrt(factor) on the transformed entries, dy elsewhere.
template <int BWD>
__global__ void __launch_bounds__(256) fds_smooth_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ dy,
                                                         const int* __restrict__ bins, const int* __restrict__ present, int N,
                                                         int D, int bucket_start, int bucket_num, const float* __restrict__ m1,
                                                         const float* __restrict__ v1, const float* __restrict__ m2,
                                                         const float* __restrict__ v2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + warp; row < N; row += gridDim.x * wpb) {
        const int b = fds_bucket(bins[row], present, bucket_start, bucket_num);
        bool active = b >= 0;
        const long long off = (long long)(active ? b : 0) * D;
        if (active) {
            float s = 0.f;
            for (int c = lane; c < D; c += 32) s += v1[off + c];
            s = warp_sum(s);
            active = !(s < 1e-10f);
        }
        float* xr = x + (long long)row * ldx;
        if (!active) {
            if (BWD) for (int c = lane; c < D; c += 32) xr[c] = dy[(long long)row * D + c];
            continue;
        }
        for (int c = lane; c < D; c += 32) {
            const float a = v1[off + c];
            if (BWD) {
                const float g = dy[(long long)row * D + c];
                xr[c] = a != 0.f ? g * sqrtf(fminf(fmaxf(__fdiv_rn(v2[off + c], a), 0.1f), 10.f)) : g;
            } else if (a != 0.f) {
                const float f = sqrtf(fminf(fmaxf(__fdiv_rn(v2[off + c], a), 0.1f), 10.f));
                xr[c] = __fadd_rn(__fmul_rn(__fsub_rn(xr[c], m1[off + c]), f), m2[off + c]);
            }
        }
    }
}

// ---- table form of the calibration (bandwidth-bound at epoch sizes): per bucket and column the factor
// sqrt(clamp(v2/v1, 0.1, 10)) and the two means are evaluated ONCE (nb x D values) instead of once per sample, with the
// identity (f = 1, m1 = 0, m2 = -0) where v1 == 0 and a per-bucket "active" flag (sum(v1) >= 1e-10); the per-sample kernel
// then is three vector loads and (x - m1) * f + m2 with the reference's three separate roundings.
// tab: [F | M1 | M2] (3 x nb x D) then active (nb) floats.
__global__ void __launch_bounds__(256) fds_table_kernel(const float* __restrict__ m1, const float* __restrict__ v1,
                                                        const float* __restrict__ m2, const float* __restrict__ v2,
                                                        float* __restrict__ tab, int nb, int D) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const long long off = (long long)b * D, plane = (long long)nb * D;
    float s = 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) s += v1[off + c];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        tab[3 * plane + b] = (t < 1e-10f) ? 0.f : 1.f;
    }
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const float a = v1[off + c];
        const bool on = a != 0.f;
        tab[off + c] = on ? sqrtf(fminf(fmaxf(__fdiv_rn(v2[off + c], a), 0.1f), 10.f)) : 1.f;
        tab[plane + off + c] = on && m1 ? m1[off + c] : 0.f;
        tab[2 * plane + off + c] = on && m2 ? m2[off + c] : -0.f;
    }
}

template <int BWD>
__global__ void __launch_bounds__(256) fds_smooth_vec_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ dy,
                                                             const int* __restrict__ bins, const int* __restrict__ present, int N,
                                                             int D, int bucket_start, int bucket_num, const float* __restrict__ tab) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const long long plane = (long long)(bucket_num - bucket_start) * D;
    for (int row = blockIdx.x * wpb + warp; row < N; row += gridDim.x * wpb) {
        const int b = fds_bucket(bins[row], present, bucket_start, bucket_num);
        const bool active = b >= 0 && tab[3 * plane + b] != 0.f;
        float* xr = x + (long long)row * ldx;
        if (!active) {
            if (BWD) for (int c = lane * 4; c < D; c += 128) *reinterpret_cast<float4*>(xr + c) = *reinterpret_cast<const float4*>(dy + (long long)row * D + c);
            continue;
        }
        const float* F = tab + (long long)b * D;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 f = *reinterpret_cast<const float4*>(F + c);
            if (BWD) {
                const float4 g = *reinterpret_cast<const float4*>(dy + (long long)row * D + c);
                *reinterpret_cast<float4*>(xr + c) = make_float4(g.x * f.x, g.y * f.y, g.z * f.z, g.w * f.w);
            } else {
                const float4 v = *reinterpret_cast<const float4*>(xr + c);
                const float4 a = *reinterpret_cast<const float4*>(F + plane + c), o = *reinterpret_cast<const float4*>(F + 2 * plane + c);
                *reinterpret_cast<float4*>(xr + c) = make_float4(__fadd_rn(__fmul_rn(__fsub_rn(v.x, a.x), f.x), o.x), __fadd_rn(__fmul_rn(__fsub_rn(v.y, a.y), f.y), o.y),
                                                                 __fadd_rn(__fmul_rn(__fsub_rn(v.z, a.z), f.z), o.z), __fadd_rn(__fmul_rn(__fsub_rn(v.w, a.w), f.w), o.w));
            }
        }
    }
}

}  // namespace

extern "C" int mmdti_fds_bin(const float* labels, int64_t ld, int N, float min_value, float bin_width, int bucket_start,
                             int bucket_num, int32_t* bins, int32_t* present, void* stream) {
    MMDTI_REQUIRE(labels && bins && present && N > 0 && ld >= 1, "fds_bin: bad arguments");
    MMDTI_REQUIRE(bucket_num > bucket_start && bucket_start >= 0, "fds_bin: need 0 <= bucket_start < bucket_num");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MMDTI_CUDA_OK(cudaMemsetAsync(present, 0, sizeof(int32_t) * (size_t)(bucket_num - bucket_start), st));
    fds_bin_kernel<<<std::min((N + 255) / 256, num_sms() * 4), 256, 0, st>>>(labels, ld, N, min_value, bin_width, bucket_start,
                                                                            bucket_num, bins, present);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_fds_smooth_fwd(float* x, int64_t ldx, const int32_t* bins, const int32_t* present, int N, int D,
                                    int bucket_start, int bucket_num, const float* m1, const float* v1, const float* m2,
                                    const float* v2, float* work, void* stream) {
    MMDTI_REQUIRE(x && bins && present && m1 && v1 && m2 && v2 && N > 0 && D > 0 && ldx >= D, "fds_smooth_fwd: bad arguments");
    if (work && D % 4 == 0 && ldx % 4 == 0 && mmdti_aligned(x, 16) && mmdti_aligned(work, 16)) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int nb = bucket_num - bucket_start;
        fds_table_kernel<<<nb, 256, 0, st>>>(m1, v1, m2, v2, work, nb, D);
        fds_smooth_vec_kernel<0><<<std::min((N + 7) / 8, num_sms() * 16), 256, 0, st>>>(x, ldx, nullptr, bins, present, N, D, bucket_start,
                                                                                       bucket_num, work);
        MMDTI_LAUNCH_OK();
        return MMDTI_OK;
    }
    fds_smooth_kernel<0><<<std::min((N + 7) / 8, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, nullptr, bins, present, N, D, bucket_start, bucket_num, m1, v1, m2, v2);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_fds_smooth_bwd(const float* dy, float* dx, const int32_t* bins, const int32_t* present, int N, int D,
                                    int bucket_start, int bucket_num, const float* v1, const float* v2, float* work, void* stream) {
    MMDTI_REQUIRE(dy && dx && bins && present && v1 && v2 && N > 0 && D > 0, "fds_smooth_bwd: bad arguments");
    if (work && D % 4 == 0 && mmdti_aligned(dy, 16) && mmdti_aligned(dx, 16) && mmdti_aligned(work, 16)) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int nb = bucket_num - bucket_start;
        fds_table_kernel<<<nb, 256, 0, st>>>(nullptr, v1, nullptr, v2, work, nb, D);
        fds_smooth_vec_kernel<1><<<std::min((N + 7) / 8, num_sms() * 16), 256, 0, st>>>(dx, D, dy, bins, present, N, D, bucket_start,
                                                                                       bucket_num, work);
        MMDTI_LAUNCH_OK();
        return MMDTI_OK;
    }
    fds_smooth_kernel<1><<<std::min((N + 7) / 8, num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dx, D, dy, bins, present, N, D, bucket_start, bucket_num, nullptr, v1, nullptr, v2);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

static int fds_segsum(int pass, const float* x, int64_t ldx, const int32_t* seg, const int32_t* order, const float* sum1,
                      const float* count, float* out,
                      int N, int D, int nb, cudaStream_t st) {
    if (D <= 32) {
        int DP = 1;
        while (DP < D) DP <<= 1;
        const int blocks = (nb * 32 + 127) / 128;
        if (pass == 1) fds_segsum_small_kernel<1><<<blocks, 128, 0, st>>>(x, ldx, seg, order, sum1, count, out, D, DP, nb);
        else fds_segsum_small_kernel<2><<<blocks, 128, 0, st>>>(x, ldx, seg, order, sum1, count, out, D, DP, nb);
    } else if (D % 4 == 0 && (ldx & 3) == 0 && mmdti_aligned(x, 16) && mmdti_aligned(out, 16) && (!sum1 || mmdti_aligned(sum1, 16))) {
        // uniform work per CTA: ranges of the bucket-sorted row order (out is zeroed by the caller, partials are added)
        const size_t smem = 8 * 512 * sizeof(float);
        const int grid = (N + SEG_RPC - 1) / SEG_RPC;
        if (pass == 1) fds_segsum_range_kernel<1><<<grid, 256, smem, st>>>(x, ldx, seg, order, sum1, count, out, D, nb);
        else fds_segsum_range_kernel<2><<<grid, 256, smem, st>>>(x, ldx, seg, order, sum1, count, out, D, nb);
    } else {
        int splits = std::max(1, std::min(64, (2 * num_sms() + nb - 1) / nb));
        splits = std::max(1, std::min(splits, (N / nb + 63) / 64));
        const size_t smem = 8 * 512 * sizeof(float);
        dim3 grid(nb, splits);
        if (pass == 1) fds_segsum_kernel<1><<<grid, 256, smem, st>>>(x, ldx, seg, order, sum1, count, out, D);
        else fds_segsum_kernel<2><<<grid, 256, smem, st>>>(x, ldx, seg, order, sum1, count, out, D);
    }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_fds_group(const int32_t* bins, const int32_t* present, int N, int bucket_start, int bucket_num, int32_t* seg,
                               int32_t* order, float* count, void* stream) {
    const int nb = bucket_num - bucket_start;
    MMDTI_REQUIRE(bins && present && seg && order && count && N > 0 && nb > 0 && nb <= 4096, "fds_group: bad arguments (nb <= 4096)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nb <= 256) {
        const size_t smem = (size_t)(GROUP_WARPS * nb + 2 * nb) * sizeof(int);
        MMDTI_CUDA_OK(cudaFuncSetAttribute(fds_group_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fds_group_par_kernel<<<1, GROUP_WARPS * 32, smem, st>>>(bins, present, N, bucket_start, bucket_num, nb, seg, order, count);
    } else {
        fds_group_kernel<<<1, 1024, 2 * nb * sizeof(int), st>>>(bins, present, N, bucket_start, bucket_num, nb, seg, order, count);
    }
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_fds_bucket_sums(const float* x, int64_t ldx, const int32_t* seg, const int32_t* order, float* sum1, int N, int D,
                                     int nb, void* stream) {
    MMDTI_REQUIRE(x && seg && order && sum1 && N > 0 && D > 0 && nb > 0 && ldx >= D, "fds_bucket_sums: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MMDTI_CUDA_OK(cudaMemsetAsync(sum1, 0, sizeof(float) * (size_t)nb * D, st));
    return fds_segsum(1, x, ldx, seg, order, nullptr, nullptr, sum1, N, D, nb, st);
}

extern "C" int mmdti_fds_bucket_m2(const float* x, int64_t ldx, const int32_t* seg, const int32_t* order, const float* sum1,
                                   const float* count, float* m2, int N, int D, int nb, void* stream) {
    MMDTI_REQUIRE(x && seg && order && sum1 && count && m2 && N > 0 && D > 0 && nb > 0 && ldx >= D, "fds_bucket_m2: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MMDTI_CUDA_OK(cudaMemsetAsync(m2, 0, sizeof(float) * (size_t)nb * D, st));
    return fds_segsum(2, x, ldx, seg, order, sum1, count, m2, N, D, nb, st);
}

extern "C" int mmdti_fds_ema(const float* count, const float* sum1, const float* m2, float* running_mean,
                             float* running_var, float* num_samples_tracked, int nb, int D, float momentum, int first_update,
                             void* stream) {
    MMDTI_REQUIRE(count && sum1 && m2 && running_mean && running_var && num_samples_tracked && nb > 0 && D > 0,
                  "fds_ema: bad arguments");
    fds_ema_kernel<<<nb, 128, 0, static_cast<cudaStream_t>(stream)>>>(sum1, m2, running_mean, running_var, num_samples_tracked, nb,
                                                                    D, momentum, first_update, count);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}

extern "C" int mmdti_fds_window(const float* in, const float* window, float* out, int nb, int D, int ks, void* stream) {
    MMDTI_REQUIRE(in && window && out && in != out && nb > 0 && D > 0 && ks >= 1 && (ks & 1), "fds_window: bad arguments");
    MMDTI_REQUIRE((ks - 1) / 2 < nb, "fds_window: reflect padding needs (ks-1)/2 < number of buckets");
    const long long n = (long long)nb * D;
    fds_window_kernel<<<(int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, window, out, nb, D, ks);
    MMDTI_LAUNCH_OK();
    return MMDTI_OK;
}
