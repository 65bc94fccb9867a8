// Shared device/host helpers for the mmdti_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/mmdti_b200.h"

// ---------------------------------------------------------------- error plumbing
void mmdti_set_error(const char* fmt, ...);
#define MMDTI_REQUIRE(cond, ...)              \
    do {                                      \
        if (!(cond)) {                        \
            mmdti_set_error(__VA_ARGS__);     \
            return MMDTI_ERR_ARG;             \
        }                                     \
    } while (0)
#define MMDTI_CUDA_OK(call)                                                          \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            mmdti_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                     \
            return MMDTI_ERR_CUDA;                                                   \
        }                                                                            \
    } while (0)
#define MMDTI_LAUNCH_OK()                                                              \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            mmdti_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                       \
            return MMDTI_ERR_CUDA;                                                     \
        }                                                                              \
    } while (0)

static inline bool mmdti_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Launch-configuration caches (cudaFuncSetAttribute / occupancy) are per-device facts: key them by the current device.
constexpr int MMDTI_MAX_DEVICES = 64;
static inline int mmdti_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < MMDTI_MAX_DEVICES ? dev : 0;
}

// Scheduler counters of the persistent kernels with dynamic work distribution (gemm_tc.cu, pair_attn.cu backward): a
// zero-initialised ring of {next item, CTAs done} pairs in device memory, one pair per launch.  A kernel re-arms its pair when
// its last CTA retires, so a CUDA-graph replay (which re-runs the same launch with the same pair) finds it zeroed; the ring is
// long enough that a pair is never shared by two launches in flight.  NULL on allocation failure.
int* mmdti_sched_slot();

// ---------------------------------------------------------------- type helpers
typedef __nv_bfloat16 bf16;

template <typename T> struct TypeOf;
template <> struct TypeOf<float> { static constexpr int id = MMDTI_F32; };
template <> struct TypeOf<bf16> { static constexpr int id = MMDTI_BF16; };
template <> struct TypeOf<__half> { static constexpr int id = MMDTI_F16; };

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

// one F2FP (cvt.rn.bf16x2.f32) / one shift + one mask: the cuda_bf16.hpp helpers cost 3 instructions per pair
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
}
// pack / unpack two 16-bit pair-tensor elements of type TP
template <typename TP> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<bf16>(float lo, float hi) { return pack_bf16(lo, hi); }
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) { return pack_f16(lo, hi); }
template <typename TP> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<bf16>(uint32_t u) { return unpack_bf16(u); }
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) { return unpack_f16(u); }

// ---------------------------------------------------------------- fast math (bf16 paths)
// exp2f / __expf / __fdividef carry range fix-ups (FSETP + predicated FMULs around the MUFU); the flush-to-zero
// approximations are a single MUFU each, accurate to 2 ulp — far below the bf16 rounding of what consumes them.
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
// exact-erf GELU and its derivative with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7); exp(-x^2/2) is shared
// between erf and the normal density.  value = x Phi(x), grad = Phi(x) + x phi(x).
__device__ __forceinline__ void gelu_fast_both(float x, float& val, float& grad) {
    const float e = fast_ex2(x * x * -0.72134752044448170368f);                      // exp(-x^2/2)
    const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.f));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float erfa = fmaf(-poly * t, e, 1.f);                                      // erf(|x| / sqrt 2)
    const float cdf = 0.5f + copysignf(0.5f * erfa, x);
    val = x * cdf;
    grad = fmaf(x * e, 0.3989422804014327f, cdf);
}
__device__ __forceinline__ float gelu_fast_val(float x) {
    float v, g;
    gelu_fast_both(x, v, g);
    return v;
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one 16-byte reduction instead of four scalar atomics (sm_90+): the L2 atomic units see a quarter of the ops
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------- tensor-core wrappers
// D(16x8,f32) += A(16x8,bf16,row) * B(8x8,bf16,col)
__device__ __forceinline__ void mma_bf16_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(b0));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                               uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                                  const void* smem_row_ptr) {
    uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row_ptr) {
    uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r0), "=r"(r1)
                 : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                            const void* smem_row_ptr) {
    uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(a));
}

// ---------------------------------------------------------------- counter-based dropout RNG
// lowbias32 avalanche hash; the keep decision of element (stream, row, col) depends only on
// (seed, stream, row, col), so forward, backward and the debug dump agree under any tiling.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t rng_stream_key(unsigned long long seed, uint32_t stream) {
    uint32_t lo = static_cast<uint32_t>(seed), hi = static_cast<uint32_t>(seed >> 32);
    return mix32(lo ^ mix32(stream * 0x9E3779B1U + hi + 0x85EBCA6BU));
}
// 64 random bits for the "quad" of `row`: the column pairs (2*q4, 2*q4+1) of the 8-column blocks kb = 2m
// (word .x) and kb = 2m+1 (word .y), m = col >> 4, q4 = (col & 7) >> 1.  Low 16 bits of a word belong to
// the even column, high 16 bits to the odd one.  One avalanche hash + one multiply-xorshift per 4 elements.
// rows < 2^21, cols < 2^11.
__device__ __forceinline__ uint2 rng_quad_bits(uint32_t key, uint32_t row, uint32_t col) {
    const uint32_t qidx = ((col >> 4) << 2) | ((col & 7u) >> 1);
    uint2 w;
    w.x = mix32(key ^ ((row << 10) | qidx));
    w.y = w.x * 0x9E3779B1U;
    w.y ^= w.y >> 15;
    return w;
}
// the 32-bit word holding the bits of element (row, col)
__device__ __forceinline__ uint32_t rng_pair_bits(uint32_t key, uint32_t row, uint32_t col) {
    const uint2 w = rng_quad_bits(key, row, col);
    return (col & 8u) ? w.y : w.x;
}
__device__ __forceinline__ bool rng_keep(uint32_t bits, uint32_t col, uint32_t thresh16) {
    uint32_t r = (col & 1u) ? (bits >> 16) : (bits & 0xffffu);
    return r >= thresh16;
}
// per-halfword keep mask (0xFFFF = keep) of a 32-bit word of random bits
__device__ __forceinline__ uint32_t rng_keep_mask2(uint32_t bits, uint32_t thresh2) { return __vcmpgeu2(bits, thresh2); }

// Device-resident RNG offset (CUDA-graph replays): when the host registered a device counter with
// mmdti_set_seed_offset(), every dropout kernel adds *counter to its seed, so a captured step draws fresh
// masks on every replay while forward / backward / debug dumps of ONE step still agree.
const unsigned long long* mmdti_seed_offset_ptr();
__device__ __forceinline__ unsigned long long rng_effective_seed(unsigned long long seed, const unsigned long long* off) {
    return off ? seed + (*off) * 0x9E3779B97F4A7C15ULL : seed;
}
__device__ __forceinline__ uint32_t rng_effective_key(uint32_t key, const unsigned long long* off) {
    return off ? mix32(key + (uint32_t)(*off) * 0x9E3779B1U + (uint32_t)((*off) >> 32)) : key;
}

// exact unsigned division by a small runtime constant: q = n / d for n*d < 2^32
struct FastDiv {
    uint32_t d, m;
    __host__ __device__ explicit FastDiv(uint32_t d_) : d(d_), m(d_ > 1 ? (0xFFFFFFFFu / d_ + 1u) : 0u) {}
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d > 1 ? __umulhi(n, m) : n; }
};

// ---------------------------------------------------------------- async copy / mbarrier / TMA bulk
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory"); }
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// pair-tensor row stride (elements): 8 * (smallest supported odd block count covering L)
__host__ __device__ inline int mmdti_pair_nkb(int L) {
    const int nkb = (L + 7) / 8;
    if (nkb <= 3) return 3;
    if (nkb <= 5) return 5;
    if (nkb <= 9) return 9;
    if (nkb <= 13) return 13;
    if (nkb <= 17) return 17;
    if (nkb <= 25) return 25;
    if (nkb <= 33) return 33;
    return -1;
}
