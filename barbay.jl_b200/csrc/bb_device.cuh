// Device-side primitives: Philox4x32-10 noise lattice, Box-Muller, scalar math.
// sm_100a only.  See DESIGN.md "Noise lattice" for the counter layout; the oracle
// restates it independently in oracle/philox_ref.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bb {

constexpr int BLOCK = 128;         // threads per CTA of the column kernels (one column per thread)
constexpr int MAX_NT_DYN = 32;     // time points supported by the runtime-T fallback kernels
constexpr int MAX_NE_DYN = 8;      // environments supported by the runtime-E fallback kernels
constexpr int MAX_SEG = 32;        // (replicate, population) segments per launch
constexpr int MAX_K_SHARED = 1024;  // upper bound on K handled by the shared-latent kernel's smem staging

enum : uint32_t { STREAM_COLUMN = 0, STREAM_SHARED = 1, STREAM_HYPER = 2, STREAM_INIT = 3 };

template <typename real> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };
template <typename real> using vec2 = typename V2<real>::type;

template <typename real> __device__ __forceinline__ vec2<real> mk2(real a, real b) {
    vec2<real> v; v.x = a; v.y = b; return v;
}

// ---------------------------------------------------------------- Philox4x32-10
// The ten round keys (k + r * W) are kernel-uniform: they are computed once on the host
// and sit in the kernel's constant bank, so a round is 2 IMAD.WIDE + 2 LOP3.
struct PhiloxKey {
    uint32_t k0[10], k1[10];
};
inline PhiloxKey make_philox_key(uint64_t seed) {
    PhiloxKey k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { k.k0[r] = a; k.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return k;
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKey &key, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r], n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// MUFU-level approximations (no denormal / IEEE slow paths): operands here are never denormal
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// uniform in (0,1) from the top 23 bits: ((x >> 9) + 0.5) / 2^23
__device__ __forceinline__ float u23(uint32_t x, float) {
    return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994039535522f;   // f - (1 - 2^-24), exact
}
__device__ __forceinline__ double u23(uint32_t x, double) {
    return (static_cast<double>(x >> 9) + 0.5) * (1.0 / 8388608.0);
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float &n0, float &n1) {
    const float u = u23(xa, 0.f), v = u23(xb, 0.f);
    const float radius = fast_sqrt(-1.3862943611198906f * fast_lg2(u));   // sqrt(-2 ln u)
    float s, c;
    __sincosf(6.2831853071795865f * v, &s, &c);
    n0 = radius * c; n1 = radius * s;
}
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, double &n0, double &n1) {
    const double u = u23(xa, 0.0), v = u23(xb, 0.0);
    const double radius = sqrt(-2.0 * log(u));
    double s, c;
    sincospi(2.0 * v, &s, &c);
    n0 = radius * c; n1 = radius * s;
}

template <typename real>
__device__ __forceinline__ void normals4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                         const PhiloxKey &key, real (&n)[4]) {
    uint32_t x[4];
    philox4x32_10(c0, c1, c2, c3, key, x);
    box_muller(x[0], x[1], n[0], n[1]);
    box_muller(x[2], x[3], n[2], n[3]);
}

// normal attached to slot `i` of a non-column stream (shared / hyper / init)
template <typename real>
__device__ __forceinline__ real stream_normal(uint32_t stream, uint32_t i, uint32_t k, uint32_t step,
                                              const PhiloxKey &key) {
    real n[4];
    normals4<real>(i >> 2, stream << 24, k, step, key, n);
    const uint32_t lane = i & 3u;
    return lane == 0 ? n[0] : lane == 1 ? n[1] : lane == 2 ? n[2] : n[3];
}

// ---------------------------------------------------------------- scalar math
__device__ __forceinline__ float bb_exp(float x) { return fast_ex2(x * 1.4426950408889634f); }
__device__ __forceinline__ double bb_exp(double x) { return exp(x); }
__device__ __forceinline__ float bb_log(float x) { return logf(x); }
__device__ __forceinline__ double bb_log(double x) { return log(x); }
__device__ __forceinline__ float bb_sqrt(float x) { return fast_sqrt(x); }
__device__ __forceinline__ float bb_rcp(float x) { return fast_rcp(x); }
__device__ __forceinline__ double bb_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ double bb_sqrt(double x) { return sqrt(x); }

// softplus(w) = log(1 + e^w), stable; StatsFuns.softplus
__device__ __forceinline__ float softplus(float w) { return fmaxf(w, 0.f) + log1pf(expf(-fabsf(w))); }
__device__ __forceinline__ double softplus(double w) { return fmax(w, 0.0) + log1p(exp(-fabs(w))); }
// d softplus / dw
__device__ __forceinline__ float sigmoidf_(float w) { return 1.f / (1.f + expf(-w)); }
__device__ __forceinline__ float sigmoid(float w) { return sigmoidf_(w); }
__device__ __forceinline__ double sigmoid(double w) { return 1.0 / (1.0 + exp(-w)); }

template <typename real> __device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace bb
