// Device-side primitives: Philox4x32-7 noise lattice, Box-Muller, scalar math.
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

// ---------------------------------------------------------------- Philox4x32-7
// Philox4x32 (Salmon et al., SC'11) with 7 rounds: the round count the Random123 authors report as the smallest
// that is Crush-resistant (BigCrush) for the 4x32 variant; 10 is their safety-margin default.  Noise generation
// is the largest item of the K = 8 step (two passes regenerate it), and the Philox chain is its serial part.
// The round keys (k + r * W) are kernel-uniform: they are computed once on the host and sit in the kernel's
// constant bank, so a round is 2 IMAD.WIDE + 2 LOP3.
constexpr int PHILOX_ROUNDS = 7;
struct PhiloxKey {
    uint32_t k0[PHILOX_ROUNDS], k1[PHILOX_ROUNDS];
    const float2 *trig;     // fp32 Box-Muller direction table in global memory (TRIG_N entries), see box_muller
};
inline PhiloxKey make_philox_key(uint64_t seed) {
    PhiloxKey k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < PHILOX_ROUNDS; ++r) { k.k0[r] = a; k.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    k.trig = nullptr;
    return k;
}

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKey &key, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
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

// One 32-bit Philox word feeds one Box-Muller pair: the top 22 bits give the radius uniform
// u = (j + 0.5) / 2^22 (largest radius 5.65 sigma), the low 10 bits the direction (cos, sin)(2 pi (a + 0.5) / 1024).
// With 1024 equispaced directions every trigonometric moment up to order 1023 equals that of a continuous
// angle, so the pair is exactly uncorrelated with exact second and fourth moments; one Philox4x32-7
// call therefore yields eight standard normals.  Identical definition in fp32 and fp64.
// fp32: the directions come from a table instead of two MUFU ops (the step is dispatch-bound): TRIG_N = 1024
// entries sqrt(2 ln 2) (cos, sin), rounded once from double, indexed by the low bits of the word as they are -- no
// sign or quadrant fix-up (a half-turn table with a sign bit cost two more ALU instructions per word, 8 words per
// column and sample: 4 % of the K = 8 step).  `tab` is the table -- the column kernels pass their shared-memory copy,
// everything else the global one of the key.
constexpr int TRIG_N = 1024;
__device__ __forceinline__ void box_muller(uint32_t x, float &n0, float &n1, const float2 *tab) {
    // j = x >> 10 dropped into the mantissa of 2.0f by one funnel shift: 2 + j 2^-22, minus (2 - 2^-23) -> (j + 0.5) / 2^22, exact
    const float u = __uint_as_float(__funnelshift_r(x, 0x100u, 10)) - 1.99999988079071044921875f;
    const float radius = fast_sqrt(-fast_lg2(u));   // sqrt(-2 ln u) / sqrt(2 ln 2): the table carries the factor
    const float2 d = tab[x & (TRIG_N - 1)];
    n0 = radius * d.x; n1 = radius * d.y;
}
__device__ __forceinline__ void box_muller(uint32_t x, double &n0, double &n1, const float2 *) {
    const double u = (static_cast<double>(x >> 10) + 0.5) * (1.0 / 4194304.0);
    const double v = (static_cast<double>(x & 0x3ffu) + 0.5) * (1.0 / 1024.0);
    const double radius = sqrt(-2.0 * log(u));
    double s, c;
    sincospi(2.0 * v, &s, &c);
    n0 = radius * c; n1 = radius * s;
}

// eight normals per counter: word w -> lanes 2w (radius * cos) and 2w + 1 (radius * sin)
template <typename real>
__device__ __forceinline__ void normals8(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                         const PhiloxKey &key, const float2 *tab, real (&n)[8]) {
    uint32_t x[4];
    philox4x32(c0, c1, c2, c3, key, x);
    box_muller(x[0], n[0], n[1], tab);
    box_muller(x[1], n[2], n[3], tab);
    box_muller(x[2], n[4], n[5], tab);
    box_muller(x[3], n[6], n[7], tab);
}

// normal attached to slot `i` of a non-column stream (shared / hyper / init)
template <typename real>
__device__ __forceinline__ real stream_normal(uint32_t stream, uint32_t i, uint32_t k, uint32_t step,
                                              const PhiloxKey &key) {
    // lane i & 7 of normals8(i >> 3, ...): only its own word goes through Box-Muller
    uint32_t x[4];
    philox4x32(i >> 3, stream << 24, k, step, key, x);
    const uint32_t w = (i >> 1) & 3u;
    const uint32_t xw = w == 0 ? x[0] : w == 1 ? x[1] : w == 2 ? x[2] : x[3];
    real n0, n1;
    box_muller(xw, n0, n1, key.trig);
    return (i & 1u) ? n1 : n0;
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
