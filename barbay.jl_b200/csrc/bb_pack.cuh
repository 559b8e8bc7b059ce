// Packs of W Monte-Carlo samples processed by one instruction stream.
//
// The fused step kernel (bb_step_kernel.cuh) walks the K samples of a column W at a time.  W = 2 with
// real = float maps every arithmetic step onto the packed fp32 instructions of sm_100 (PTX
// add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): one issue slot does the work of two, which is
// what the issue-bound K = 8 step needs.  Per-column scalars (mu, sigma, priors, counts) enter as
// broadcast operands -- FFMA2 takes a scalar register for either half (`R.F32`), so no duplicate
// registers are spent on them.  W = 1 and real = double are the same code with plain scalars.
#pragma once
#include "bb_device.cuh"

namespace bb {

template <typename real, int W> struct Pack;

// ---------------------------------------------------------------- W = 1: a scalar
template <typename real> struct Pack<real, 1> {
    real v;
};
template <typename real> __device__ __forceinline__ Pack<real, 1> pbc1(real a) { return Pack<real, 1>{a}; }

// ---------------------------------------------------------------- W = 2, double: two scalars
template <> struct Pack<double, 2> {
    double a, b;
};

// ---------------------------------------------------------------- W = 2, float: one 64-bit register pair
template <> struct Pack<float, 2> {
    float2 v;
};

template <typename P> struct PackTraits;
template <typename real> struct PackTraits<Pack<real, 1>> { using scalar = real; static constexpr int W = 1; };
template <typename real> struct PackTraits<Pack<real, 2>> { using scalar = real; static constexpr int W = 2; };

// ---- construction / access
template <typename real, int W> __device__ __forceinline__ Pack<real, W> pk_bc(real a);
template <> __device__ __forceinline__ Pack<float, 1> pk_bc<float, 1>(float a) { return {a}; }
template <> __device__ __forceinline__ Pack<double, 1> pk_bc<double, 1>(double a) { return {a}; }
template <> __device__ __forceinline__ Pack<double, 2> pk_bc<double, 2>(double a) { return {a, a}; }
template <> __device__ __forceinline__ Pack<float, 2> pk_bc<float, 2>(float a) { return {make_float2(a, a)}; }
__device__ __forceinline__ Pack<float, 2> pk_make(float a, float b) { return {make_float2(a, b)}; }
__device__ __forceinline__ Pack<double, 2> pk_make(double a, double b) { return {a, b}; }

template <int I> __device__ __forceinline__ float pk_get(const Pack<float, 1> &p) { return p.v; }
template <int I> __device__ __forceinline__ double pk_get(const Pack<double, 1> &p) { return p.v; }
template <int I> __device__ __forceinline__ double pk_get(const Pack<double, 2> &p) { return I == 0 ? p.a : p.b; }
template <int I> __device__ __forceinline__ float pk_get(const Pack<float, 2> &p) { return I == 0 ? p.v.x : p.v.y; }
template <int I> __device__ __forceinline__ void pk_set(Pack<float, 1> &p, float x) { p.v = x; }
template <int I> __device__ __forceinline__ void pk_set(Pack<double, 1> &p, double x) { p.v = x; }
template <int I> __device__ __forceinline__ void pk_set(Pack<double, 2> &p, double x) { if (I == 0) p.a = x; else p.b = x; }
template <int I> __device__ __forceinline__ void pk_set(Pack<float, 2> &p, float x) { if (I == 0) p.v.x = x; else p.v.y = x; }

// sum of the halves (folding a pack of per-sample sums at the end of the sample loop)
__device__ __forceinline__ float pk_hsum(const Pack<float, 1> &p) { return p.v; }
__device__ __forceinline__ double pk_hsum(const Pack<double, 1> &p) { return p.v; }
__device__ __forceinline__ double pk_hsum(const Pack<double, 2> &p) { return p.a + p.b; }
__device__ __forceinline__ float pk_hsum(const Pack<float, 2> &p) { return pk_get<0>(p) + pk_get<1>(p); }

// ---- arithmetic
#define BB_PK_SCALAR_OPS(real)                                                                                         \
    __device__ __forceinline__ Pack<real, 1> operator+(Pack<real, 1> a, Pack<real, 1> b) { return {a.v + b.v}; }      \
    __device__ __forceinline__ Pack<real, 1> operator-(Pack<real, 1> a, Pack<real, 1> b) { return {a.v - b.v}; }      \
    __device__ __forceinline__ Pack<real, 1> operator*(Pack<real, 1> a, Pack<real, 1> b) { return {a.v * b.v}; }      \
    __device__ __forceinline__ Pack<real, 1> pk_fma(Pack<real, 1> a, Pack<real, 1> b, Pack<real, 1> c) {              \
        return {fma(a.v, b.v, c.v)};                                                                                   \
    }
BB_PK_SCALAR_OPS(float)
BB_PK_SCALAR_OPS(double)
#undef BB_PK_SCALAR_OPS

__device__ __forceinline__ Pack<double, 2> operator+(Pack<double, 2> a, Pack<double, 2> b) { return {a.a + b.a, a.b + b.b}; }
__device__ __forceinline__ Pack<double, 2> operator-(Pack<double, 2> a, Pack<double, 2> b) { return {a.a - b.a, a.b - b.b}; }
__device__ __forceinline__ Pack<double, 2> operator*(Pack<double, 2> a, Pack<double, 2> b) { return {a.a * b.a, a.b * b.b}; }
__device__ __forceinline__ Pack<double, 2> pk_fma(Pack<double, 2> a, Pack<double, 2> b, Pack<double, 2> c) {
    return {fma(a.a, b.a, c.a), fma(a.b, b.b, c.b)};
}

// the sm_100 packed-fp32 intrinsics (crt/sm_100_rt.h) = PTX add / mul / fma.rn.f32x2
__device__ __forceinline__ Pack<float, 2> operator+(Pack<float, 2> a, Pack<float, 2> b) { return {__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ Pack<float, 2> operator-(Pack<float, 2> a, Pack<float, 2> b) {
    return {__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))};
}
__device__ __forceinline__ Pack<float, 2> operator*(Pack<float, 2> a, Pack<float, 2> b) { return {__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ Pack<float, 2> pk_fma(Pack<float, 2> a, Pack<float, 2> b, Pack<float, 2> c) {
    return {__ffma2_rn(a.v, b.v, c.v)};
}

// e^x per half.  fp32: one packed multiply by log2(e), then one MUFU.EX2 per half; fp64: libm.
__device__ __forceinline__ Pack<float, 1> pk_exp(Pack<float, 1> x) { return {bb_exp(x.v)}; }
__device__ __forceinline__ Pack<double, 1> pk_exp(Pack<double, 1> x) { return {exp(x.v)}; }
__device__ __forceinline__ Pack<double, 2> pk_exp(Pack<double, 2> x) { return {exp(x.a), exp(x.b)}; }
__device__ __forceinline__ Pack<float, 2> pk_exp(Pack<float, 2> x) {
    const Pack<float, 2> y = x * pk_bc<float, 2>(1.4426950408889634f);
    return pk_make(fast_ex2(pk_get<0>(y)), fast_ex2(pk_get<1>(y)));
}
// e^(s x) with the scale folded into the log2(e) multiply
__device__ __forceinline__ Pack<float, 1> pk_exp_scaled(Pack<float, 1> x, float s) {
    return {fast_ex2(x.v * (s * 1.4426950408889634f))};
}
__device__ __forceinline__ Pack<double, 1> pk_exp_scaled(Pack<double, 1> x, double s) { return {exp(s * x.v)}; }
__device__ __forceinline__ Pack<double, 2> pk_exp_scaled(Pack<double, 2> x, double s) { return {exp(s * x.a), exp(s * x.b)}; }
__device__ __forceinline__ Pack<float, 2> pk_exp_scaled(Pack<float, 2> x, float s) {
    const Pack<float, 2> y = x * pk_bc<float, 2>(s * 1.4426950408889634f);
    return pk_make(fast_ex2(pk_get<0>(y)), fast_ex2(pk_get<1>(y)));
}

}  // namespace bb
