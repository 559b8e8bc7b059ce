// Derived `bc_fitness` rows of the hierarchical models, on the device.
//
// utils.advi_to_df -> process_hierarchical_samples! (src/utils.jl:1284-1343) draws n (default 10 000) samples of
// theta, log-tau and theta-tilde from their fitted Normals for every (barcode[, environment], replicate), forms
// s = theta + exp(log-tau) * theta-tilde and reports the sample MEDIAN (in the `mean` column, :1315) and the sample
// standard deviation.  On the host that is 3 x 10^4 normals per output row -- minutes at 10^6 barcodes, more than the
// whole fit.  Here one CTA per output row draws the samples from the Philox lattice into shared memory, takes the
// exact median with a radix select and the standard deviation from shifted double sums.
#pragma once
#include "bb_aux_kernels.cuh"

namespace bb {

constexpr uint32_t STREAM_DERIVED = 4;
constexpr int DERIVED_THREADS = 256;

template <typename real> struct DerivedArgs {
    const int *cols;             // [ncols] device column index of every mutant column of this shard
    int ncols, E, cpad, n;       // n samples per output row
    const vec2<real> *bc_th;     // [3 E][cpad]: (theta-tilde, log-tau, log-sigma) per environment
    const vec2<real> *hy_th;     // [H]
    const int *hgroup;           // [cpad]
    const int *map_tau;          // [cpad] rows (3 e + 1) of map_bc: reference index of log-tau
    long long off_tau;           // reference offset of the log-tau block
    PhiloxKey key;
    double *med, *sd;            // [E M R] in the order of the log-tau rows
};

__device__ __forceinline__ uint32_t float_key(float f) {       // order-preserving map float -> uint32
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of vals[0, n): four 8-bit passes, most significant first
static __device__ __forceinline__ float radix_select(const float *vals, int n, int k, unsigned *hist, unsigned *bcast) {
    uint32_t prefix = 0u, mask = 0u;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t kk = float_key(vals[i]);
            if ((kk & mask) == prefix) atomicAdd(&hist[(kk >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned cum = 0u;
            int b = 0;
            for (; b < 255; ++b) {
                if (cum + hist[b] > (unsigned)k) break;
                cum += hist[b];
            }
            bcast[0] = (unsigned)b; bcast[1] = cum;
        }
        __syncthreads();
        prefix |= bcast[0] << (8 * pass);
        mask |= 255u << (8 * pass);
        k -= (int)bcast[1];
        __syncthreads();
    }
    return key_float(prefix);
}

template <typename real>
__global__ void __launch_bounds__(DERIVED_THREADS) derived_fitness_kernel(const DerivedArgs<real> a) {
    extern __shared__ float dvals[];             // [n]
    __shared__ unsigned hist[256], bcast[2];
    __shared__ double red[2][DERIVED_THREADS / 32];
    const int tid = threadIdx.x;
    for (long long item = blockIdx.x; item < (long long)a.ncols * a.E; item += gridDim.x) {
        const int c = a.cols[item / a.E], e = (int)(item % a.E);
        const long long slot = (long long)a.map_tau[(size_t)(3 * e + 1) * a.cpad + c] - a.off_tau;
        const vec2<real> tt = a.bc_th[(size_t)(3 * e) * a.cpad + c], ta = a.bc_th[(size_t)(3 * e + 1) * a.cpad + c];
        const vec2<real> th = a.hy_th[a.hgroup[c] + e];
        const float m_th = (float)th.x, s_th = (float)softplus_d((double)th.y);
        const float m_ta = (float)ta.x, s_ta = (float)softplus_d((double)ta.y);
        const float m_tt = (float)tt.x, s_tt = (float)softplus_d((double)tt.y);
        const float shift = m_th + expf(m_ta) * m_tt;
        double s1 = 0.0, s2 = 0.0;
        for (int i = tid; i < a.n; i += DERIVED_THREADS) {
            uint32_t x[4];
            philox4x32((uint32_t)slot, (STREAM_DERIVED << 24) | (uint32_t)((unsigned long long)slot >> 32), (uint32_t)i, 0u, a.key, x);
            float n0, n1, n2, n3;
            box_muller(x[0], n0, n1, a.key.trig);
            box_muller(x[1], n2, n3, a.key.trig);
            const float s = fmaf(s_th, n0, m_th) + expf(fmaf(s_ta, n1, m_ta)) * fmaf(s_tt, n2, m_tt);
            dvals[i] = s;
            const double d = (double)(s - shift);
            s1 += d; s2 += d * d;
        }
        s1 = warp_sum<double>(s1); s2 = warp_sum<double>(s2);
        if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
        __syncthreads();
        const float lo = radix_select(dvals, a.n, (a.n - 1) / 2, hist, bcast);
        const float hi = (a.n & 1) ? lo : radix_select(dvals, a.n, a.n / 2, hist, bcast);
        if (tid == 0) {
            double t1 = 0.0, t2 = 0.0;
            for (int w = 0; w < DERIVED_THREADS / 32; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
            const double var = a.n > 1 ? (t2 - t1 * t1 / a.n) / (a.n - 1) : 0.0;
            a.med[slot] = 0.5 * ((double)lo + (double)hi);
            a.sd[slot] = sqrt(fmax(var, 0.0));
        }
        __syncthreads();
    }
}

}  // namespace bb
