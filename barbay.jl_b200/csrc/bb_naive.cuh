// BarBay.stats.naive_prior (src/stats.jl:1175-1359) on the device, from the PACKED count array -- the priors every
// documented workflow computes first and feeds into the models as matrix priors (docs/src/examples.md:121-140).
// SURVEY.md section 8f rank 3.  Input: DataArrays.bc_count (already incremented by the pseudocount, stats.jl:1185)
// in Julia memory order, as bb_desc.bc_count.  Three kernels, all fp64, results independent of the launch shape:
//   np_totals_kernel   n_t = sum_b r_tb per (replicate, t): exact 64-bit integer sums (bc_total, utils.jl:431-432)
//   np_loglam_kernel   log.(bc_count)[:]                                      (stats.jl:1345-1352)
//   np_neutral_kernel  per (replicate, t): mean and sd (n - 1) over the neutral barcodes of
//                      log((r_{t+1,n} / n_{t+1}) / (r_tn / n_t)) with the +-Inf entries left out
//                      (stats.jl:1199-1262, 1298, 1338) -- one CTA each, two fixed-order passes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

namespace bb {

constexpr int NP_THREADS = 256;
constexpr int NP_MAX_T = 64;         // time points per replicate handled by the totals kernel's shared accumulators

struct NpRep {                       // one replicate's block of the packed counts
    long long off;                   // offset of its T x B block
    int T;
    int out_off;                     // offset of its T - 1 entries in s_pop / logsig_pop
    int tot_off;                     // offset of its T totals
};

// totals[tot_off + t] += sum over the block's columns; grid = (chunks, replicates)
static __global__ void __launch_bounds__(NP_THREADS)
np_totals_kernel(const long long *cnt, const NpRep *reps, int B, unsigned long long *totals) {
    __shared__ unsigned long long acc[NP_MAX_T];
    const NpRep r = reps[blockIdx.y];
    for (int i = threadIdx.x; i < r.T; i += blockDim.x) acc[i] = 0ull;
    __syncthreads();
    const long long n = (long long)r.T * B;
    // element i of the block is (t = i % T, b = i / T): consecutive threads read consecutive words
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
    for (long long i0 = lo; i0 < hi; i0 += (long long)blockDim.x * 4) {
        unsigned long long v[4]; int t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + (long long)u * blockDim.x + threadIdx.x;
            v[u] = i < hi ? (unsigned long long)cnt[r.off + i] : 0ull;
            t[u] = i < hi ? (int)(i % r.T) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u]) atomicAdd(&acc[t[u]], v[u]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < r.T; i += blockDim.x)
        if (acc[i]) atomicAdd(&totals[r.tot_off + i], acc[i]);
}

static __global__ void __launch_bounds__(NP_THREADS)
np_loglam_kernel(const long long *cnt, long long n, double *out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = log((double)cnt[i]);
}

// fixed-order block sum: thread partials -> shared tree (the same tree whatever the data)
static __device__ __forceinline__ double np_block_sum(double v, double *sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = NP_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// grid = (max T - 1, replicates); CTA (t, r) handles the ratio t -> t + 1 of replicate r
static __global__ void __launch_bounds__(NP_THREADS)
np_neutral_kernel(const long long *cnt, const NpRep *reps, int N, const unsigned long long *totals,
                  double *s_pop, double *logsig_pop) {
    __shared__ double sh[NP_THREADS];
    const NpRep r = reps[blockIdx.y];
    const int t = blockIdx.x;
    if (t >= r.T - 1) return;
    const double n0 = (double)totals[r.tot_off + t], n1 = (double)totals[r.tot_off + t + 1];
    const long long *c = cnt + r.off;
    auto ratio = [&](int n) {
        const double f0 = (double)c[(long long)n * r.T + t] / n0, f1 = (double)c[(long long)n * r.T + t + 1] / n1;
        return log(f1 / f0);
    };
    double s = 0.0, m = 0.0;
    for (int n = threadIdx.x; n < N; n += NP_THREADS) {
        const double x = ratio(n);
        if (!isinf(x)) { s += x; m += 1.0; }
    }
    const double cntf = np_block_sum(m, sh);
    const double mean = np_block_sum(s, sh) / cntf;
    double q = 0.0;
    for (int n = threadIdx.x; n < N; n += NP_THREADS) {
        const double x = ratio(n);
        if (!isinf(x)) q += (x - mean) * (x - mean);
    }
    const double var = np_block_sum(q, sh) / (cntf - 1.0);
    if (threadIdx.x == 0) {
        s_pop[r.out_off + t] = -mean;                    // stats.jl:1298
        logsig_pop[r.out_off + t] = -sqrt(var);          // stats.jl:1338 (the reference's -std, not log(std))
    }
}

#define NP_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    throw std::runtime_error(std::string("bb_naive_prior: ") + cudaGetErrorString(e_)); } while (0)

// Host driver: copies the packed counts in, runs the three kernels on `device`, copies the priors out.
// s_pop / logsig_pop: sum_r (T_r - 1) entries (time fastest, then replicate); loglam: sum_r T_r B entries in the
// memory order of bc_count.  *launches (nullable) += kernels launched.
inline void naive_prior_device(const int64_t *bc_count, int n_rep, const int32_t *n_time, int n_neutral, int n_bc,
                               double *s_pop, double *logsig_pop, double *loglam, long long *launches) {
    if (!bc_count || !n_time || !s_pop || !logsig_pop || !loglam) throw std::runtime_error("bb_naive_prior: NULL argument");
    if (n_rep < 1 || n_neutral < 1 || n_bc < 0) throw std::runtime_error("bb_naive_prior: bad sizes");
    const int B = n_neutral + n_bc;
    std::vector<NpRep> reps(n_rep);
    long long off = 0; int out_off = 0, tot_off = 0, tmax = 0;
    for (int r = 0; r < n_rep; ++r) {
        if (n_time[r] < 2 || n_time[r] > NP_MAX_T)
            throw std::runtime_error("bb_naive_prior: every replicate needs between 2 and 64 time points");
        reps[r] = {off, n_time[r], out_off, tot_off};
        off += (long long)n_time[r] * B; out_off += n_time[r] - 1; tot_off += n_time[r];
        tmax = n_time[r] > tmax ? n_time[r] : tmax;
    }
    long long *d_cnt = nullptr; NpRep *d_reps = nullptr; unsigned long long *d_tot = nullptr;
    double *d_ll = nullptr, *d_sp = nullptr, *d_ls = nullptr;
    struct Free { void **p; ~Free() { if (*p) cudaFree(*p); } };
    Free f0{(void **)&d_cnt}, f1{(void **)&d_reps}, f2{(void **)&d_tot}, f3{(void **)&d_ll}, f4{(void **)&d_sp}, f5{(void **)&d_ls};
    NP_CUDA(cudaMalloc(&d_cnt, off * sizeof(long long)));
    NP_CUDA(cudaMalloc(&d_reps, n_rep * sizeof(NpRep)));
    NP_CUDA(cudaMalloc(&d_tot, tot_off * sizeof(unsigned long long)));
    NP_CUDA(cudaMalloc(&d_ll, off * sizeof(double)));
    NP_CUDA(cudaMalloc(&d_sp, out_off * sizeof(double)));
    NP_CUDA(cudaMalloc(&d_ls, out_off * sizeof(double)));
    NP_CUDA(cudaMemcpy(d_cnt, bc_count, off * sizeof(long long), cudaMemcpyHostToDevice));
    NP_CUDA(cudaMemcpy(d_reps, reps.data(), n_rep * sizeof(NpRep), cudaMemcpyHostToDevice));
    NP_CUDA(cudaMemset(d_tot, 0, tot_off * sizeof(unsigned long long)));
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const long long per_rep = (long long)tmax * B;
    int chunks = (int)((per_rep + NP_THREADS * 16 - 1) / (NP_THREADS * 16));
    chunks = chunks < 1 ? 1 : (chunks > 4 * sms ? 4 * sms : chunks);
    np_totals_kernel<<<dim3(chunks, n_rep), NP_THREADS>>>(d_cnt, d_reps, B, d_tot);
    long long lb = (off + NP_THREADS - 1) / NP_THREADS;
    np_loglam_kernel<<<(int)(lb < 8 * sms ? (lb < 1 ? 1 : lb) : 8 * sms), NP_THREADS>>>(d_cnt, off, d_ll);
    np_neutral_kernel<<<dim3(tmax - 1, n_rep), NP_THREADS>>>(d_cnt, d_reps, n_neutral, d_tot, d_sp, d_ls);
    NP_CUDA(cudaGetLastError());
    if (launches) *launches += 3;
    NP_CUDA(cudaMemcpy(loglam, d_ll, off * sizeof(double), cudaMemcpyDeviceToHost));
    NP_CUDA(cudaMemcpy(s_pop, d_sp, out_off * sizeof(double), cudaMemcpyDeviceToHost));
    NP_CUDA(cudaMemcpy(logsig_pop, d_ls, out_off * sizeof(double), cudaMemcpyDeviceToHost));
}

}  // namespace bb
