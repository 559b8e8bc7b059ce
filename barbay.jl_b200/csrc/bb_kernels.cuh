// Column kernels of the ADVI step (sm_100a).
//
// One ADVI step = pass 1 (per-(replicate, time, sample) partial sums over all
// columns: Lambda_t and the weighted log-ratio sums) -> reduce (+ NCCL all-reduce)
// -> shared kernel (c_t, residual sums, population-latent gradient + update) ->
// pass 2 (per-column analytic gradient of the log-joint for all K samples, fused
// with the AdaGrad-family update of (mu, omega)).  The algebra is derived in
// DESIGN.md §3; it restates the log-joint of src/model_fitness_normal.jl:131-271
// (and the replicate / multienv / genotype variants) of the reference.
//
// Both passes regenerate the same eps from the Philox lattice instead of storing
// z: storing would add 2*K*w bytes per latent to a step whose algorithmic traffic
// is 8*w bytes per latent.
#pragma once
#include <type_traits>

#include "bb_types.cuh"

namespace bb {
#ifndef BB_P1_UNROLL
#define BB_P1_UNROLL 2
#endif
#ifndef BB_P2_MIN_BLOCKS
#define BB_P2_MIN_BLOCKS 6
#endif
#ifndef BB_P2_UNROLL
#define BB_P2_UNROLL 1
#endif
#ifndef BB_FUSE_MIN_BLOCKS
#define BB_FUSE_MIN_BLOCKS 3
#endif
#ifndef BB_FUSE_UNROLL
#define BB_FUSE_UNROLL 1
#endif

constexpr int kP1Unroll = BB_P1_UNROLL;   // MC samples of pass 1 interleaved per thread (ILP)
constexpr int kP2Unroll = BB_P2_UNROLL;
constexpr int kFuseUnroll = BB_FUSE_UNROLL;   // sample loops of the fused step kernel

template <int NT> struct TD { static constexpr int MAX = NT > 0 ? NT : MAX_NT_DYN; };
template <int NE> struct ED { static constexpr int MAX = NE > 0 ? NE : MAX_NE_DYN; };

template <int NT, int NE, bool HIER> struct Shape {
    static constexpr int PER = HIER ? 3 : 2;
    static constexpr int MAXT = TD<NT>::MAX;
    static constexpr int MAXE = ED<NE>::MAX;
    static constexpr int MAXJ = PER * MAXE;
    static constexpr int MAXC = ((MAXT + MAXJ + 7) / 8) * 8;   // noise slots, padded to whole Philox calls (8 normals)
};

__device__ __forceinline__ int find_segment(const SegList &sl, int blk) {
    int s = 0;
    for (int i = 1; i < sl.nseg; ++i)
        if (blk >= sl.seg[i].blk0) s = i;
    return s;
}

// eps for every latent of one column and one MC sample
template <typename real, int MAXC, bool SUP>
__device__ __forceinline__ void column_noise(real (&eps)[MAXC], int nclass, int nt, uint32_t colid, uint32_t k,
                                             uint32_t step, const PhiloxKey &key, const float2 *tab,
                                             const SupArgs<real> &sup, int c, int cpad, int tmax, int nj) {
    if constexpr (SUP) {
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            if (i >= nclass) break;
            eps[i] = (i < nt) ? sup.eps_lam[((size_t)k * tmax + i) * cpad + c]
                              : sup.eps_bc[((size_t)k * nj + (i - nt)) * cpad + c];
        }
    } else {
#pragma unroll
        for (int q = 0; q < MAXC / 8; ++q) {
            if (q * 8 >= nclass) break;
            real n[8];
            normals8<real>(colid, (STREAM_COLUMN << 24) | (uint32_t)q, k, step, key, tab, n);
#pragma unroll
            for (int l = 0; l < 8; ++l) eps[8 * q + l] = n[l];
        }
    }
}

// ===================================================================== staging
// Thread-private double buffering through shared memory with cp.async (LDGSTS): while a
// thread computes the K samples of one column it already has the next column's rows in
// flight, so the ~1 us HBM latency is paid once per kernel, not once per tile.  Each thread
// copies and reads only its own slot [row][tid]: no block barrier, no bank conflicts.
template <int BYTES> __device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// fp32 kernels keep the Box-Muller direction table (bb_device.cuh) at the head of their dynamic shared
// memory; the fp64 kernels use sincospi and reserve nothing.  stage() starts the copy as ONE cp.async group
// (all loads in flight at once: the kernel start is pure latency); the caller waits for that group and
// publishes it with a __syncthreads before the first draw.
template <typename real> struct TrigTab {
    static constexpr size_t BYTES = std::is_same<real, float>::value ? TRIG_N * sizeof(float2) : 0;
    template <bool LOAD> static __device__ __forceinline__ const float2 *stage(unsigned char *smem, const PhiloxKey &key) {
        if constexpr (BYTES == 0) return nullptr;
        float2 *t = reinterpret_cast<float2 *>(smem);
        if constexpr (LOAD) {
#pragma unroll
            for (int i = 0; i < TRIG_N / BLOCK; ++i)
                cp_async<sizeof(float2)>(t + i * BLOCK + threadIdx.x, key.trig + i * BLOCK + threadIdx.x);
        }
        cp_async_commit();
        return t;
    }
};

// softplus(w) = max(w, 0) + log1p(e^{-|w|}) and its derivative sigmoid(w).  fp32: x = e^{-|w|} from one
// MUFU ex2; log1p(x) is the alternating series to x^6 below x = 1/16 (truncation < 1e-8 relative, so a
// small sigma keeps full relative accuracy) and ln2 * lg2(1 + x) above it -- two MUFU ops, and the
// polynomial runs beside the lg2.  fp64: libm.
__device__ __forceinline__ float softplus_f32(float w, float &x) {
    x = fast_ex2(-fabsf(w) * 1.4426950408889634f);                  // e^{-|w|} in (0, 1]
    const float big = 0.6931471805599453f * fast_lg2(1.0f + x);
    float p = fmaf(x, -1.0f / 6.0f, 0.2f);
    p = fmaf(x, p, -0.25f);
    p = fmaf(x, p, 1.0f / 3.0f);
    p = fmaf(x, p, -0.5f);
    p = fmaf(x, p, 1.0f);
    const float l1p = x < 0.0625f ? x * p : big;
    return fmaxf(w, 0.0f) + l1p;
}
__device__ __forceinline__ void softplus_sigmoid(float w, float &sp, float &sgm) {
    float x;
    sp = softplus_f32(w, x);
    const float ru = fast_rcp(1.0f + x);
    sgm = w >= 0.0f ? ru : x * ru;
}
__device__ __forceinline__ void softplus_sigmoid(double w, double &sp, double &sgm) {
    sp = softplus(w);
    sgm = sigmoid(w);
}
template <typename real> __device__ __forceinline__ real softplus_only(real w) {
    real sp, sg;
    softplus_sigmoid(w, sp, sg);
    return sp;
}

// ===================================================================== pass 1
// Accumulator layout of one sample: `pva` rows of BLOCK reals.  Scalar layout: slot v of thread t at
// v * BLOCK + t.  Pair layout (compile-time T, one environment): slots 2p, 2p+1 of thread t sit side by side
// at (p * BLOCK + t) * 2, so a sample's read-modify-write is one 64/128-bit LDS + STS per pair.
template <int NT, int NE> struct AccLayout {
    static constexpr bool PAIRS = NT > 0 && NE == 1;
    static __device__ __forceinline__ int rows(int pvs) { return PAIRS ? (pvs + 1) & ~1 : pvs; }
};

// Block reduction of the thread-private accumulators -> part[(k0 + k) * pv + v][block], in double and in a
// fixed order.  A warp owns whole samples (k = warp, warp + 4, ...) and reduces four slots at a time with
// a transposed butterfly: after the xor-16 and xor-8 exchanges every lane carries ONE of the four rows
// (row = lane >> 3 of its octet class), so the four 32-lane sums cost 6 double shuffles instead of 20.
// The kernel tail is pure latency: no divisions, few instructions.
template <typename real, bool PAIRS>
__device__ __forceinline__ void flush_rows(const real *sacc, int nk, int pvs, int pva, int k0, int pv, double *part) {
    using r2 = vec2<real>;
    constexpr int NW = BLOCK / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    for (int k = warp; k < nk; k += NW) {
        const real *base = sacc + (size_t)k * pva * BLOCK;
        for (int v0 = 0; v0 < pvs; v0 += 4) {
            double s[4] = {0.0, 0.0, 0.0, 0.0};
            if constexpr (PAIRS) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (v0 + 2 * q < pvs) {
                        const r2 *rp = reinterpret_cast<const r2 *>(base) + (size_t)((v0 >> 1) + q) * BLOCK;
#pragma unroll
                        for (int j = 0; j < NW; ++j) {
                            const r2 t = rp[j * 32 + lane];
                            s[2 * q] += (double)t.x; s[2 * q + 1] += (double)t.y;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (v0 + u < pvs) {
                        const real *rp = base + (size_t)(v0 + u) * BLOCK;
#pragma unroll
                        for (int j = 0; j < NW; ++j) s[u] += (double)rp[j * 32 + lane];
                    }
                }
            }
            // rows {0,1} stay in lanes 0-15, rows {2,3} in lanes 16-31; then row parity by bit 3
            double ka = hi16 ? s[2] : s[0], kb = hi16 ? s[3] : s[1];
            ka += __shfl_xor_sync(0xffffffffu, hi16 ? s[0] : s[2], 16);
            kb += __shfl_xor_sync(0xffffffffu, hi16 ? s[1] : s[3], 16);
            double kk = hi8 ? kb : ka;
            kk += __shfl_xor_sync(0xffffffffu, hi8 ? ka : kb, 8);
            kk += __shfl_xor_sync(0xffffffffu, kk, 4);
            kk += __shfl_xor_sync(0xffffffffu, kk, 2);
            kk += __shfl_xor_sync(0xffffffffu, kk, 1);
            const int v = v0 + (lane >> 3);
            if ((lane & 7) == 0 && v < pvs)
                part[((size_t)(k0 + k) * pv + v) * gridDim.x + blockIdx.x] = kk;   // [k][v][block]: coalesced for the reducer
        }
    }
}

// One MC sample of one column for pass 1: z = mu + sigma eps and the column's contribution to every slot,
// added into the thread-private accumulators `acc` (already offset to this sample and thread).
// zth[e]: the hyper latent's draw (hierarchical models).
template <typename real, int NT, int NE, bool HIER>
__device__ __forceinline__ void pass1_sample(const real *eps, const real *mu, const real *sg, const real *mub,
                                             const real *sgb, const real *zth, bool neutral, int nt, int ne,
                                             const int *env_of_t, real *acc) {
    using S = Shape<NT, NE, HIER>;
    using r2 = vec2<real>;
    constexpr bool PAIRS = AccLayout<NT, NE>::PAIRS;
    constexpr int NV = PAIRS ? ((3 * NT - 2 + 1) & ~1) : 2;
    real v[NV];
    if constexpr (PAIRS) v[NV - 1] = real(0);        // pad slot of an odd population
    auto add = [&](int slot, real x) {
        if constexpr (PAIRS) v[slot] = x; else acc[slot * BLOCK] += x;
    };
    real z[S::MAXT];
#pragma unroll
    for (int t = 0; t < S::MAXT; ++t) {
        if (t >= nt) break;
        z[t] = fma(sg[t], eps[t], mu[t]);
        add(t, bb_exp(z[t]));
    }
    if (neutral) {
#pragma unroll
        for (int t = 0; t < S::MAXT - 1; ++t) {
            if (t >= nt - 1) break;
            const real d = z[t + 1] - z[t];
            add(nt + t, d);
            add(2 * nt - 1 + t, d * d);
        }
    } else {
        real zs[S::MAXE], w[S::MAXE];
#pragma unroll
        for (int e = 0; e < S::MAXE; ++e) {
            if (e >= ne) break;
            if constexpr (HIER) {
                const real ztt = fma(sgb[3 * e], eps[nt + 3 * e], mub[3 * e]);
                const real ztau = fma(sgb[3 * e + 1], eps[nt + 3 * e + 1], mub[3 * e + 1]);
                zs[e] = fma(bb_exp(ztau), ztt, zth[e]);
                w[e] = bb_exp(real(-2) * fma(sgb[3 * e + 2], eps[nt + 3 * e + 2], mub[3 * e + 2]));
            } else {
                zs[e] = fma(sgb[2 * e], eps[nt + 2 * e], mub[2 * e]);
                w[e] = bb_exp(real(-2) * fma(sgb[2 * e + 1], eps[nt + 2 * e + 1], mub[2 * e + 1]));
            }
        }
#pragma unroll
        for (int t = 0; t < S::MAXT - 1; ++t) {
            if (t >= nt - 1) break;
            const int e = NE == 1 ? 0 : env_of_t[t + 1];
            add(nt + t, w[e] * (z[t + 1] - z[t] - zs[e]));
            if (NE != 1) add(2 * nt - 1 + t, w[e]);
        }
        if (NE == 1) add(2 * nt - 1, w[0]);
    }
    if constexpr (PAIRS) {
        r2 *a2 = reinterpret_cast<r2 *>(acc);
        const int np = neutral ? (3 * NT - 1) / 2 : NT;
#pragma unroll
        for (int q = 0; q < NV / 2; ++q) {
            if (q >= np) break;
            r2 t2 = a2[q * BLOCK];
            t2.x += v[2 * q]; t2.y += v[2 * q + 1];
            a2[q * BLOCK] = t2;
        }
    }
}

// Slots per sample (pv = nt + 2 (nt-1)):
//   [0, nt)              Lambda_t partial            (both populations)
//   [nt, 2nt-1)          neutral: sum d_t            mutant: sum w (d_t - s)
//   [2nt-1, 3nt-2)       neutral: sum d_t^2          mutant: sum w        (E == 1: slot 2nt-1 only)
// smem: [kchunk][pv][BLOCK] thread-private accumulators, then 2 staging buffers [nt + nj][BLOCK] of (mu, omega)
template <typename real, int NT, int NE, bool HIER, bool SUP>
__global__ void __launch_bounds__(BLOCK) pass1_kernel(const P1Args<real> a) {
    using S = Shape<NT, NE, HIER>;
    using r2 = vec2<real>;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const float2 *strig = TrigTab<real>::template stage<!SUP>(smem_all, a.key);   // published by the first barrier below
    unsigned char *smem_raw = smem_all + TrigTab<real>::BYTES;
    real *sacc = reinterpret_cast<real *>(smem_raw);

    const int tid = threadIdx.x;
    const int sidx = find_segment(a.segs, blockIdx.x);
    const Seg seg = a.segs.seg[sidx];
    const int nt = NT > 0 ? NT : seg.nt;
    const int ne = NE > 0 ? NE : a.ne;
    const int nj = S::PER * ne;
    const int pv = a.pv;
    const ColArrays<real> &C = a.cols;
    const int cpad = C.cpad;
    const int nblk = seg.blk1 - seg.blk0;
    const int ntile = (seg.ncol + BLOCK - 1) / BLOCK;
    // accumulator slots this block's population actually uses; the smem budget (a.acc_slots rows of
    // BLOCK reals) is sized by the host for the mutant population, the few neutral blocks sweep K in
    // smaller chunks instead of inflating every block's shared memory
    const int pvs = seg.neutral ? 3 * nt - 2 : (NE == 1 ? 2 * nt : 3 * nt - 2);
    constexpr bool PAIRS = AccLayout<NT, NE>::PAIRS;
    const int pva = AccLayout<NT, NE>::rows(pvs);
    const int kchunk = max(1, min(a.K, a.acc_slots / pva));
    const size_t acc_bytes = (((size_t)a.acc_slots * BLOCK * sizeof(real)) + 15) / 16 * 16;
    r2 *stage = reinterpret_cast<r2 *>(smem_raw + acc_bytes);          // [2][nt + nj][BLOCK]
    const int stage_stride = (C.tmax + C.nj) * BLOCK;

    auto prefetch = [&](int tile, int buf) {
        const int i = tile * BLOCK + tid;
        if (tile < ntile && i < seg.ncol) {
            const int c = seg.col0 + i;
            r2 *dst = stage + (size_t)buf * stage_stride + tid;
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                cp_async<sizeof(r2)>(dst + t * BLOCK, C.lam_th + (size_t)t * cpad + c);
            }
            if (!seg.neutral) {
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    cp_async<sizeof(r2)>(dst + (nt + j) * BLOCK, C.bc_th + (size_t)j * cpad + c);
                }
            }
        }
        cp_async_commit();
    };

    cp_async_wait<0>();                 // the direction table; published by the first barrier below
    for (int kc0 = 0; kc0 < a.K; kc0 += kchunk) {
        const int kc1 = min(a.K, kc0 + kchunk);
        for (int i = tid; i < (kc1 - kc0) * pva * BLOCK; i += BLOCK) sacc[i] = real(0);
        __syncthreads();

        int buf = 0;
        const int nbuf = a.nbuf;
        prefetch(blockIdx.x - seg.blk0, 0);
        for (int tile = blockIdx.x - seg.blk0; tile < ntile; tile += nblk, buf ^= (nbuf - 1)) {
            if (nbuf == 2) { prefetch(tile + nblk, buf ^ 1); cp_async_wait<1>(); }
            else { if (tile != (int)blockIdx.x - seg.blk0) prefetch(tile, 0); cp_async_wait<0>(); }
            const int i = tile * BLOCK + tid;
            if (i >= seg.ncol) continue;
            const int c = seg.col0 + i;
            const uint32_t colid = C.col_id ? C.col_id[c] : seg.colid0 + (uint32_t)i;
            const r2 *src = stage + (size_t)buf * stage_stride + tid;

            real mu[S::MAXT], sg[S::MAXT];
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                const r2 th = src[t * BLOCK];
                mu[t] = th.x; sg[t] = softplus_only<real>(th.y);
                if constexpr (SUP) if (a.sup.z_direct) { mu[t] = real(0); sg[t] = real(1); }
            }
            real mub[S::MAXJ], sgb[S::MAXJ];
            int hbase = 0;
            if (!seg.neutral) {
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    const r2 th = src[(nt + j) * BLOCK];
                    mub[j] = th.x; sgb[j] = softplus_only<real>(th.y);
                    if constexpr (SUP) if (a.sup.z_direct) { mub[j] = real(0); sgb[j] = real(1); }
                }
                if constexpr (HIER) hbase = C.hgroup[c];
            }
            const int nclass = seg.neutral ? nt : nt + nj;

            // hierarchical models: the hyper latent's draw of sample k + 1 is loaded while sample k is computed
            // (an L2 round trip per sample otherwise sits on the critical path of every warp)
            real zth_nxt[S::MAXE];
            auto load_zth = [&](int k) {
                if constexpr (HIER) {
                    if (!seg.neutral) {
#pragma unroll
                        for (int e = 0; e < S::MAXE; ++e) {
                            if (e >= ne) break;
                            zth_nxt[e] = a.hy_zeps[(size_t)k * a.hz_k + (size_t)(hbase + e) * a.hz_h].x;
                        }
                    }
                }
            };
            load_zth(kc0);
#pragma unroll kP1Unroll
            for (int k = kc0; k < kc1; ++k) {
                real zth[S::MAXE];
                if constexpr (HIER) {
#pragma unroll
                    for (int e = 0; e < S::MAXE; ++e) zth[e] = zth_nxt[e];
                    if (k + 1 < kc1) load_zth(k + 1);
                }
                real eps[S::MAXC];
                column_noise<real, S::MAXC, SUP>(eps, nclass, nt, colid, (uint32_t)k, a.step, a.key, strig,
                                                 a.sup, c, cpad, C.tmax, C.nj);
                pass1_sample<real, NT, NE, HIER>(eps, mu, sg, mub, sgb, zth, seg.neutral, nt, ne, a.env_of_t,
                                                 sacc + (size_t)(k - kc0) * pva * BLOCK + (PAIRS ? 2 * tid : tid));
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        // block reduction of the private columns, in double, fixed order
        flush_rows<real, PAIRS>(sacc, kc1 - kc0, pvs, pva, kc0, pv, a.part);
        __syncthreads();
    }
}

// As-written neutral pairing of the ragged replicate model (replicates.jl:599-605): the shared-latent kernel pairs the
// neutral ratios itself (shared_body, bb_aux_kernels.cuh) and needs d[t] = z[t+1] - z[t] of every neutral column and
// sample -- the same draws as pass 1 (same lattice call, or the same supplied noise).  A few hundred columns: a kernel
// of its own (block k = sample k) keeps pass 1 exactly as it is for every other fit.
template <typename real, bool SUP>
__global__ void __launch_bounds__(BLOCK) aw_export_kernel(const AwArgs<real> a) {
    using r2 = vec2<real>;
    constexpr bool F32 = std::is_same<real, float>::value;
    __shared__ float2 strig[F32 ? TRIG_N : 1];
    if constexpr (F32) {
        for (int i = threadIdx.x; i < TRIG_N; i += BLOCK) strig[i] = a.key.trig[i];
        __syncthreads();
    }
    const ColArrays<real> &C = a.cols;
    const int k = blockIdx.x;
    for (int si = 0; si < a.segs.nseg; ++si) {
        const Seg &seg = a.segs.seg[si];
        const int nt = seg.nt;
        for (int i = threadIdx.x; i < seg.ncol; i += BLOCK) {
            const int c = seg.col0 + i;
            const uint32_t colid = C.col_id ? C.col_id[c] : seg.colid0 + (uint32_t)i;
            real eps[MAX_NT_DYN];
            column_noise<real, MAX_NT_DYN, SUP>(eps, nt, nt, colid, (uint32_t)k, a.step, a.key, strig, a.sup, c, C.cpad,
                                                C.tmax, C.nj);
            real *out = a.d + (((size_t)seg.rep * a.K + k) * a.N + i) * (C.tmax - 1);
            real zprev = real(0);
#pragma unroll
            for (int t = 0; t < MAX_NT_DYN; ++t) {
                if (t >= nt) break;
                const r2 th = C.lam_th[(size_t)t * C.cpad + c];
                real mu = th.x, sg = softplus_only<real>(th.y);
                if constexpr (SUP) if (a.sup.z_direct) { mu = real(0); sg = real(1); }
                const real z = fma(sg, eps[t], mu);
                if (t > 0) out[t - 1] = z - zprev;
                zprev = z;
            }
        }
    }
}

// Pass-1 accumulation of samples [k0, k1) of one non-hierarchical column into the thread-private
// accumulators `acc0` (offset to this thread, sample k0; same layout as pass1_kernel).  Used by the fused step kernel.
template <typename real, int NT, int NE>
__device__ __forceinline__ void pass1_column_direct(const real *mu, const real *sg, const real *mub, const real *sgb,
                                                    bool neutral, int nt, int ne, int k0, int k1, uint32_t colid,
                                                    uint32_t step, const PhiloxKey &key, const float2 *tab,
                                                    const int *env_of_t, real *acc0, int pva) {
    using S = Shape<NT, NE, false>;
    const int nclass = neutral ? nt : nt + 2 * ne;
    const SupArgs<real> nosup{};
#pragma unroll kFuseUnroll
    for (int k = k0; k < k1; ++k) {
        real eps[S::MAXC];
        column_noise<real, S::MAXC, false>(eps, nclass, nt, colid, (uint32_t)k, step, key, tab, nosup, 0, 0, 0, 0);
        pass1_sample<real, NT, NE, false>(eps, mu, sg, mub, sgb, nullptr, neutral, nt, ne, env_of_t,
                                          acc0 + (size_t)(k - k0) * pva * BLOCK);
    }
}

// ===================================================================== optimiser
// AdvancedVI 0.2 optimisers.jl (restated in oracle/advi_ref.py):
//   DecayedADAGrad   acc = post*acc + pre*g^2 ; delta = eta*g / (sqrt(acc) + 1e-8)
//   TruncatedADAGrad s   = sum of the last n g^2; delta = eta*g / (tau + sqrt(s) + 1e-8)
// The window sum is kept in running form, s <- max(s - evicted, 0) + g^2.  The clamp sits BEFORE the new term: in
// fp32, once the huge squared gradients of the first steps (1e14 for a 1e7-read barcode) have been evicted, the
// cancellation residue of s - evicted can be any value of magnitude ulp(1e14) ~ 1e7, negative included; clamping
// s - evicted + g^2 as a whole would then return 0 for a latent whose current g is large and turn the bounded
// AdaGrad step (|delta| <= eta, as s >= g^2 in exact arithmetic) into eta * g / tau -- a jump that ends in inf / NaN
// (seen after ~2 300 steps of the documented naive-prior workflow).  With the clamp inside, s >= g^2 always holds.
template <typename real>
__device__ __forceinline__ void opt_apply(const OptArgsT<real> &o, real g, real &theta, real &acc, real &ring_slot) {
    const real g2 = g * g;
    real denom;
    if (o.kind == 1) {
        acc = fma(o.post, acc, o.tau * g2);
        denom = bb_sqrt(acc) + real(1e-8);
    } else {
        acc = fmax(acc - ring_slot, real(0)) + g2;
        ring_slot = g2;
        denom = o.tau + bb_sqrt(acc) + real(1e-8);
    }
    theta -= o.eta * g * bb_rcp(denom);
}

// finish one latent: gradients of the objective, update (or emit), in place.
// MODE 0: DecayedADAGrad update | 1: TruncatedADAGrad update | 2: emit (dELBO/dmu, dELBO/domega) | 3: nothing.
// The mode is block-uniform, so each instantiation is a straight line of ~40 instructions.
template <typename real, int MODE>
__device__ __forceinline__ void finish_latent_mode(const OptArgsT<real> &o, real invK, real sgrad, real sgrade,
                                                   real sigma, vec2<real> &th, vec2<real> ac, vec2<real> rg, vec2<real> *th_ptr,
                                                   vec2<real> *acc_ptr, vec2<real> *ring_ptr, vec2<real> *gout_ptr) {
    // rg: the ring slot evicted this step (MODE 1 only); th is updated in place (MODE 0 / 1)
    if constexpr (MODE == 3) return;
    // d ELBO / d mu = mean_k g ; d ELBO / d omega = (mean_k g eps + 1/sigma) sigmoid(omega), and
    // sigmoid(w) = exp(w - softplus(w)): sigma is already at hand from the prologue
    const real sgm = bb_exp(th.y - sigma);
    const real gm = sgrad * invK;
    const real go = (sgrade * invK + bb_rcp(sigma)) * sgm;
    if constexpr (MODE == 2) {
        *gout_ptr = mk2<real>(gm, go);
    } else {
        const real g0 = -gm, g1 = -go;                  // the engine minimises -ELBO
        const real q0 = g0 * g0, q1 = g1 * g1;
        real d0, d1;
        if constexpr (MODE == 0) {
            ac.x = fma(o.post, ac.x, o.tau * q0);
            ac.y = fma(o.post, ac.y, o.tau * q1);
            d0 = bb_sqrt(ac.x) + real(1e-8);
            d1 = bb_sqrt(ac.y) + real(1e-8);
        } else {
            ac.x = fmax(ac.x - rg.x, real(0)) + q0;
            ac.y = fmax(ac.y - rg.y, real(0)) + q1;
            *ring_ptr = mk2<real>(q0, q1);
            d0 = o.tau + bb_sqrt(ac.x) + real(1e-8);
            d1 = o.tau + bb_sqrt(ac.y) + real(1e-8);
        }
        th.x -= o.eta * g0 * bb_rcp(d0);
        th.y -= o.eta * g1 * bb_rcp(d1);
        *th_ptr = th;
        *acc_ptr = ac;
    }
}

// The same update for the `n` (<= N) latents of one class of a column, stage by stage over all of them: the
// latents are independent, so their MUFU chains (ex2, rcp | sqrt x2 | rcp x2 | softplus of the new omega)
// overlap instead of running one latent after the other.  Rows are strided: smem rows by BLOCK, global
// rows by `cpad`; all pointers are already offset to this thread's column.  FUSE: returns the updated
// mu / sigma in place for the next step's pass 1.
template <typename real, int MODE, int N, bool FUSE>
__device__ __forceinline__ void finish_batch(const OptArgsT<real> &o, real invK, int n, const real *sgrad,
                                             const real *sgrade, real *mu, real *sigma, const vec2<real> *sth,
                                             const vec2<real> *sac, const vec2<real> *srg, vec2<real> *g_th,
                                             vec2<real> *g_acc, vec2<real> *g_ring, vec2<real> *g_out, size_t cpad) {
    using r2 = vec2<real>;
    if constexpr (MODE == 3) return;
    r2 th[N], ac[N], rg[MODE == 1 ? N : 1];
    real g0[N], g1[N];
    // every load of the batch is issued before its first store: unstaged accumulators / ring slots come from
    // global memory, and a load behind a (possibly aliasing) store would pay its DRAM latency alone
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i >= n) break;
        th[i] = sth[i * BLOCK];
        if constexpr (MODE < 2) ac[i] = sac ? sac[i * BLOCK] : g_acc[(size_t)i * cpad];
        if constexpr (MODE == 1) rg[i] = srg ? srg[i * BLOCK] : g_ring[(size_t)i * cpad];   // the ring slot evicted this step
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i >= n) break;
        // d ELBO / d mu = mean_k g ; d ELBO / d omega = (mean_k g eps + 1/sigma) sigmoid(omega), and
        // sigmoid(w) = exp(w - softplus(w)): sigma is already at hand from the prologue
        const real sgm = bb_exp(th[i].y - sigma[i]);
        const real gm = sgrad[i] * invK;
        const real go = (sgrade[i] * invK + bb_rcp(sigma[i])) * sgm;
        g0[i] = -gm; g1[i] = -go;                       // the engine minimises -ELBO
        if constexpr (MODE == 2) g_out[(size_t)i * cpad] = mk2<real>(gm, go);
    }
    if constexpr (MODE < 2) {
        real d0[N], d1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i >= n) break;
            const real q0 = g0[i] * g0[i], q1 = g1[i] * g1[i];
            if constexpr (MODE == 0) {
                ac[i].x = fma(o.post, ac[i].x, o.tau * q0);
                ac[i].y = fma(o.post, ac[i].y, o.tau * q1);
                d0[i] = bb_sqrt(ac[i].x) + real(1e-8);
                d1[i] = bb_sqrt(ac[i].y) + real(1e-8);
            } else {
                ac[i].x = fmax(ac[i].x - rg[i].x, real(0)) + q0;
                ac[i].y = fmax(ac[i].y - rg[i].y, real(0)) + q1;
                g_ring[(size_t)i * cpad] = mk2<real>(q0, q1);
                d0[i] = o.tau + bb_sqrt(ac[i].x) + real(1e-8);
                d1[i] = o.tau + bb_sqrt(ac[i].y) + real(1e-8);
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i >= n) break;
            th[i].x -= o.eta * g0[i] * bb_rcp(d0[i]);
            th[i].y -= o.eta * g1[i] * bb_rcp(d1[i]);
            g_th[(size_t)i * cpad] = th[i];
            g_acc[(size_t)i * cpad] = ac[i];
        }
        if constexpr (FUSE) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (i >= n) break;
                mu[i] = th[i].x; sigma[i] = softplus_only<real>(th[i].y);
            }
        }
    }
}

template <typename real>
__device__ __forceinline__ void finish_latent(const OptArgsT<real> &o, real invK, real sgrad, real sgrade,
                                              vec2<real> th, vec2<real> ac, vec2<real> *th_ptr,
                                              vec2<real> *acc_ptr, vec2<real> *ring_ptr, vec2<real> *gout_ptr) {
    const real sigma = softplus_only<real>(th.y);
    if (o.update) {
        const vec2<real> z2 = mk2<real>(0, 0);
        if (o.kind == 1) finish_latent_mode<real, 0>(o, invK, sgrad, sgrade, sigma, th, ac, z2, th_ptr, acc_ptr, ring_ptr, gout_ptr);
        else finish_latent_mode<real, 1>(o, invK, sgrad, sgrade, sigma, th, ac, *ring_ptr, th_ptr, acc_ptr, ring_ptr, gout_ptr);
    } else if (gout_ptr) {
        finish_latent_mode<real, 2>(o, invK, sgrad, sgrade, sigma, th, ac, mk2<real>(0, 0), th_ptr, acc_ptr, ring_ptr, gout_ptr);
    }
}

// ===================================================================== pass 2
// smem: ctx [K][3][tmax] of this block's replicate | ELBO private columns [K+1][BLOCK] (double) |
//       2 staging buffers { theta [nt+nj][BLOCK], acc [nt+nj][BLOCK], counts [nt][BLOCK],
//                           priors [nt+nj][BLOCK] (matrix priors only) }
// FUSE: after updating a column, accumulate the pass-1 sums of the NEXT step (noise of step + 1, the
// freshly updated theta) -- the software-pipelined step: theta and its accumulators are read and
// written exactly once per ADVI step (non-hierarchical models; the hyper latents of the hierarchical
// ones are only known after every member column has been processed).
template <typename real, int NT, int NE, bool HIER, bool SUP, bool ELBO, bool FUSE = false>
// register budget: 6 CTAs per SM (80 registers) for the one-environment kernels; the runtime-environment
// kernels carry up to 8 x (s, log sigma[, log tau]) per column and would spill kilobytes at 80 -> 3 CTAs (168)
__global__ void __launch_bounds__(BLOCK, FUSE ? BB_FUSE_MIN_BLOCKS : (NE == 1 ? BB_P2_MIN_BLOCKS : 3))
pass2_kernel(const P2Args<real> a) {
    using S = Shape<NT, NE, HIER>;
    using r2 = vec2<real>;
    extern __shared__ __align__(16) unsigned char smem_all2[];
    const float2 *strig = TrigTab<real>::template stage<!SUP>(smem_all2, a.key);   // published by the barrier after the ctx load
    unsigned char *smem_raw = smem_all2 + TrigTab<real>::BYTES;
    real *sctx = reinterpret_cast<real *>(smem_raw);
    // per-sample context rows {c_t - sbar_t | (U_{t-1} - U_t) / Lambda_t | exp(-2 logsigma-bar_t)}: staged as
    // [K][CS] with 16-byte rows so a sample's context is a few vector loads (compile-time T), not 3 T - 2 scalars
    constexpr int VW = 16 / (int)sizeof(real);
    const int TT = NT > 0 ? NT : a.tmax_ctx;
    const int CS = ((3 * TT + VW - 1) / VW) * VW;
    constexpr int CSC = NT > 0 ? ((3 * NT + VW - 1) / VW) * VW : VW;
    const size_t ctx_bytes = (size_t)a.K * (((3 * a.tmax_ctx + VW - 1) / VW) * VW) * sizeof(real);
    double *sel = reinterpret_cast<double *>(smem_raw + ctx_bytes);
    const size_t sel_bytes = ELBO ? (size_t)(a.K + 1) * BLOCK * sizeof(double) : 0;

    const int tid = threadIdx.x;
    const int sidx = find_segment(a.segs, blockIdx.x);
    const Seg seg = a.segs.seg[sidx];
    const int nt = NT > 0 ? NT : seg.nt;
    const int ne = NE > 0 ? NE : a.ne;
    const int nj = S::PER * ne;
    const ColArrays<real> &C = a.cols;
    const int cpad = C.cpad;
    const int nblk = seg.blk1 - seg.blk0;
    const int ntile = (seg.ncol + BLOCK - 1) / BLOCK;
    constexpr bool want_elbo = ELBO;
    const real invK = real(1) / real(a.K);
    const bool lam_mat = C.lam_pr != nullptr, bc_mat = C.bc_pr != nullptr;

    // staging geometry (rows sized by the arrays' tmax / nj so every block agrees with the host)
    const int rows = C.tmax + C.nj;
    const size_t th_bytes = (size_t)rows * BLOCK * sizeof(r2);
    const size_t cn_bytes = (size_t)C.tmax * BLOCK * sizeof(int);
    // what is staged (host decides by the shared-memory budget): theta + counts always; accumulators,
    // matrix priors and the evicted TruncatedADAGrad ring slot when they fit.  What the sample loop reads
    // (theta, priors, counts) is prefetched one tile ahead into one of `nbuf` buffers; what only the
    // update epilogue reads (accumulators, ring slot) has ONE buffer, filled while the samples run.
    const int nac = a.stage_acc ? 1 : 0;
    const int npr = (a.stage_acc && a.stage_pr) ? 1 : 0;
    const int nrg = (a.stage_acc && a.stage_ring) ? 1 : 0;
    const size_t buf_bytes = (1 + npr) * th_bytes + cn_bytes;            // theta, [priors], counts
    const int nbuf = a.nbuf;
    unsigned char *stage0 = smem_raw + ctx_bytes + sel_bytes;
    unsigned char *epi0 = stage0 + (size_t)nbuf * buf_bytes;             // [acc], [ring]
    // fused step: thread-private pass-1 accumulators behind the staging buffers, a.acc_slots rows of
    // BLOCK reals sized for the mutant population ([K][pvs]); mutant blocks accumulate inline right
    // after updating a column, the few neutral blocks (3 nt - 2 slots per sample) sweep their own
    // columns again in sample chunks once their updates are written
    const int pvs = seg.neutral ? 3 * nt - 2 : (NE == 1 ? 2 * nt : 3 * nt - 2);
    constexpr bool PAIRS = AccLayout<NT, NE>::PAIRS;
    const int pva = AccLayout<NT, NE>::rows(pvs);
    real *facc = reinterpret_cast<real *>(epi0 + (size_t)(nac + nrg) * th_bytes);
    if constexpr (FUSE)
        for (int i = tid; i < a.acc_slots * BLOCK; i += BLOCK) facc[i] = real(0);

    // element offsets are 32-bit (t * cpad + c < 2^31: checked by the host), one IMAD.WIDE per address; the
    // optional sets (matrix priors, ring slot) sit in their own uniformly-branched loops, not predicated
    auto prefetch = [&](int tile, int buf) {
        const int i = tile * BLOCK + tid;
        if (tile < ntile && i < seg.ncol) {
            const uint32_t c = (uint32_t)(seg.col0 + i);
            unsigned char *base = stage0 + (size_t)buf * buf_bytes;
            r2 *sth = reinterpret_cast<r2 *>(base) + tid;
            r2 *spr = reinterpret_cast<r2 *>(base + th_bytes) + tid;
            int *scn = reinterpret_cast<int *>(base + (1 + npr) * th_bytes) + tid;
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                const uint32_t o = (uint32_t)t * (uint32_t)cpad + c;
                cp_async<sizeof(r2)>(sth + t * BLOCK, C.lam_th + o);
                cp_async<4>(scn + t * BLOCK, C.cnt + o);
            }
            if (!seg.neutral) {
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    cp_async<sizeof(r2)>(sth + (nt + j) * BLOCK, C.bc_th + ((uint32_t)j * (uint32_t)cpad + c));
                }
            }
            if (npr) {
                if (lam_mat) {
#pragma unroll
                    for (int t = 0; t < S::MAXT; ++t) {
                        if (t >= nt) break;
                        cp_async<sizeof(r2)>(spr + t * BLOCK, C.lam_pr + ((uint32_t)t * (uint32_t)cpad + c));
                    }
                }
                if (bc_mat && !seg.neutral) {
#pragma unroll
                    for (int j = 0; j < S::MAXJ; ++j) {
                        if (j >= nj) break;
                        cp_async<sizeof(r2)>(spr + (nt + j) * BLOCK, C.bc_pr + ((uint32_t)j * (uint32_t)cpad + c));
                    }
                }
            }
        }
        cp_async_commit();
    };
    auto prefetch_epi = [&](int tile) {          // always commits a (possibly empty) group
        const int i = tile * BLOCK + tid;
        if (a.l2_ring && i < seg.ncol) {
            // TruncatedADAGrad with the ring slot not staged (shared-memory budget): pull it into L2 now, the
            // update epilogue reads it after the K samples
            const uint32_t c = (uint32_t)(seg.col0 + i);
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(C.lam_ring + ((uint32_t)t * (uint32_t)cpad + c)));
            }
            if (!seg.neutral) {
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(C.bc_ring + ((uint32_t)j * (uint32_t)cpad + c)));
                }
            }
        }
        if (nac && i < seg.ncol) {
            const uint32_t c = (uint32_t)(seg.col0 + i);
            r2 *sac = reinterpret_cast<r2 *>(epi0) + tid;
            r2 *srg = reinterpret_cast<r2 *>(epi0 + th_bytes) + tid;
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                cp_async<sizeof(r2)>(sac + t * BLOCK, C.lam_acc + ((uint32_t)t * (uint32_t)cpad + c));
            }
            if (!seg.neutral) {
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    cp_async<sizeof(r2)>(sac + (nt + j) * BLOCK, C.bc_acc + ((uint32_t)j * (uint32_t)cpad + c));
                }
            }
            if (nrg) {
#pragma unroll
                for (int t = 0; t < S::MAXT; ++t) {
                    if (t >= nt) break;
                    cp_async<sizeof(r2)>(srg + t * BLOCK, C.lam_ring + ((uint32_t)t * (uint32_t)cpad + c));
                }
                if (!seg.neutral) {
#pragma unroll
                    for (int j = 0; j < S::MAXJ; ++j) {
                        if (j >= nj) break;
                        cp_async<sizeof(r2)>(srg + (nt + j) * BLOCK, C.bc_ring + ((uint32_t)j * (uint32_t)cpad + c));
                    }
                }
            }
        }
        cp_async_commit();
    };

    prefetch(blockIdx.x - seg.blk0, 0);
    // launched as a programmatic dependent of the tail kernel (fused step): everything above ran beside it,
    // its output (the context) is needed from here on.  A no-op for ordinary launches.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.abort && *reinterpret_cast<const volatile int *>(a.abort)) { cp_async_wait<0>(); return; }
    for (int i = tid; i < a.K * 3 * TT; i += BLOCK) {
        const int k = i / (3 * TT), r = i - k * 3 * TT, j = r / TT, t = r - j * TT;
        sctx[k * CS + j * TT + t] = a.ctx[(((size_t)seg.rep * a.K + k) * 3 + j) * a.tmax_ctx + t];
    }
    if (want_elbo)
        for (int k = 0; k <= a.K; ++k) sel[k * BLOCK + tid] = 0.0;
    cp_async_wait<1>();                 // the direction table (the first tile may still be in flight)
    __syncthreads();

    // ratio terms per environment (the "-1" of d/dlog-sigma and the -log-sigma of the density)
    int n_of_e[S::MAXE];
#pragma unroll
    for (int e = 0; e < S::MAXE; ++e) n_of_e[e] = 0;
    if (NE == 1) n_of_e[0] = nt - 1;
    else for (int t = 1; t < nt; ++t) n_of_e[a.env_of_t[t]] += 1;

    int buf = 0;
    for (int tile = blockIdx.x - seg.blk0; tile < ntile; tile += nblk, buf ^= (nbuf - 1)) {
        // in flight, oldest first: {theta(tile)} -> + {acc(tile)} + {theta(next tile)}
        if (nbuf == 2) { prefetch_epi(tile); prefetch(tile + nblk, buf ^ 1); cp_async_wait<2>(); }
        else { if (tile != (int)blockIdx.x - seg.blk0) prefetch(tile, 0); prefetch_epi(tile); cp_async_wait<0>(); }
        const int i = tile * BLOCK + tid;
        if (i >= seg.ncol) continue;
        const int c = seg.col0 + i;
        const uint32_t colid = C.col_id ? C.col_id[c] : seg.colid0 + (uint32_t)i;
        unsigned char *base = stage0 + (size_t)buf * buf_bytes;
        const r2 *sth = reinterpret_cast<const r2 *>(base) + tid;
        const r2 *spr = reinterpret_cast<const r2 *>(base + th_bytes) + tid;
        const int *scn = reinterpret_cast<const int *>(base + (1 + npr) * th_bytes) + tid;
        const r2 *sac = reinterpret_cast<const r2 *>(epi0) + tid;
        const r2 *srg = reinterpret_cast<const r2 *>(epi0 + th_bytes) + tid;

        real mu[S::MAXT], sg[S::MAXT], sgr[S::MAXT], sge[S::MAXT], cnt[S::MAXT];
        double lsig_sum = 0.0;
#pragma unroll
        for (int t = 0; t < S::MAXT; ++t) {
            if (t >= nt) break;
            const r2 th = sth[t * BLOCK];
            mu[t] = th.x; sg[t] = softplus_only<real>(th.y);
            if constexpr (SUP) if (a.sup.z_direct) { mu[t] = real(0); sg[t] = real(1); }
            cnt[t] = (real)scn[t * BLOCK];
            sgr[t] = real(0); sge[t] = real(0);
            if (want_elbo) lsig_sum += (double)bb_log(sg[t]);
        }
        real mub[S::MAXJ], sgb[S::MAXJ], sgrb[S::MAXJ], sgeb[S::MAXJ];
        real hc[S::MAXE], hce[S::MAXE];
        int hbase = 0;
        if (!seg.neutral) {
#pragma unroll
            for (int j = 0; j < S::MAXJ; ++j) {
                if (j >= nj) break;
                const r2 th = sth[(nt + j) * BLOCK];
                mub[j] = th.x; sgb[j] = softplus_only<real>(th.y);
                if constexpr (SUP) if (a.sup.z_direct) { mub[j] = real(0); sgb[j] = real(1); }
                sgrb[j] = real(0); sgeb[j] = real(0);
                if (want_elbo) lsig_sum += (double)bb_log(sgb[j]);
            }
            if constexpr (HIER) {
                hbase = C.hgroup[c];
#pragma unroll
                for (int e = 0; e < S::MAXE; ++e) { hc[e] = real(0); hce[e] = real(0); }
            }
        }
        const int nclass = seg.neutral ? nt : nt + nj;

        // the K samples; instantiated twice so the common vector-prior case carries no per-latent prior loads
        auto sample_loop = [&](auto matpr_tag) {
        constexpr bool MATPR = decltype(matpr_tag)::value;
        // hierarchical models: (z, eps) of the hyper latent for sample k + 1 is loaded while sample k is computed
        r2 hz_nxt[S::MAXE];
        auto load_hz = [&](int k) {
            if constexpr (HIER) {
                if (!seg.neutral) {
#pragma unroll
                    for (int e = 0; e < S::MAXE; ++e) {
                        if (e >= ne) break;
                        hz_nxt[e] = a.hy_zeps[(size_t)k * a.hz_k + (size_t)(hbase + e) * a.hz_h];
                    }
                }
            }
        };
        load_hz(0);
#pragma unroll(FUSE ? kFuseUnroll : kP2Unroll)
        for (int k = 0; k < a.K; ++k) {
            r2 hz_cur[S::MAXE];
            if constexpr (HIER) {
#pragma unroll
                for (int e = 0; e < S::MAXE; ++e) hz_cur[e] = hz_nxt[e];
                if (k + 1 < a.K) load_hz(k + 1);
            }
            real eps[S::MAXC];
            column_noise<real, S::MAXC, SUP>(eps, nclass, nt, colid, (uint32_t)k, a.step, a.key, strig, a.sup,
                                             c, cpad, C.tmax, C.nj);
            const real *crow = sctx + (size_t)k * CS;
            real cv[CSC];
            if constexpr (NT > 0) {
                using vw = typename std::conditional<std::is_same<real, float>::value, float4, double2>::type;
#pragma unroll
                for (int q = 0; q < CSC / VW; ++q) {
                    const vw v = reinterpret_cast<const vw *>(crow)[q];
                    if constexpr (VW == 4) { cv[4 * q] = v.x; cv[4 * q + 1] = v.y; cv[4 * q + 2] = v.z; cv[4 * q + 3] = v.w; }
                    else { cv[2 * q] = v.x; cv[2 * q + 1] = v.y; }
                }
            }
            auto cE = [&](int t) -> real { if constexpr (NT > 0) return cv[t]; else return crow[t]; };            // c_t - sbar_t
            auto cG = [&](int t) -> real { if constexpr (NT > 0) return cv[NT + t]; else return crow[TT + t]; };  // (U_{t-1} - U_t) / Lambda_t
            auto cW = [&](int t) -> real { if constexpr (NT > 0) return cv[2 * NT + t]; else return crow[2 * TT + t]; };   // exp(-2 logsigma-bar_t)
            real z[S::MAXT], g[S::MAXT];
            real lp = real(0);
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                z[t] = fma(sg[t], eps[t], mu[t]);
                const real lam = bb_exp(z[t]);
                r2 p = C.lam_pr_s;
                if constexpr (MATPR) if (lam_mat) p = npr ? spr[t * BLOCK] : C.lam_pr[(size_t)t * cpad + c];
                const real dz = z[t] - p.x;
                // Poisson (collapsed Poisson x Multinomial) + logLambda coupling + Normal prior
                g[t] = (cnt[t] - lam) + lam * cG(t) - dz * p.y;
                if (want_elbo) lp += cnt[t] * z[t] - lam - real(0.5) * dz * dz * p.y;
            }
            if (seg.neutral) {
                real uprev = real(0);
#pragma unroll
                for (int t = 0; t < S::MAXT - 1; ++t) {
                    if (t >= nt - 1) break;
                    real ce = cE(t), cw = cW(t);
                    if (a.aw_zs) {                 // as-written pairing: population latent p instead of t
                        const int p = (i * (nt - 1) + t) / a.aw_N;
                        const real *zsr = a.aw_zs + ((size_t)seg.rep * a.K + k) * a.tmax_ctx;
                        ce += zsr[t] - zsr[p];
                        cw = crow[2 * TT + p];
                    }
                    const real res = z[t + 1] - z[t] - ce;
                    const real u = cw * res;
                    g[t] += u - uprev;
                    uprev = u;
                }
                g[nt - 1] -= uprev;
            } else {
                real zb[S::MAXJ], zs[S::MAXE], zl[S::MAXE], w[S::MAXE], gs[S::MAXE], gq[S::MAXE], extau[S::MAXE];
                real epsth[S::MAXE];
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    zb[j] = fma(sgb[j], eps[nt + j], mub[j]);
                }
#pragma unroll
                for (int e = 0; e < S::MAXE; ++e) {
                    if (e >= ne) break;
                    if constexpr (HIER) {
                        const r2 hz = hz_cur[e];
                        extau[e] = bb_exp(zb[3 * e + 1]);
                        zs[e] = fma(extau[e], zb[3 * e], hz.x);
                        epsth[e] = hz.y;
                        zl[e] = zb[3 * e + 2];
                    } else {
                        zs[e] = zb[2 * e];
                        zl[e] = zb[2 * e + 1];
                    }
                    w[e] = bb_exp(real(-2) * zl[e]);
                    gs[e] = real(0); gq[e] = real(0);
                }
                real uprev = real(0);
#pragma unroll
                for (int t = 0; t < S::MAXT - 1; ++t) {
                    if (t >= nt - 1) break;
                    const int e = NE == 1 ? 0 : a.env_of_t[t + 1];
                    const real res = z[t + 1] - z[t] - zs[e] - cE(t);
                    const real u = w[e] * res;
                    gs[e] += u;
                    gq[e] += u * res;
                    g[t] += u - uprev;
                    uprev = u;
                }
                g[nt - 1] -= uprev;
                real gb[S::MAXJ];
#pragma unroll
                for (int e = 0; e < S::MAXE; ++e) {
                    if (e >= ne) break;
                    const real gls = gq[e] - (real)n_of_e[e];
                    if constexpr (HIER) {
                        gb[3 * e] = gs[e] * extau[e];
                        gb[3 * e + 1] = gs[e] * extau[e] * zb[3 * e];
                        gb[3 * e + 2] = gls;
                        hc[e] += gs[e];
                        hce[e] += gs[e] * epsth[e];
                        if constexpr (SUP) if (a.sup.dump_hcontrib)
                            a.sup.dump_hcontrib[((size_t)k * ne + e) * cpad + c] = gs[e];
                    } else {
                        gb[2 * e] = gs[e];
                        gb[2 * e + 1] = gls;
                    }
                    if (want_elbo) lp += -(real)n_of_e[e] * zl[e] - real(0.5) * gq[e];
                }
#pragma unroll
                for (int j = 0; j < S::MAXJ; ++j) {
                    if (j >= nj) break;
                    r2 p = C.bc_pr_s[j % S::PER];
                    if constexpr (MATPR) if (bc_mat) p = npr ? spr[(nt + j) * BLOCK] : C.bc_pr[(size_t)j * cpad + c];
                    const real dz = zb[j] - p.x;
                    gb[j] -= dz * p.y;
                    if (want_elbo) lp -= real(0.5) * dz * dz * p.y;
                    sgrb[j] += gb[j];
                    sgeb[j] = fma(gb[j], eps[nt + j], sgeb[j]);
                    if constexpr (SUP) if (a.sup.dump_bc) a.sup.dump_bc[((size_t)k * C.nj + j) * cpad + c] = gb[j];
                }
            }
#pragma unroll
            for (int t = 0; t < S::MAXT; ++t) {
                if (t >= nt) break;
                sgr[t] += g[t];
                sge[t] = fma(g[t], eps[t], sge[t]);
                if constexpr (SUP) if (a.sup.dump_lam) a.sup.dump_lam[((size_t)k * C.tmax + t) * cpad + c] = g[t];
            }
            if (want_elbo) sel[k * BLOCK + tid] += (double)lp;
        }
        };
        if (lam_mat || bc_mat) sample_loop(std::true_type{});
        else sample_loop(std::false_type{});
        if (want_elbo) sel[a.K * BLOCK + tid] += lsig_sum;
        if (nbuf == 2) cp_async_wait<1>();      // this tile's accumulators (the next tile's theta may still fly)

        // fused optimiser update of every latent of the column (theta / accumulators re-read from the
        // stage); the mode is kernel-uniform, so branch once around straight-line per-latent code
        auto finish_all = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            finish_batch<real, MODE, S::MAXT, FUSE>(a.opt, invK, nt, sgr, sge, mu, sg, sth, nac ? sac : nullptr,
                                                    nrg ? srg : nullptr, C.lam_th + c, C.lam_acc + c,
                                                    C.lam_ring + c, a.gout_lam + c, (size_t)cpad);
            if (!seg.neutral)
                finish_batch<real, MODE, S::MAXJ, FUSE>(a.opt, invK, nj, sgrb, sgeb, mub, sgb, sth + nt * BLOCK,
                                                        nac ? sac + nt * BLOCK : nullptr,
                                                        nrg ? srg + nt * BLOCK : nullptr, C.bc_th + c, C.bc_acc + c,
                                                        C.bc_ring + c, a.gout_bc + c, (size_t)cpad);
        };
        if (a.opt.update) {
            if (a.opt.kind == 1) finish_all(std::integral_constant<int, 0>{});
            else finish_all(std::integral_constant<int, 1>{});
        } else if (a.gout_lam) {
            finish_all(std::integral_constant<int, 2>{});
        }
        if (!seg.neutral) {
            if constexpr (HIER) {
#pragma unroll
                for (int e = 0; e < S::MAXE; ++e) {
                    if (e >= ne) break;
                    a.hcontrib[(size_t)e * cpad + c] = mk2<real>(hc[e], hce[e]);
                }
            }
        }
        if constexpr (FUSE && !HIER)
            if (!seg.neutral)
                pass1_column_direct<real, NT, NE>(mu, sg, mub, sgb, false, nt, ne, 0, a.K, colid, a.step + 1, a.key,
                                                  strig, a.env_of_t, facc + (PAIRS ? 2 * tid : tid), pva);
    }
    cp_async_wait<0>();
    if constexpr (FUSE && !HIER) {
        auto flush = [&](int k0, int k1) {          // block reduction of the private columns -> part
            __syncthreads();
            flush_rows<real, PAIRS>(facc, k1 - k0, pvs, pva, k0, a.pv, a.part);
            __syncthreads();
        };
        if (!seg.neutral) {
            flush(0, a.K);
        } else {
            const int kchunk = max(1, min(a.K, a.acc_slots / pva));
            for (int kc0 = 0; kc0 < a.K; kc0 += kchunk) {
                const int kc1 = min(a.K, kc0 + kchunk);
                if (kc0 > 0) {
                    for (int i = tid; i < (kc1 - kc0) * pva * BLOCK; i += BLOCK) facc[i] = real(0);
                    __syncthreads();
                }
                for (int tile = blockIdx.x - seg.blk0; tile < ntile; tile += nblk) {
                    const int i = tile * BLOCK + tid;
                    if (i >= seg.ncol) continue;
                    const int c = seg.col0 + i;
                    const uint32_t colid = C.col_id ? C.col_id[c] : seg.colid0 + (uint32_t)i;
                    real mu[S::MAXT], sg[S::MAXT];
#pragma unroll
                    for (int t = 0; t < S::MAXT; ++t) {
                        if (t >= nt) break;
                        const r2 th = C.lam_th[(size_t)t * cpad + c];      // this thread's own update, already written
                        mu[t] = th.x; sg[t] = softplus_only<real>(th.y);
                    }
                    pass1_column_direct<real, NT, NE>(mu, sg, mu, sg, true, nt, ne, kc0, kc1, colid, a.step + 1,
                                                      a.key, strig, a.env_of_t, facc + (PAIRS ? 2 * tid : tid), pva);
                }
                flush(kc0, kc1);
            }
        }
    }
    if (want_elbo) {
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        for (int row = warp; row <= a.K; row += BLOCK / 32) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < BLOCK / 32; ++j) s += sel[row * BLOCK + j * 32 + lane];
            s = warp_sum<double>(s);
            if (lane == 0) a.epart[(size_t)blockIdx.x * (a.K + 1) + row] = s;
        }
    }
}

}  // namespace bb
