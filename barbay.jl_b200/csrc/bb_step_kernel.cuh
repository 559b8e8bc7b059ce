// The fused ADVI step kernel of the non-hierarchical models (fitness_normal, multienv_fitness_normal;
// src/model_fitness_normal.jl:131-271, src/model_multienv_fitness_normal.jl:145-302 of the reference), sm_100a.
//
// Same algebra and software pipeline as pass2_kernel<..., FUSE> (bb_kernels.cuh): per
// column, regenerate the noise of step i, evaluate the gradient of the log-joint for all K samples, apply the
// AdaGrad-family update to every (mu, omega) of the column, then -- with the fresh theta and the noise of step
// i + 1 -- accumulate the partial sums the next step's shared phase needs.  What is new:
//
//  * the K samples are walked W = 2 at a time on packed fp32 arithmetic (bb_pack.cuh: FFMA2 / FADD2 / FMUL2),
//    the prior and count terms folded into per-column constants: the step is issue-bound, not HBM-bound, at K = 8;
//  * tiles are staged by bulk async copies (cp.async.bulk = TMA 1-D, SASS UBLKCP) completed on mbarriers:
//    one elected thread moves each contiguous SoA row of the tile (128 columns x 8 B = 1 KB) instead of 128
//    threads issuing one 8-byte LDGSTS each;
//  * PERSISTENT mode (nsteps > 1, cooperative launch): the kernel loops over ADVI steps.  Between steps the
//    CTAs reduce their partial sums through two levels of tickets, the last one posts the rank's sums to every
//    peer GPU over NVLink peer memory, and EVERY CTA then completes the sums, evaluates the shared-latent
//    phases redundantly in its own shared memory (c_t, U_t, G_t, the population-latent update) and carries on:
//    no kernel boundary and no second grid-wide hand-off per step.  This is the small-shard path of the
//    8-GPU strong-scaling configuration, where a step is tens of microseconds.
#pragma once
#include "bb_aux_kernels.cuh"
#include "bb_pack.cuh"

namespace bb {

#ifndef BB_STEP_MIN_BLOCKS
#define BB_STEP_MIN_BLOCKS 3
#endif
#ifndef BB_STEP_UNROLL2
#define BB_STEP_UNROLL2 1       // packs of pass 2 interleaved per thread (ILP: the step is latency-bound on small shards)
#endif
#ifndef BB_STEP_UNROLL1
#define BB_STEP_UNROLL1 1       // same for the pass-1 loop
#endif
constexpr int kStepUnroll2 = BB_STEP_UNROLL2, kStepUnroll1 = BB_STEP_UNROLL1;

// ---------------------------------------------------------------- mbarrier / bulk-copy primitives
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    // a bulk copy that never lands (it cannot, short of a programming error) must not hang the GPU: trap after ~2 s
    const long long t0 = clock64();
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) asm volatile("trap;");
    } while (!ok);
}
// global -> shared bulk copy (TMA 1-D); 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// fixed-order sums of `n` doubles `stride` apart, two outputs at a time: the loads of a batch (2 x B) are in flight
// together -- one L2 round trip (0.55 us between the last arrival and the next step, r2_tunechain.log) per batch
// instead of one per element -- and the additions keep the element order, so the result does not depend on the batching
template <int B>
__device__ __forceinline__ void ordered_sum2_cg(const double *p0, const double *p1, int n, size_t stride, double &s0, double &s1) {
    s0 = 0.0; s1 = 0.0;
    for (int i = 0; i < n; i += B) {
        double v0[B], v1[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const size_t o = (size_t)min(i + u, n - 1) * stride;
            v0[u] = __ldcg(p0 + o); v1[u] = __ldcg(p1 + o);
        }
#pragma unroll
        for (int u = 0; u < B; ++u)
            if (i + u < n) { s0 += v0[u]; s1 += v1[u]; }
    }
}

// ---------------------------------------------------------------- arguments
struct StepSync {                  // global control block of the persistent mode (zeroed once, tickets are monotone)
    unsigned group_ticket[64];
    unsigned final_ticket;
    int err;                       // 1: a peer never posted its sums (bounded spin) -- the launch stops stepping
    int steps_done;                // steps completed by the launch that raised err (diagnostic)
    int pad;
    unsigned long long stat[16];   // CTA 0, summed over the in-kernel tails: cycles {column phase, arrive -> sums, sums -> context}, tails
};

template <typename real> struct StepArgs {
    SegList segs;
    ColArrays<real> cols;          // persistent + TruncatedADAGrad: lam_ring / bc_ring are the ring BASES (slot 0)
    int K, P;                      // P = K * NQ * tmax: length of one partial-sum vector
    int env_of_t[MAX_NT_DYN];
    PhiloxKey key;
    uint32_t step;                 // first step of this launch
    int nsteps;                    // 1: plain fused step (launched behind tail_kernel); > 1: persistent, cooperative
    const real *ctx;               // [K][3][tmax] context of step `step` (written by tail_kernel)
    OptArgsT<real> opt;
    int stage_pr, stage_ring, l2_ring;
    int nbuf;                      // 2: the next tile is prefetched while this one is computed; 1: single staging buffer
    int stage_acc;                 // 1: accumulators (and the ring slot) are staged for the epilogue; 0: read from global (L2-prefetched)
    int acc_rows;                  // budget of the pass-1 accumulators: rows of BLOCK x (2 slots x W samples)
    int tail_scratch;              // doubles of shared_body's working arrays (the in-kernel tail borrows the accumulators)
    double *xpart;                 // [gridDim.x][P] block partial sums of the next step, output space [k][q][t]
    int ring_n, ring_slot;         // TruncatedADAGrad: step s uses slot (ring_slot + s - step) % ring_n
    // ---- persistent mode
    StepSync *sync;
    double *gpart;                 // [ngroups][P]
    int gsize, ngroups;
    XchgPostArgs xp;               // peers' exchange buffers / flags (world == 1: the local one); seq = first in-kernel exchange
    const double *xbuf;            // local exchange buffer [2][world][P]
    const unsigned long long *xflag;
    const int *abort;              // set by the tail kernel ahead of this launch when its exchange failed -> return
    SharedArgs<real> sa;           // sh_th / sh_acc / sh_pr global; sh_ring_rd = sh_ring_wr = ring BASE; eps_sh = precomputed noise [nsteps][K][2 nst]
};

// output-space index of accumulator slot v of a column population (see bb_kernels.cuh "Slots per sample")
// neutral: [0,T) Lambda | [T,2T-1) sum d | [2T-1,3T-2) sum d^2 ; mutant: [0,T) Lambda | [T,2T-1) A | W: 2T-1 (E == 1) or [2T-1,3T-2)
template <int NT, int NE>
__device__ __forceinline__ void slot_to_out(bool neutral, int v, int &q, int &t, int &rep) {
    rep = 1;
    if (v < NT) { q = Q_LAM; t = v; return; }
    if (v < 2 * NT - 1) { q = neutral ? Q_DN : Q_A; t = v - NT; return; }
    t = v - (2 * NT - 1);
    if (neutral) { q = Q_D2N; return; }
    q = Q_W;
    if (NE == 1) rep = NT - 1;      // one W for every time point
}

// ---------------------------------------------------------------- noise for a pack of samples
// Both passes over a column regenerate its draws from the lattice.  Storing them between the passes was built and
// measured in round 2 (binary16 draws, 16 B per column and sample): 167 us against 142 us for regenerating -- the
// second pass then waits on its own loads -- see DESIGN.md section 5.
template <typename real, int W, int MAXC>
__device__ __forceinline__ void column_noise_pack(Pack<real, W> (&eps)[MAXC], int nclass, uint32_t colid, uint32_t k0,
                                                  uint32_t step, const PhiloxKey &key, const float2 *tab) {
#pragma unroll
    for (int q = 0; q < MAXC / 8; ++q) {
        if (q * 8 >= nclass) break;
        if constexpr (W == 1) {
            real n[8];
            normals8<real>(colid, (STREAM_COLUMN << 24) | (uint32_t)q, k0, step, key, tab, n);
#pragma unroll
            for (int l = 0; l < 8; ++l) eps[8 * q + l].v = n[l];
        } else if constexpr (std::is_same<real, float>::value) {
            uint32_t xa[4], xb[4];
            philox4x32(colid, (STREAM_COLUMN << 24) | (uint32_t)q, k0, step, key, xa);
            philox4x32(colid, (STREAM_COLUMN << 24) | (uint32_t)q, k0 + 1u, step, key, xb);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                // box_muller (bb_device.cuh) of both samples; the uniform's offset is one packed add
                const Pack<float, 2> u = pk_make(__uint_as_float(__funnelshift_r(xa[w], 0x100u, 10)),
                                                 __uint_as_float(__funnelshift_r(xb[w], 0x100u, 10))) +
                                         pk_bc<float, 2>(-1.99999988079071044921875f);
                const float sa = fast_sqrt(-fast_lg2(pk_get<0>(u))), sb = fast_sqrt(-fast_lg2(pk_get<1>(u)));
                const float2 da = tab[xa[w] & (TRIG_N - 1)], db = tab[xb[w] & (TRIG_N - 1)];
                eps[8 * q + 2 * w] = pk_make(sa * da.x, sb * db.x);
                eps[8 * q + 2 * w + 1] = pk_make(sa * da.y, sb * db.y);
            }
        } else {
            real na[8], nb[8];
            normals8<real>(colid, (STREAM_COLUMN << 24) | (uint32_t)q, k0, step, key, tab, na);
            normals8<real>(colid, (STREAM_COLUMN << 24) | (uint32_t)q, k0 + 1u, step, key, tab, nb);
#pragma unroll
            for (int l = 0; l < 8; ++l) eps[8 * q + l] = pk_make(na[l], nb[l]);
        }
    }
}

// two accumulator slots of one pack of samples, side by side: one 128-bit (fp32, W = 2) shared-memory access
template <typename P> struct __align__(2 * sizeof(P)) SlotPair { P a, b; };

template <typename real, int W> __device__ __forceinline__ Pack<real, W> pk_zero() { return pk_bc<real, W>(real(0)); }

// acc += (a, b).  fp32 packs use two 64-bit accesses: ptxas cannot keep two FADD2 results in one aligned register
// quad and would pay four moves per 128-bit store.
template <typename P> __device__ __forceinline__ void slot_add(SlotPair<P> *p, P a, P b) {
    SlotPair<P> s = *p;
    s.a = s.a + a; s.b = s.b + b;
    *p = s;
}
template <> __device__ __forceinline__ void slot_add<Pack<float, 2>>(SlotPair<Pack<float, 2>> *p, Pack<float, 2> a, Pack<float, 2> b) {
    const uint32_t addr = smem_u32(p);
    float2 x, y;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x.x), "=f"(x.y) : "r"(addr) : "memory");
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(y.x), "=f"(y.y) : "r"(addr) : "memory");
    x = __fadd2_rn(x, a.v);
    y = __fadd2_rn(y, b.v);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x.x), "f"(x.y) : "memory");
    asm volatile("st.shared.v2.f32 [%0+8], {%1, %2};" ::"r"(addr), "f"(y.x), "f"(y.y) : "memory");
}

// per-column environment tables of the multi-environment model: `sel` picks entry e of a register array with a
// uniform predicate chain (no dynamically indexed registers -> no local memory)
template <typename T, int N> __device__ __forceinline__ T sel(const T (&arr)[N], int e) {
    T r = arr[0];
#pragma unroll
    for (int i = 1; i < N; ++i) if (e == i) r = arr[i];
    return r;
}

// ---------------------------------------------------------------- pass 1 of one pack of samples
// z = mu + sigma eps for W samples at once and the column's contribution to every slot, added into the
// thread-private accumulators `acc` (offset to this pack and thread; rows of BLOCK SlotPairs).
template <typename real, int NT, int NE, int W>
__device__ __forceinline__ void pass1_pack(const Pack<real, W> *eps, const real *mu, const real *sg, const real *mub,
                                           const real *sgb, bool neutral, const int *env_of_t,
                                           SlotPair<Pack<real, W>> *acc) {
    using P = Pack<real, W>;
    constexpr int NV = (3 * NT - 2 + 1) & ~1;
    P v[NV];
    v[NV - 1] = pk_zero<real, W>();
    P z[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        z[t] = pk_fma(pk_bc<real, W>(sg[t]), eps[t], pk_bc<real, W>(mu[t]));
        v[t] = pk_exp(z[t]);
    }
    int np;
    if (neutral) {
#pragma unroll
        for (int t = 0; t < NT - 1; ++t) {
            const P d = z[t + 1] - z[t];
            v[NT + t] = d;
            v[2 * NT - 1 + t] = d * d;
        }
        np = (3 * NT - 1) / 2;
    } else {
        P zs[NE], w[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            zs[e] = pk_fma(pk_bc<real, W>(sgb[2 * e]), eps[NT + 2 * e], pk_bc<real, W>(mub[2 * e]));
            const P zl = pk_fma(pk_bc<real, W>(sgb[2 * e + 1]), eps[NT + 2 * e + 1], pk_bc<real, W>(mub[2 * e + 1]));
            w[e] = pk_exp_scaled(zl, real(-2));
        }
#pragma unroll
        for (int t = 0; t < NT - 1; ++t) {
            const int e = NE == 1 ? 0 : env_of_t[t + 1];
            const P we = sel(w, e), zse = sel(zs, e);
            v[NT + t] = we * ((z[t + 1] - z[t]) - zse);
            if (NE != 1) v[2 * NT - 1 + t] = we;
        }
        if (NE == 1) { v[2 * NT - 1] = w[0]; np = NT; }
        else np = (3 * NT - 1) / 2;
    }
#pragma unroll
    for (int q = 0; q < NV / 2; ++q) {
        if (q >= np) break;
        slot_add<P>(acc + q * BLOCK, v[2 * q], v[2 * q + 1]);
    }
}

// ---------------------------------------------------------------- block reduction of the accumulators
// facc: [npack][nqp][BLOCK] SlotPairs -> xpart[out index of (sample, slot)] in double, fixed order: a warp owns
// whole packs, sums the four threads-of-lane entries, and reduces the 2 W rows of a SlotPair with a transposed
// butterfly (bb_kernels.cuh flush_rows).
template <typename real, int NT, int NE, int W>
__device__ __forceinline__ void flush_packs(const SlotPair<Pack<real, W>> *facc, int npack, int pvs, int k0, bool neutral,
                                            int tmax, double *out) {
    using P = Pack<real, W>;
    constexpr int NW = BLOCK / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqp = (pvs + 1) >> 1;
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    for (int kp = warp; kp < npack; kp += NW) {
        for (int q = 0; q < nqp; ++q) {
            const SlotPair<P> *rp = facc + ((size_t)kp * nqp + q) * BLOCK;
            double s[4] = {0.0, 0.0, 0.0, 0.0};       // rows: (slot 2q, sample 0) (slot 2q, sample 1) (slot 2q+1, sample 0) (slot 2q+1, sample 1)
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const SlotPair<P> t = rp[j * 32 + lane];
                if constexpr (W == 2) {
                    s[0] += (double)pk_get<0>(t.a); s[1] += (double)pk_get<1>(t.a);
                    s[2] += (double)pk_get<0>(t.b); s[3] += (double)pk_get<1>(t.b);
                } else {
                    s[0] += (double)t.a.v; s[2] += (double)t.b.v;
                }
            }
            double kk;
            int row;
            if constexpr (W == 2) {
                double ka = hi16 ? s[2] : s[0], kb = hi16 ? s[3] : s[1];
                ka += __shfl_xor_sync(0xffffffffu, hi16 ? s[0] : s[2], 16);
                kb += __shfl_xor_sync(0xffffffffu, hi16 ? s[1] : s[3], 16);
                kk = hi8 ? kb : ka;
                kk += __shfl_xor_sync(0xffffffffu, hi8 ? ka : kb, 8);
                row = lane >> 3;                       // 0..3 = (slot parity << 1) | sample
            } else {
                kk = hi16 ? s[2] : s[0];
                kk += __shfl_xor_sync(0xffffffffu, hi16 ? s[0] : s[2], 16);
                kk += __shfl_xor_sync(0xffffffffu, kk, 8);
                row = (lane >> 4) << 1;
            }
            kk += __shfl_xor_sync(0xffffffffu, kk, 4);
            kk += __shfl_xor_sync(0xffffffffu, kk, 2);
            kk += __shfl_xor_sync(0xffffffffu, kk, 1);
            const int v = 2 * q + (row >> 1), k = k0 + kp * W + (row & 1);
            if ((lane & 7) == 0 && (W == 2 || (lane & 8) == 0) && v < pvs) {
                int oq, ot, rep;
                slot_to_out<NT, NE>(neutral, v, oq, ot, rep);
                for (int r = 0; r < rep; ++r) out[((size_t)k * NQ + oq) * tmax + ot + r] = kk;
            }
        }
    }
}

// ---------------------------------------------------------------- the kernel
// Shared memory: direction table (fp32) | mbarriers | packed context [K/W][3T] | persistent: ctx_lin, shared
// latents | 2 staging buffers {theta [T+J][BLOCK], [priors], counts [T][BLOCK]} | epilogue buffer {acc, [ring]} |
// pass-1 accumulators (aliased by the in-kernel tail's scratch between the flush and the next column phase)
template <typename real, int NT, int NE, int W>
__global__ void __launch_bounds__(BLOCK, BB_STEP_MIN_BLOCKS) step_kernel(const StepArgs<real> a) {
    using S = Shape<NT, NE, false>;
    using r2 = vec2<real>;
    using P = Pack<real, W>;
    using SP = SlotPair<P>;
    constexpr int NJ = 2 * NE, ROWS = NT + NJ;
    extern __shared__ __align__(128) unsigned char step_smem[];
    const int tid = threadIdx.x;
    const bool persist = a.nsteps > 1;

    // ---- shared-memory carve-up (every block agrees with the host: sizes depend on kernel-uniform values only)
    // the Box-Muller direction table is STATIC shared memory: its address is an immediate of the LDS
    constexpr bool F32 = std::is_same<real, float>::value;
    __shared__ float2 strig[F32 ? TRIG_N : 1];
    unsigned char *sp = step_smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sp); sp += 128;         // per warp: [0,1] staging buffers full, [2] epilogue buffer full
    constexpr int CSP = 3 * NT;                                           // context entries per pack
    P *sctx = reinterpret_cast<P *>(sp);
    const int npack = a.K / W;
    sp += ((size_t)npack * CSP * sizeof(P) + 127) / 128 * 128;
    double2 *s_sh_th = nullptr, *s_sh_acc = nullptr;
    real *s_sig = nullptr;                  // sigma [2n] of the population latents
    // thread j < K * 2n carries (z, eps) of population latent j % 2n, sample j / 2n, for the NEXT in-kernel tail in two
    // registers across the column phase (shared memory is the occupancy limiter: three CTAs fit with 288 bytes to spare)
    real my_z = real(0), my_eps = real(0);
    const int npr = a.stage_pr ? 1 : 0, nrg = a.stage_ring ? 1 : 0;
    constexpr size_t th_bytes = (size_t)ROWS * BLOCK * sizeof(r2), cn_bytes = (size_t)NT * BLOCK * sizeof(int);
    const size_t buf_bytes = (1 + npr) * th_bytes + cn_bytes;
    unsigned char *stage0 = sp; sp += (size_t)a.nbuf * buf_bytes;
    unsigned char *epi0 = sp; sp += a.stage_acc ? (size_t)(1 + nrg) * th_bytes : 0;
    SP *facc = reinterpret_cast<SP *>(sp);
    sp += (size_t)a.acc_rows * BLOCK * sizeof(SP);
    if (persist) {                          // the small persistent state sits behind the (128-byte aligned) big regions
        s_sh_th = reinterpret_cast<double2 *>(sp); sp += (size_t)2 * (NT - 1) * sizeof(double2);
        s_sh_acc = reinterpret_cast<double2 *>(sp); sp += (size_t)2 * (NT - 1) * sizeof(double2);
        s_sig = reinterpret_cast<real *>(sp);
    }

    const int sidx = find_segment(a.segs, blockIdx.x);
    const Seg seg = a.segs.seg[sidx];
    const ColArrays<real> &C = a.cols;
    const int cpad = C.cpad;
    const int nblk = seg.blk1 - seg.blk0;
    const int ntile = (seg.ncol + BLOCK - 1) / BLOCK;
    const int first = blockIdx.x - seg.blk0;
    const real invK = real(1) / real(a.K);
    const bool lam_mat = C.lam_pr != nullptr, bc_mat = C.bc_pr != nullptr;
    const bool neutral = seg.neutral != 0;
    const int pvs = neutral ? 3 * NT - 2 : (NE == 1 ? 2 * NT : 3 * NT - 2);
    const int nqp = (pvs + 1) >> 1;
    const int nrows = neutral ? NT : ROWS;                       // theta rows of this population

    const int warp = tid >> 5, lane = tid & 31;
    uint64_t *wbar = bars + 3 * warp;        // this warp's barriers: every warp runs its own copy pipeline, no block barrier
    if (tid < 3 * (BLOCK / 32)) {
        mbar_init(bars + tid, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (F32) {
#pragma unroll
        for (int i = 0; i < TRIG_N / BLOCK; ++i) strig[i * BLOCK + tid] = a.key.trig[i * BLOCK + tid];
    }
    for (int i = tid; i < a.acc_rows * BLOCK; i += BLOCK) facc[i] = SP{pk_zero<real, W>(), pk_zero<real, W>()};
    __syncthreads();

    // ---- bulk staging of one tile: rows are contiguous in the class-major SoA, so a row of the tile is ONE copy
    uint32_t ph_full[2] = {0u, 0u}, ph_epi = 0u;
    // One copy per (array row, warp): the warp's 32 columns of a row are 256 B (fp32 pairs) of contiguous global memory.
    // Lane l issues row l of the set (all rows of a tile in ONE instruction slot), lane 0 arms the barrier first.
    // Segments are padded to 32 columns, so a warp either has a full 32-column slice or none at all.
    auto warp_has = [&](int tile) { return tile < ntile && tile * BLOCK + warp * 32 < seg.ncol; };
    auto issue_pre = [&](int tile, int buf) {
        const uint32_t c = (uint32_t)(seg.col0 + tile * BLOCK + warp * 32);
        unsigned char *base = stage0 + (size_t)buf * buf_bytes;
        r2 *sth = reinterpret_cast<r2 *>(base) + warp * 32;
        r2 *spr = reinterpret_cast<r2 *>(base + th_bytes) + warp * 32;
        int *scn = reinterpret_cast<int *>(base + (1 + npr) * th_bytes) + warp * 32;
        constexpr uint32_t rb = 32 * sizeof(r2), cb = 32 * sizeof(int);
        const int n_pl = (npr && lam_mat) ? NT : 0, n_pb = (npr && bc_mat && !neutral) ? NJ : 0;
        if (lane == 0) mbar_expect_tx(wbar + buf, (uint32_t)(nrows + n_pl + n_pb) * rb + (uint32_t)NT * cb);
        __syncwarp();
        for (int row = lane; row < nrows + NT + n_pl + n_pb; row += 32) {
            if (row < NT) bulk_g2s(sth + row * BLOCK, C.lam_th + ((uint32_t)row * (uint32_t)cpad + c), rb, wbar + buf);
            else if (row < 2 * NT) bulk_g2s(scn + (row - NT) * BLOCK, C.cnt + ((uint32_t)(row - NT) * (uint32_t)cpad + c), cb, wbar + buf);
            else if (row < NT + nrows) bulk_g2s(sth + (row - NT) * BLOCK, C.bc_th + ((uint32_t)(row - 2 * NT) * (uint32_t)cpad + c), rb, wbar + buf);
            else if (row < NT + nrows + n_pl) {
                const int t = row - NT - nrows;
                bulk_g2s(spr + t * BLOCK, C.lam_pr + ((uint32_t)t * (uint32_t)cpad + c), rb, wbar + buf);
            } else {
                const int j = row - NT - nrows - n_pl;
                bulk_g2s(spr + (NT + j) * BLOCK, C.bc_pr + ((uint32_t)j * (uint32_t)cpad + c), rb, wbar + buf);
            }
        }
    };
    auto issue_epi = [&](int tile, const r2 *lam_ring, const r2 *bc_ring) {
        const uint32_t c = (uint32_t)(seg.col0 + tile * BLOCK + warp * 32);
        r2 *sac = reinterpret_cast<r2 *>(epi0) + warp * 32;
        r2 *srg = reinterpret_cast<r2 *>(epi0 + th_bytes) + warp * 32;
        constexpr uint32_t rb = 32 * sizeof(r2);
        if (lane == 0) mbar_expect_tx(wbar + 2, (uint32_t)(1 + nrg) * (uint32_t)nrows * rb);
        __syncwarp();
        for (int row = lane; row < (1 + nrg) * nrows; row += 32) {
            const int rr = row < nrows ? row : row - nrows;
            r2 *dst = (row < nrows ? sac : srg) + rr * BLOCK;
            const r2 *src = row < nrows ? (rr < NT ? C.lam_acc + ((uint32_t)rr * (uint32_t)cpad + c)
                                                   : C.bc_acc + ((uint32_t)(rr - NT) * (uint32_t)cpad + c))
                                        : (rr < NT ? lam_ring + ((uint32_t)rr * (uint32_t)cpad + c)
                                                   : bc_ring + ((uint32_t)(rr - NT) * (uint32_t)cpad + c));
            bulk_g2s(dst, src, rb, wbar + 2);
        }
    };

    // ratio terms per environment (the "-1" of d/dlog-sigma)
    int n_of_e[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) n_of_e[e] = 0;
    if (NE == 1) n_of_e[0] = NT - 1;
    else
        for (int t = 1; t < NT; ++t) {
#pragma unroll
            for (int e = 0; e < NE; ++e) if (a.env_of_t[t] == e) n_of_e[e] += 1;
        }

    // context of the first step: from tail_kernel (launched ahead of this kernel; programmatic dependent launch)
    if (warp_has(first)) issue_pre(first, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (a.abort && *reinterpret_cast<const volatile int *>(a.abort)) {
        if (warp_has(first)) mbar_wait(wbar + 0, 0u);      // the copies in flight must land before the CTA retires
        return;
    }
    auto pack_ctx = [&](const real *lin) {           // linear [K][3][NT] -> packs [K/W][3 NT]
        for (int i = tid; i < npack * CSP; i += BLOCK) {
            const int kp = i / CSP, j = i - kp * CSP;
            if constexpr (W == 2) sctx[i] = pk_make(lin[(size_t)(2 * kp) * CSP + j], lin[(size_t)(2 * kp + 1) * CSP + j]);
            else sctx[i].v = lin[(size_t)kp * CSP + j];
        }
    };
    pack_ctx(a.ctx);
    constexpr int N2 = 2 * (NT - 1);
    if (persist) {
        // the population latents live in shared memory for the whole launch (every CTA keeps the same replica);
        // z = mu + sigma eps of the first in-kernel tail (step + 1) is drawn here, later ones at the end of each tail
        for (int i = tid; i < N2; i += BLOCK) { s_sh_th[i] = a.sa.sh_th[i]; s_sh_acc[i] = a.sa.sh_acc[i]; }
        for (int j = tid; j < a.K * N2; j += BLOCK) {
            const int i = j % N2;
            const double2 th = a.sa.sh_th[i];
            const real sg = softplus_only<real>((real)th.y);
            const real e = (real)a.sa.eps_sh[(size_t)a.K * N2 + j];
            my_eps = e; my_z = fma(sg, e, (real)th.x);          // K * 2n <= BLOCK (checked by the host)
            if (j < N2) s_sig[i] = sg;
        }
    }
    __syncthreads();

    long long st_col = 0, st_wait = 0, st_ctx = 0, st_a = 0, st_b = 0;
    int abort_flag = 0;
    for (int si = 0; si < a.nsteps; ++si) {
        const uint32_t step = a.step + (uint32_t)si;
        const long long tc0 = clock64();
        // this step's TruncatedADAGrad ring slot
        const int rslot = a.ring_n > 0 ? (a.ring_slot + si) % a.ring_n : 0;
        r2 *lam_ring = C.lam_ring ? C.lam_ring + (size_t)rslot * C.tmax * cpad : nullptr;
        r2 *bc_ring = C.bc_ring ? C.bc_ring + (size_t)rslot * C.nj * cpad : nullptr;
        if (si > 0) {
            // (persistent) the buffers are free and this CTA's theta / accumulator stores of the last step are ordered
            // before the bulk copies below by the fences of the in-kernel tail
            if (warp_has(first)) issue_pre(first, 0);
        }
        int buf = 0;
        for (int tile = first; tile < ntile; tile += nblk, buf ^= (a.nbuf - 1)) {
            if (!warp_has(tile)) continue;   // (warp-uniform) a partial last tile leaves whole warps without columns
            __syncwarp();                    // the warp is done with the previous tile's buffers
            if (a.stage_acc) issue_epi(tile, lam_ring, bc_ring);
            if (a.nbuf == 2 && warp_has(tile + nblk)) issue_pre(tile + nblk, buf ^ 1);
            const int i = tile * BLOCK + tid;
            const bool active = i < seg.ncol;
            const int c = seg.col0 + (active ? i : 0);
            if (!a.stage_acc && active) {    // the epilogue reads the accumulators from global memory: have them in L2 by then
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(C.lam_acc + ((uint32_t)t * (uint32_t)cpad + c)));
                if (!neutral) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(C.bc_acc + ((uint32_t)j * (uint32_t)cpad + c)));
                }
            }
            if (a.l2_ring && active) {
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(lam_ring + ((uint32_t)t * (uint32_t)cpad + c)));
                if (!neutral) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(bc_ring + ((uint32_t)j * (uint32_t)cpad + c)));
                }
            }
            mbar_wait(wbar + buf, ph_full[buf]); ph_full[buf] ^= 1u;
            if (active) {
                const uint32_t colid = C.col_id ? C.col_id[c] : seg.colid0 + (uint32_t)i;
                unsigned char *base = stage0 + (size_t)buf * buf_bytes;
                const r2 *sth = reinterpret_cast<const r2 *>(base) + tid;
                const r2 *spr = reinterpret_cast<const r2 *>(base + th_bytes) + tid;
                const int *scn = reinterpret_cast<const int *>(base + (1 + npr) * th_bytes) + tid;
                const r2 *sac = reinterpret_cast<const r2 *>(epi0) + tid;
                const r2 *srg = reinterpret_cast<const r2 *>(epi0 + th_bytes) + tid;

                // per-column constants: g_t = c0_t + lam (G_t - 1) + z_t npy_t with c0 = r + m / s^2, npy = -1 / s^2
                real mu[NT], sg[NT], c0[NT], npy[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const r2 th = sth[t * BLOCK];
                    mu[t] = th.x; sg[t] = softplus_only<real>(th.y);
                    r2 p = C.lam_pr_s;
                    if (lam_mat) p = npr ? spr[t * BLOCK] : C.lam_pr[(size_t)t * cpad + c];
                    c0[t] = fma(p.x, p.y, (real)scn[t * BLOCK]);
                    npy[t] = -p.y;
                }
                real mub[NJ], sgb[NJ], cb0[NJ], npb[NJ];
                if (!neutral) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const r2 th = sth[(NT + j) * BLOCK];
                        mub[j] = th.x; sgb[j] = softplus_only<real>(th.y);
                        r2 p = C.bc_pr_s[j & 1];
                        if (bc_mat) p = npr ? spr[(NT + j) * BLOCK] : C.bc_pr[(size_t)j * cpad + c];
                        // s: + m / s^2 ; log sigma: + m / s^2 - (number of ratios of the environment)
                        cb0[j] = p.x * p.y - ((j & 1) ? (real)n_of_e[j >> 1] : real(0));
                        npb[j] = -p.y;
                    }
                }
                const int nclass = neutral ? NT : ROWS;

                P sgr[NT], sge[NT], sgrb[NJ], sgeb[NJ];
#pragma unroll
                for (int t = 0; t < NT; ++t) { sgr[t] = pk_zero<real, W>(); sge[t] = pk_zero<real, W>(); }
#pragma unroll
                for (int j = 0; j < NJ; ++j) { sgrb[j] = pk_zero<real, W>(); sgeb[j] = pk_zero<real, W>(); }

                // ---- the K samples, W at a time
#pragma unroll kStepUnroll2
                for (int kp = 0; kp < npack; ++kp) {
                    P eps[S::MAXC];
                    column_noise_pack<real, W, S::MAXC>(eps, nclass, colid, (uint32_t)(kp * W), step, a.key, strig);
                    const P *crow = sctx + (size_t)kp * CSP;       // {c_t - sbar_t | G_t | wbar_t}
                    P z[NT], g[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        z[t] = pk_fma(pk_bc<real, W>(sg[t]), eps[t], pk_bc<real, W>(mu[t]));
                        const P lam = pk_exp(z[t]);
                        // Poisson (collapsed Poisson x Multinomial) + logLambda coupling + Normal prior
                        g[t] = pk_fma(z[t], pk_bc<real, W>(npy[t]), pk_fma(lam, crow[NT + t], pk_bc<real, W>(c0[t]) - lam));
                    }
                    if (neutral) {
                        P uprev;
#pragma unroll
                        for (int t = 0; t < NT - 1; ++t) {
                            const P res = (z[t + 1] - z[t]) - crow[t];
                            const P u = crow[2 * NT + t] * res;
                            g[t] = t == 0 ? g[t] + u : g[t] + (u - uprev);
                            uprev = u;
                        }
                        g[NT - 1] = g[NT - 1] - uprev;
                    } else {
                        P zs[NE], zl[NE], w[NE], gs[NE], gq[NE];
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            zs[e] = pk_fma(pk_bc<real, W>(sgb[2 * e]), eps[NT + 2 * e], pk_bc<real, W>(mub[2 * e]));
                            zl[e] = pk_fma(pk_bc<real, W>(sgb[2 * e + 1]), eps[NT + 2 * e + 1], pk_bc<real, W>(mub[2 * e + 1]));
                            w[e] = pk_exp_scaled(zl[e], real(-2));
                            gs[e] = pk_zero<real, W>(); gq[e] = pk_zero<real, W>();
                        }
                        P uprev;
#pragma unroll
                        for (int t = 0; t < NT - 1; ++t) {
                            const int e = NE == 1 ? 0 : a.env_of_t[t + 1];
                            const P res = ((z[t + 1] - z[t]) - sel(zs, e)) - crow[t];
                            const P u = sel(w, e) * res;
#pragma unroll
                            for (int ee = 0; ee < NE; ++ee)
                                if (NE == 1 || e == ee) { gs[ee] = gs[ee] + u; gq[ee] = pk_fma(u, res, gq[ee]); }
                            g[t] = t == 0 ? g[t] + u : g[t] + (u - uprev);
                            uprev = u;
                        }
                        g[NT - 1] = g[NT - 1] - uprev;
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            const P gb0 = pk_fma(zs[e], pk_bc<real, W>(npb[2 * e]), gs[e] + pk_bc<real, W>(cb0[2 * e]));
                            const P gb1 = pk_fma(zl[e], pk_bc<real, W>(npb[2 * e + 1]), gq[e] + pk_bc<real, W>(cb0[2 * e + 1]));
                            sgrb[2 * e] = sgrb[2 * e] + gb0; sgeb[2 * e] = pk_fma(gb0, eps[NT + 2 * e], sgeb[2 * e]);
                            sgrb[2 * e + 1] = sgrb[2 * e + 1] + gb1; sgeb[2 * e + 1] = pk_fma(gb1, eps[NT + 2 * e + 1], sgeb[2 * e + 1]);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < NT; ++t) { sgr[t] = sgr[t] + g[t]; sge[t] = pk_fma(g[t], eps[t], sge[t]); }
                }

                // fold the packs: sum over all K samples
                real fgr[NT], fge[NT], fgrb[NJ], fgeb[NJ];
#pragma unroll
                for (int t = 0; t < NT; ++t) { fgr[t] = pk_hsum(sgr[t]); fge[t] = pk_hsum(sge[t]); }
#pragma unroll
                for (int j = 0; j < NJ; ++j) { fgrb[j] = pk_hsum(sgrb[j]); fgeb[j] = pk_hsum(sgeb[j]); }

                if (a.stage_acc) mbar_wait(wbar + 2, ph_epi);      // this tile's accumulators (and ring slot)
                // fused optimiser update of every latent of the column
                auto finish_all = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
                    finish_batch<real, MODE, NT, true>(a.opt, invK, NT, fgr, fge, mu, sg, sth, a.stage_acc ? sac : nullptr,
                                                       (nrg && a.stage_acc) ? srg : nullptr,
                                                       C.lam_th + c, C.lam_acc + c, lam_ring + c, nullptr, (size_t)cpad);
                    if (!neutral)
                        finish_batch<real, MODE, NJ, true>(a.opt, invK, NJ, fgrb, fgeb, mub, sgb, sth + NT * BLOCK,
                                                           a.stage_acc ? sac + NT * BLOCK : nullptr,
                                                           (nrg && a.stage_acc) ? srg + NT * BLOCK : nullptr, C.bc_th + c,
                                                           C.bc_acc + c, bc_ring + c, nullptr, (size_t)cpad);
                };
                if (a.opt.kind == 1) finish_all(std::integral_constant<int, 0>{});
                else finish_all(std::integral_constant<int, 1>{});

                // pass 1 of the next step with the fresh theta (mutant columns; neutral blocks sweep again below)
                if (!neutral) {
#pragma unroll kStepUnroll1
                    for (int kp = 0; kp < npack; ++kp) {
                        P eps[S::MAXC];
                        column_noise_pack<real, W, S::MAXC>(eps, nclass, colid, (uint32_t)(kp * W), step + 1u, a.key, strig);
                        pass1_pack<real, NT, NE, W>(eps, mu, sg, mub, sgb, false, a.env_of_t, facc + (size_t)kp * nqp * BLOCK + tid);
                    }
                }
            } else if (a.stage_acc) {
                mbar_wait(wbar + 2, ph_epi);
            }
            if (a.stage_acc) ph_epi ^= 1u;
            if (a.nbuf == 1 && warp_has(tile + nblk)) {       // single buffer: refill it once the whole warp has read this tile
                __syncwarp();
                issue_pre(tile + nblk, 0);
            }
        }

        // ---- block partial sums of the next step -> xpart[block][k][q][t]
        double *xout = a.xpart + (size_t)blockIdx.x * a.P;
        for (int i = tid; i < a.P; i += BLOCK) xout[i] = 0.0;
        if (!neutral) {
            __syncthreads();
            flush_packs<real, NT, NE, W>(facc, npack, pvs, 0, false, NT, xout);
            __syncthreads();
            for (int i = tid; i < npack * nqp * BLOCK; i += BLOCK) facc[i] = SP{pk_zero<real, W>(), pk_zero<real, W>()};
        } else {
            const int pchunk = max(1, min(npack, a.acc_rows / nqp));
            for (int kc0 = 0; kc0 < npack; kc0 += pchunk) {
                const int kc1 = min(npack, kc0 + pchunk);
                __syncthreads();
                for (int tile = first; tile < ntile; tile += nblk) {
                    const int i = tile * BLOCK + tid;
                    if (i >= seg.ncol) continue;
                    const int c = seg.col0 + i;
                    const uint32_t colid = C.col_id ? C.col_id[c] : seg.colid0 + (uint32_t)i;
                    real mu[NT], sg[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const r2 th = C.lam_th[(size_t)t * cpad + c];      // this thread's own update, already written
                        mu[t] = th.x; sg[t] = softplus_only<real>(th.y);
                    }
#pragma unroll 1
                    for (int kp = kc0; kp < kc1; ++kp) {
                        P eps[S::MAXC];
                        column_noise_pack<real, W, S::MAXC>(eps, NT, colid, (uint32_t)(kp * W), step + 1u, a.key, strig);
                        pass1_pack<real, NT, NE, W>(eps, mu, sg, mu, sg, true, a.env_of_t, facc + (size_t)(kp - kc0) * nqp * BLOCK + tid);
                    }
                }
                __syncthreads();
                flush_packs<real, NT, NE, W>(facc, kc1 - kc0, pvs, kc0 * W, true, NT, xout);
                __syncthreads();
                for (int i = tid; i < (kc1 - kc0) * nqp * BLOCK; i += BLOCK) facc[i] = SP{pk_zero<real, W>(), pk_zero<real, W>()};
            }
        }
        if (si + 1 == a.nsteps) break;

        // =================================================== in-kernel tail (persistent mode)
        // partial sums -> group -> rank -> all ranks; every CTA then runs the shared-latent phases itself
        const unsigned long long seq = a.xp.seq + (unsigned long long)si;
        const int parity = (int)(seq & 1ull);
        const int world = a.xp.world;
        int &s_flag = *reinterpret_cast<int *>(bars + 3 * (BLOCK / 32));     // spare slot of the barrier block
        __threadfence();
        fence_proxy_async();             // this CTA's theta / accumulator stores vs. the next step's bulk copies
        __syncthreads();
        const long long tc1 = clock64();
        const int grp = blockIdx.x / a.gsize;
        const int gsz = min(a.gsize, (int)gridDim.x - grp * a.gsize);
        // Measured on the chain of the last-arriving CTA (profiles/r2_tunechain.log, 1/8 of cfg2): group sum 2.2 us with
        // batches of 8 loads, tickets + fences 1.4, rank sum + stores 1.2, system fence + release stores 2.7 -- the release
        // stores alone 1.3 (they order the CTA's earlier stores themselves, through the barrier); one fence by thread 0
        // instead of one per thread: no difference; relaxed polling + one acquire fence: 3 us slower.
        if (tid == 0) {
            const unsigned t = atomicAdd(&a.sync->group_ticket[grp], 1u);
            s_flag = ((t + 1u) % (unsigned)gsz == 0u) ? 1 : 0;
        }
        __syncthreads();
        if (s_flag) {                    // last CTA of the group: sum the group's partials, fixed order
            __threadfence();
            const double *src = a.xpart + (size_t)grp * a.gsize * a.P;
            const long long tr0 = clock64();
            for (int j = tid; j < a.P; j += 2 * BLOCK) {       // two outputs per thread
                const int j2 = min(j + BLOCK, a.P - 1);
                double s0, s1;
                ordered_sum2_cg<16>(src + j, src + j2, gsz, (size_t)a.P, s0, s1);
                a.gpart[(size_t)grp * a.P + j] = s0;
                if (j + BLOCK < a.P) a.gpart[(size_t)grp * a.P + j + BLOCK] = s1;
            }
            const long long tr1 = clock64();
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned t = atomicAdd(&a.sync->final_ticket, 1u);
                s_flag = ((t + 1u) % (unsigned)a.ngroups == 0u) ? 2 : 0;
            }
            __syncthreads();
            if (s_flag == 2) {           // last group: the rank's sums go to every peer (world == 1: to the local buffer)
                __threadfence();
                const long long tr2 = clock64();
                for (int j = tid; j < a.P; j += 2 * BLOCK) {
                    const int j2 = min(j + BLOCK, a.P - 1);
                    double s0, s1;
                    ordered_sum2_cg<16>(a.gpart + j, a.gpart + j2, a.ngroups, (size_t)a.P, s0, s1);
                    const size_t o0 = xchg_off(a.P, parity * world + a.xp.rank, j), o1 = xchg_off(a.P, parity * world + a.xp.rank, j2);
                    for (int r = 0; r < world; ++r) {
                        a.xp.peer_buf[r][o0] = s0;
                        if (j + BLOCK < a.P) a.xp.peer_buf[r][o1] = s1;
                    }
                }
                const long long tr3 = clock64();
                __syncthreads();         // the release stores below are cumulative over the CTA's stores above
                if (tid < world) st_release_sys(a.xp.peer_flag[tid] + (parity * world + a.xp.rank), seq);
                if (tid == 0) {          // the last-arriving CTA of the grid: its reduction chain and its column phase
                    const long long tr4 = clock64();
                    atomicAdd(&a.sync->stat[6], (unsigned long long)(tr4 - tr0));
                    atomicAdd(&a.sync->stat[7], (unsigned long long)(tc1 - tc0));
                    atomicAdd(&a.sync->stat[8], (unsigned long long)(tr1 - tr0));      // group sum
                    atomicAdd(&a.sync->stat[9], (unsigned long long)(tr2 - tr1));      // fence, final ticket, fence
                    atomicAdd(&a.sync->stat[10], (unsigned long long)(tr3 - tr2));     // rank sum + stores to the peers
                    atomicAdd(&a.sync->stat[11], (unsigned long long)(tr4 - tr3));     // fence + flags
                }
            }
        }
        // fetched while the sums are on their way: priors and the evicted ring slot of population latent `tid`, the noise
        // of the tail after this one
        double2 pri_s = make_double2(0.0, 1.0), pri_l = pri_s, ring_old = make_double2(0.0, 0.0);
        double2 *ring_wr = nullptr;
        real eps_next = real(0);
        {
            const int t = tid % (NT - 1);
            pri_s = a.sa.sh_pr[t]; pri_l = a.sa.sh_pr[(NT - 1) + t];
            if (a.ring_n > 0 && tid < 2 * (NT - 1)) {          // shared-latent ring: n + 1 slots indexed by the step itself
                const uint32_t n1 = (uint32_t)a.ring_n + 1u;
                ring_old = __ldcg(a.sa.sh_ring_rd + (size_t)((step + 2u) % n1) * 2 * (NT - 1) + tid);
                ring_wr = a.sa.sh_ring_wr + (size_t)((step + 1u) % n1) * 2 * (NT - 1);
            }
            if (si + 2 < a.nsteps && tid < a.K * 2 * (NT - 1))
                eps_next = (real)a.sa.eps_sh[(size_t)(si + 2) * a.K * 2 * (NT - 1) + tid];
        }
        // every CTA: wait for all ranks' sums of this exchange
        if (tid < world) {
            const unsigned long long *f = a.xflag + (parity * world + tid);
            const long long t0 = clock64();
            while (ld_acquire_sys(f) != seq) {
                __nanosleep(40);         // hundreds of CTAs poll the same line: leave the L2 slice room for the flag's store
                if (clock64() - t0 > 4000000000LL) { a.sync->err = 1; a.sync->steps_done = si; abort_flag = 1; break; }   // ~2 s
                if (*reinterpret_cast<volatile int *>(&a.sync->err)) { abort_flag = 1; break; }
            }
        }
        abort_flag = __syncthreads_or(abort_flag);
        if (abort_flag) break;           // no update is applied with incomplete sums; the host reports the error
        const long long tc2 = clock64();
        // ---- the shared-latent phases, every CTA for itself (same arithmetic as shared_body, bb_aux_kernels.cuh, in the
        // kernel's own precision and with everything that does not depend on the sums -- noise, sigma, priors, the evicted
        // ring slot -- fetched before or cached across the wait): ~4 block barriers on the critical path
        double *tot = reinterpret_cast<double *>(facc);          // the accumulators are flushed: reuse as scratch
        real *w_lgl = reinterpret_cast<real *>(tot + a.P);       // [K][T] log Lambda
        real *w_il = w_lgl + a.K * NT;                           // [K][T] 1 / Lambda
        real *w_g = w_il + a.K * NT;                             // [K][2n] per-sample gradients of the shared latents
        real *w_u = w_g + a.K * N2;                              // [K][n] sum_all w res
        real *s_z = w_u + a.K * (NT - 1), *s_eps = s_z + a.K * N2;   // [K][2n] this tail's draws, from their owners' registers
        if (tid < a.K * N2) { s_z[tid] = my_z; s_eps[tid] = my_eps; }
        for (int j = tid; j < a.P; j += 2 * BLOCK) {             // all ranks' sums, rank order; both outputs' loads in flight together
            const int j2 = min(j + BLOCK, a.P - 1);
            double s0 = 0.0, s1 = 0.0;
            for (int r0 = 0; r0 < world; r0 += 8) {
                double v0[8], v1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = min(r0 + u, world - 1);
                    v0[u] = __ldcg(a.xbuf + xchg_off(a.P, parity * world + r, j));
                    v1[u] = __ldcg(a.xbuf + xchg_off(a.P, parity * world + r, j2));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (r0 + u < world) { s0 += v0[u]; s1 += v1[u]; }
            }
            tot[j] = s0;
            if (j + BLOCK < a.P) tot[j + BLOCK] = s1;
        }
        __syncthreads();
        const long long tca = clock64();
        for (int j = tid; j < a.K * NT; j += BLOCK) {
            const int k = j / NT, t = j - k * NT;
            const real lam = (real)tot[((size_t)k * NQ + Q_LAM) * NT + t];
            w_lgl[j] = bb_log(lam); w_il[j] = real(1) / lam;
        }
        __syncthreads();
        const real nneu = (real)a.sa.n_neutral;
        real *ctxw = reinterpret_cast<real *>(sctx);             // pack (kp, j) half h at ((kp * CSP + j) * W + h)
        for (int j = tid; j < a.K * (NT - 1); j += BLOCK) {
            const int k = j / (NT - 1), t = j - k * (NT - 1);
            const double *S = tot + (size_t)k * NQ * NT;
            const real zs = s_z[k * N2 + t], zl = s_z[k * N2 + (NT - 1) + t];
            const real c = w_lgl[k * NT + t + 1] - w_lgl[k * NT + t];
            const real wbar = bb_exp(real(-2) * zl);
            const real av = zs - c;
            const real dn = (real)S[Q_DN * NT + t], d2n = (real)S[Q_D2N * NT + t];
            const real am = (real)S[Q_A * NT + t], wm = (real)S[Q_W * NT + t];
            const real qn = d2n + real(2) * av * dn + nneu * av * av;
            const real u = wbar * (dn + nneu * av) + (am + av * wm);
            w_g[k * N2 + t] = -u - (zs - pri_s.x) * pri_s.y;
            w_g[k * N2 + (NT - 1) + t] = wbar * qn - nneu - (zl - pri_l.x) * pri_l.y;
            w_u[k * (NT - 1) + t] = u;
            const int kp = k / W, h = k - kp * W;
            ctxw[((size_t)kp * CSP + t) * W + h] = c - zs;
            ctxw[((size_t)kp * CSP + 2 * NT + t) * W + h] = wbar;
        }
        __syncthreads();
        for (int j = tid; j < a.K * NT; j += BLOCK) {
            const int k = j / NT, t = j - k * NT;
            const real up = t > 0 ? w_u[k * (NT - 1) + t - 1] : real(0), un = t < NT - 1 ? w_u[k * (NT - 1) + t] : real(0);
            const int kp = k / W, h = k - kp * W;
            ctxw[((size_t)kp * CSP + NT + t) * W + h] = (up - un) * w_il[j];
        }
        if (tid < N2) {                                          // gradient and optimiser update of population latent `tid`
            real sg = real(0), sge = real(0);
            for (int k = 0; k < a.K; ++k) { const real g = w_g[k * N2 + tid]; sg += g; sge = fma(g, s_eps[k * N2 + tid], sge); }
            double2 th = s_sh_th[tid], ac = s_sh_acc[tid];
            const real sigma = s_sig[tid], invK = real(1) / real(a.K);
            const real g0 = -(sg * invK), g1 = -((sge * invK + real(1) / sigma) * bb_exp((real)th.y - sigma));
            // the accumulators and the update itself in double, like shared_body (the tail kernel): the population
            // latents' first gradients are ~1e6 (they sum over every barcode), and TruncatedADAGrad's running window sum
            // in fp32 kept the cancellation residue of `sum - evicted` (~ulp(1e12)) for the rest of the run -- step
            // sizes of the log-sigma latents off by large factors from the fourth step on (tests/_debug_trunc.py)
            const double q0 = (double)g0 * (double)g0, q1 = (double)g1 * (double)g1;
            double d0, d1;
            if (a.sa.opt.kind == 1) {
                ac.x = fma(a.sa.opt.post, ac.x, a.sa.opt.tau * q0);
                ac.y = fma(a.sa.opt.post, ac.y, a.sa.opt.tau * q1);
                d0 = sqrt(ac.x) + 1e-8; d1 = sqrt(ac.y) + 1e-8;
            } else {
                ac.x = fmax(ac.x - ring_old.x, 0.0) + q0;
                ac.y = fmax(ac.y - ring_old.y, 0.0) + q1;
                if (blockIdx.x == 0) ring_wr[tid] = make_double2(q0, q1);
                d0 = a.sa.opt.tau + sqrt(ac.x) + 1e-8;
                d1 = a.sa.opt.tau + sqrt(ac.y) + 1e-8;
            }
            th.x = th.x - a.sa.opt.eta * (double)g0 / d0;
            th.y = th.y - a.sa.opt.eta * (double)g1 / d1;
            s_sh_th[tid] = th; s_sh_acc[tid] = ac;
            s_sig[tid] = softplus_only<real>((real)th.y);
        }
        __syncthreads();
        st_a += tca - tc2; st_b += clock64() - tca;
        // z = mu + sigma eps of the next in-kernel tail, from the noise fetched before the wait
        if (si + 2 < a.nsteps && tid < a.K * N2) {
            const int i = tid % N2;
            my_eps = eps_next; my_z = fma(s_sig[i], eps_next, (real)s_sh_th[i].x);
        }
        __syncthreads();
        // the scratch aliased the accumulators: zero them again for the next column phase
        for (int i = tid; i < a.acc_rows * BLOCK; i += BLOCK) facc[i] = SP{pk_zero<real, W>(), pk_zero<real, W>()};
        const long long tc3 = clock64();
        st_col += tc1 - tc0; st_wait += tc2 - tc1; st_ctx += tc3 - tc2;
    }
    if (persist) {
        if (blockIdx.x == 0) {
            // the shared latents after this launch's in-kernel tails, for the next launch (or any other entry point)
            for (int i = tid; i < 2 * (NT - 1); i += BLOCK) { a.sa.sh_th[i] = s_sh_th[i]; a.sa.sh_acc[i] = s_sh_acc[i]; }
        }
        if (blockIdx.x == gridDim.x - 1) {        // a block of the (large) mutant population reports the timings
            if (tid == 0) {
                atomicAdd(&a.sync->stat[4], (unsigned long long)st_a);
                atomicAdd(&a.sync->stat[5], (unsigned long long)st_b);
                atomicAdd(&a.sync->stat[0], (unsigned long long)st_col);
                atomicAdd(&a.sync->stat[1], (unsigned long long)st_wait);
                atomicAdd(&a.sync->stat[2], (unsigned long long)st_ctx);
                atomicAdd(&a.sync->stat[3], (unsigned long long)(a.nsteps - 1));
            }
        }
    }
}

template <typename real> using StepKernelFn = void (*)(const StepArgs<real>);

}  // namespace bb
