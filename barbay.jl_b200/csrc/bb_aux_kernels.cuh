// Small kernels around the column passes: partial-sum reduction, the shared
// (population-level) latents, the hyper latents of the hierarchical models,
// initialisation and the reference-order <-> device-layout permutations.
#pragma once
#include "bb_kernels.cuh"

namespace bb {

// sums layout: sums[((r * K + k) * 5 + q) * tmax + t]
//   q = 0 Lambda_t | 1 neutral sum d_t | 2 neutral sum d_t^2 | 3 mutant sum w (d_t - s) | 4 mutant sum w
enum { Q_LAM = 0, Q_DN = 1, Q_D2N = 2, Q_A = 3, Q_W = 4, NQ = 5 };

struct ReduceArgs {
    SegList segs;
    int K, tmax, pv, nt;     // pv / nt of this launch group
    int nblk;                // CTAs of the pass-1 launch (row length of part)
    int w_single;            // E == 1: mutant W stored once (slot 2nt-1)
    unsigned rep_mask;       // replicates belonging to this launch group (bit r)
    const double *part;      // [K][pv][nblk]
    const double *xpart;     // step kernel (bb_step_kernel.cuh): [nblk][K * NQ * tmax] in output space, or nullptr
    double *sums;
};

// one warp per output (r, k, q, t): fixed-order sum over the CTAs of the matching segments
__device__ __forceinline__ void reduce_rows(const ReduceArgs &a, int R) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nout = R * a.K * NQ * a.tmax;
    if (warp >= nout) return;
    const int t = warp % a.tmax, q = (warp / a.tmax) % NQ, k = (warp / (a.tmax * NQ)) % a.K,
              r = warp / (a.tmax * NQ * a.K);
    const int nt = a.nt;
    if (t >= (q == Q_LAM ? nt : nt - 1)) return;
    if (!((a.rep_mask >> r) & 1u)) return;      // replicate handled by another launch group (ragged T)
    double s = 0.0;
    if (a.xpart) {
        // the step kernel's blocks write output-space vectors (R == 1): plain fixed-order sum over the blocks
        const int P = a.K * NQ * a.tmax;
        for (int b = lane; b < a.nblk; b += 32) s += a.xpart[(size_t)b * P + warp];
        s = warp_sum<double>(s);
        if (lane == 0) a.sums[warp] = s;
        return;
    }
    for (int si = 0; si < a.segs.nseg; ++si) {
        const Seg &sg = a.segs.seg[si];
        if (sg.rep != r) continue;
        int v;
        if (q == Q_LAM) v = t;
        else if (q == Q_DN || q == Q_D2N) { if (!sg.neutral) continue; v = (q == Q_DN ? nt : 2 * nt - 1) + t; }
        else { if (sg.neutral) continue; v = q == Q_A ? nt + t : (a.w_single ? 2 * nt - 1 : 2 * nt - 1 + t); }
        const double *row = a.part + ((size_t)k * a.pv + v) * a.nblk;
        for (int b = sg.blk0 + lane; b < sg.blk1; b += 32) s += row[b];
    }
    s = warp_sum<double>(s);
    if (lane == 0) a.sums[((size_t)(r * a.K + k) * NQ + q) * a.tmax + t] = s;
}
static __global__ void __launch_bounds__(128) reduce_kernel(const ReduceArgs a, int R) { reduce_rows(a, R); }

// ------------------------------------------------------------------ peer-memory exchange of the step's sums
// One-shot all-reduce over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC):
// after its local reduction every rank stores its `P` partial sums into slot [parity][rank] of EVERY
// peer's exchange buffer and then raises flag [parity][rank] = seq on that peer.  The shared-latent
// kernel of each rank waits for the `world` flags of this exchange, adds the slots in rank order
// (deterministic and identical on all ranks) and carries on -- the collective is fused into the kernel
// that consumes it, nothing but two NVLink stores is on the critical path.  Two parities suffice: a
// rank cannot start exchange seq + 2 before every peer has finished reading seq (it needs their
// seq + 1 data first).
constexpr int MAX_WORLD = 16;

// Slot addressing of the exchange buffers: a slot's P doubles are laid out as 16-double (128-byte) lines 4 KB apart.
// Every CTA of the persistent step kernel reads all `world` slots at the same moment; dense slots (1.6 KB at cfg2) sit
// in one or two L2 slices and that read alone took 3.8 us of every step (r2 profile), spread lines are served by
// different slices in parallel.
constexpr int XCHG_LINE = 16, XCHG_STRIDE = 512;          // doubles per line, doubles between consecutive lines
__host__ __device__ inline size_t xchg_lines(int P) { return (size_t)(P + XCHG_LINE - 1) / XCHG_LINE; }
__host__ __device__ inline size_t xchg_slot_doubles(int P) { return xchg_lines(P) * XCHG_STRIDE; }
__host__ __device__ inline size_t xchg_off(int P, int slot, int j) {
    return ((size_t)slot * xchg_lines(P) + (size_t)(j >> 4)) * XCHG_STRIDE + (size_t)(j & 15);
}

struct XchgPostArgs {
    const double *sums;                       // this rank's partial sums [P]
    int P, world, rank, parity;
    unsigned long long seq;
    double *peer_buf[MAX_WORLD];              // peer r's exchange buffer [2][world][P] (own buffer for r == rank)
    unsigned long long *peer_flag[MAX_WORLD]; // peer r's flags [2][world]
};

static __global__ void __launch_bounds__(256) xchg_post_kernel(const XchgPostArgs a) {
    const int tid = threadIdx.x;
    for (int r = 0; r < a.world; ++r) {
        double *dst = a.peer_buf[r];
        for (int i = tid; i < a.P; i += blockDim.x) dst[xchg_off(a.P, a.parity * a.world + a.rank, i)] = a.sums[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) {
        unsigned long long *f = a.peer_flag[tid] + (a.parity * a.world + a.rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(a.seq) : "memory");
    }
}

struct XchgWaitArgs {
    const double *buf;                        // local exchange buffer [2][world][P], or nullptr (no exchange)
    const unsigned long long *flag;           // local flags [2][world]
    int P, world, parity;
    unsigned long long seq;
    int *err;                                 // set to 1 if a peer never showed up (bounded spin)
};

// executed by one CTA at the top of the consumer kernel; totals land in `sums`.  Returns false when a peer never
// posted (bounded spin, or an earlier exchange of this call already failed): the caller applies NO update, the column
// kernel behind it sees the same flag and returns, and the host raises the error at the end of bb_step.
__device__ __forceinline__ bool xchg_wait_and_sum(const XchgWaitArgs &x, double *sums) {
    if (!x.buf) return true;
    const int tid = threadIdx.x;
    int bad = 0;
    if (tid < x.world) {
        const unsigned long long *f = x.flag + (x.parity * x.world + tid);
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v == x.seq) break;
            if (*reinterpret_cast<volatile int *>(x.err) || clock64() - t0 > 6000000000LL) { *x.err = 1; bad = 1; break; }   // ~3 s: never hang the GPU
        }
    }
    if (__syncthreads_or(bad)) return false;
    __threadfence_system();
    for (int i = tid; i < x.P; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < x.world; ++r)
            s += *reinterpret_cast<const volatile double *>(x.buf + xchg_off(x.P, x.parity * x.world + r, i));
        sums[i] = s;
    }
    __syncthreads();
    return true;
}

// All-reduce (sum) of a short device vector over the peer exchange buffers: ONE CTA posts this rank's `m` values to
// every peer, waits for theirs and writes the rank-ordered sum back in place.  Used for the ELBO terms / trace when the
// handles were wired by bb_peer_attach (no NCCL communicator in the process).
static __global__ void __launch_bounds__(256) peer_allreduce_kernel(double *vec, int m, const XchgPostArgs xp,
                                                                    const XchgWaitArgs xw) {
    const int tid = threadIdx.x;
    for (int r = 0; r < xp.world; ++r) {
        double *dst = xp.peer_buf[r] + (size_t)(xp.parity * xp.world + xp.rank) * xp.P;     // auxiliary region: dense slots
        for (int i = tid; i < m; i += blockDim.x) dst[i] = vec[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < xp.world) {
        unsigned long long *f = xp.peer_flag[tid] + (xp.parity * xp.world + xp.rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(xp.seq) : "memory");
    }
    XchgWaitArgs w = xw;
    w.P = xp.P;
    int bad = 0;
    if (tid < w.world) {
        const unsigned long long *f = w.flag + (w.parity * w.world + tid);
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v == w.seq) break;
            if (clock64() - t0 > 6000000000LL) { *w.err = 1; bad = 1; break; }
        }
    }
    if (__syncthreads_or(bad)) return;
    __threadfence_system();
    for (int i = tid; i < m; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < w.world; ++r)
            s += *reinterpret_cast<const volatile double *>(w.buf + (size_t)(w.parity * w.world + r) * w.P + i);
        vec[i] = s;
    }
}

// ------------------------------------------------------------------ shared latents
template <typename real> struct SharedArgs {
    int R, K, tmax, nst;          // nst = sum_r (T_r - 1)
    int nt[MAX_SEG], sh0[MAX_SEG];
    double n_neutral;             // N over all shards
    double *sums;                 // all-reduced (NCCL), or this rank's partials to be completed by `xchg`
    XchgWaitArgs xchg;
    double2 *sh_th, *sh_acc;      // [2 nst] (s-bar block, then log-sigma-bar block)
    // TruncatedADAGrad ring of the shared latents, n + 1 slots of [2 nst]: step s evicts slot (s + 1) % (n + 1)
    // (= the g^2 of step s - n) and writes slot s % (n + 1).  Read and write never share a slot, so the CTAs of
    // the persistent step kernel can all read while one of them (ring_writer) writes.
    const double2 *sh_ring_rd;
    double2 *sh_ring_wr;
    int ring_writer;
    const double2 *sh_pr;         // (mean, 1/var)
    PhiloxKey key;
    uint32_t step;
    const double *eps_sh;         // supplied noise [K][2 nst] or nullptr
    int z_direct;
    real *ctx;                    // [R][K][3][tmax]
    double *scratch;              // per-sample gradients [K][2 nst], then eps, z [K][2 nst] each, u, lp [R][K][tmax] each
    double2 *gout;                // [2 nst] (dELBO/dmu, dELBO/domega) when !opt.update
    double *dump;                 // [K][2 nst] per-sample d log pi/dz or nullptr
    double *elbo_sh;              // [K+1]: neutral-likelihood + shared-prior log-density per k; sum log sigma
    OptArgsT<double> opt;
    int leader;                   // 1: this rank reports the shared latents' ELBO terms (rank 0)
    // as-written neutral pairing of the ragged replicate model (replicates.jl:599-605; single shard): the neutral
    // columns' log-ratio differences [R][K][N][tmax-1] written by pass 1, the s-bar draws handed to pass 2
    // [R][K][tmax], working arrays [3][R][K][tmax]; aw_d == nullptr: regular pairing
    const real *aw_d;
    real *aw_zs;
    double *aw_tmp;
    int aw_N;
};

__device__ __forceinline__ double softplus_d(double w) { return fmax(w, 0.0) + log1p(exp(-fabs(w))); }

// One CTA; every phase is parallel over its natural index and separated by a block barrier.
//   phase 0  (k, i)        noise and z of the shared latents                 -> eps_z scratch
//   phase 1  (r, k, t<T-1) c_t, residual sums u_t, per-sample gradients, ctx -> scratch
//   phase 2  (r, k, t<T)   G_Lambda_t = (u_{t-1} - u_t) / Lambda_t           -> ctx
//   phase 3  (i)           mean over samples, optimiser update (or emit)
// phase 0 of the shared-latent kernel: noise and z of the shared latents -> scratch (eps, z, log sigma).
// Independent of the step's sums: the merged tail kernel runs it on its own CTA beside the reduction.
template <typename real>
__device__ __forceinline__ void shared_phase0(const SharedArgs<real> &a, double *scratch) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n2 = 2 * a.nst;
    double *eps_t = scratch + (size_t)a.K * n2;
    double *z_t = eps_t + (size_t)a.K * n2;
    double *lsig_t = z_t + (size_t)a.K * n2 + (size_t)2 * a.R * a.K * a.tmax;
    for (int j = tid; j < a.K * n2; j += nthr) {
        const int k = j / n2, i = j % n2;
        const double e = a.eps_sh ? a.eps_sh[j]
                                  : stream_normal<double>(STREAM_SHARED, (uint32_t)i, (uint32_t)k, a.step, a.key);
        const double2 th = a.sh_th[i];
        const double sigma = softplus_d(th.y);
        eps_t[j] = e;
        z_t[j] = a.z_direct ? e : th.x + sigma * e;
        if (k == 0) lsig_t[i] = log(sigma);
    }
}

// `sums`: where the (all-reduced) totals are / will be; `scratch`: working arrays (global, or shared memory
// in the merged tail kernel); do_phase0 = false: eps / z / log sigma are already in `scratch`
template <typename real>
__device__ __forceinline__ void shared_body(const SharedArgs<real> &a, double *sums, double *scratch, bool do_phase0) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n2 = 2 * a.nst;
    double *eps_t = scratch + (size_t)a.K * n2;              // [K][n2] eps
    double *z_t = eps_t + (size_t)a.K * n2;                  // [K][n2] z
    double *u_t = z_t + (size_t)a.K * n2;                    // [R][K][tmax] sum_all w res
    double *lp_t = u_t + (size_t)a.R * a.K * a.tmax;         // [R][K][tmax] log-density pieces
    double *lsig_t = lp_t + (size_t)a.R * a.K * a.tmax;      // [n2] log sigma before the update
    if (do_phase0) shared_phase0<real>(a, scratch);
    if (!xchg_wait_and_sum(a.xchg, sums)) return;            // multi-GPU: complete the sums over NVLink peer memory
    __syncthreads();
    if (a.aw_d) {
        // ---- phase 1, as-written pairing (model_fitness_normal_hierarchical_replicates.jl:599-605): vec(logGamma_n) is
        // time-fastest, element kk = n (T_r - 1) + t, while the mean / variance vectors repeat every population latent N
        // times in a row, element kk -> latent kk div N.  Ratio (n, t) is therefore scored against
        // Normal(-sbar_p, sigma_p), p = (n (T_r - 1) + t) div N, and latent p collects the N ratios kk in [p N, (p+1) N).
        // The neutral residual sums cannot be formed from per-time totals (c_t varies inside a latent's set), so the
        // few neutral columns export their differences and the sums are taken here.
        const int RKt = a.R * a.K * a.tmax, N = a.aw_N;
        double *c_a = a.aw_tmp, *wb_a = c_a + RKt, *zs_a = wb_a + RKt;
        for (int j = tid; j < RKt; j += nthr) {
            const int t = j % a.tmax, k = (j / a.tmax) % a.K, r = j / (a.tmax * a.K);
            if (t >= a.nt[r] - 1) continue;
            const double *S = sums + (size_t)(r * a.K + k) * NQ * a.tmax;
            const double zs = z_t[(size_t)k * n2 + a.sh0[r] + t], zl = z_t[(size_t)k * n2 + a.nst + a.sh0[r] + t];
            c_a[j] = log(S[Q_LAM * a.tmax + t + 1]) - log(S[Q_LAM * a.tmax + t]);
            wb_a[j] = exp(-2.0 * zl);
            zs_a[j] = zs;
            a.aw_zs[j] = (real)zs;
        }
        __syncthreads();
        for (int j = tid; j < RKt; j += nthr) {
            const int t = j % a.tmax, k = (j / a.tmax) % a.K, r = j / (a.tmax * a.K);
            const int nt = a.nt[r], ntm = nt - 1;
            u_t[j] = 0.0; lp_t[j] = 0.0;
            if (t >= ntm) continue;
            const int is = a.sh0[r] + t, il = a.nst + a.sh0[r] + t;
            const size_t row = (size_t)(r * a.K + k) * a.tmax;
            const double *S = sums + (size_t)(r * a.K + k) * NQ * a.tmax;
            const real *d = a.aw_d + (size_t)(r * a.K + k) * N * (a.tmax - 1);
            real *ctx = a.ctx + (size_t)(r * a.K + k) * 3 * a.tmax;
            const double zs = zs_a[j], zl = z_t[(size_t)k * n2 + il], c = c_a[j], wbar = wb_a[j];
            double s1 = 0.0, s2 = 0.0;             // latent p = t: its N ratios
            for (int kk = t * N; kk < (t + 1) * N; ++kk) {
                const int n = kk / ntm, tp = kk - n * ntm;
                const double res = (double)d[(size_t)n * (a.tmax - 1) + tp] - c_a[row + tp] + zs;
                s1 += res; s2 += res * res;
            }
            double uc = 0.0;                       // time t: sum_n w res of its N ratios (the coupling through c_t)
            for (int n = 0; n < N; ++n) {
                const int p = (n * ntm + t) / N;
                uc += wb_a[row + p] * ((double)d[(size_t)n * (a.tmax - 1) + t] - c + zs_a[row + p]);
            }
            const double mut = S[Q_A * a.tmax + t] + (zs - c) * S[Q_W * a.tmax + t];   // mutants: regular pairing
            const double2 ps = a.sh_pr[is], pl = a.sh_pr[il];
            scratch[(size_t)k * n2 + is] = -(wbar * s1 + mut) - (zs - ps.x) * ps.y;
            scratch[(size_t)k * n2 + il] = wbar * s2 - a.n_neutral - (zl - pl.x) * pl.y;
            u_t[j] = uc + mut;
            lp_t[j] = -a.n_neutral * zl - 0.5 * wbar * s2 - 0.5 * (zs - ps.x) * (zs - ps.x) * ps.y -
                      0.5 * (zl - pl.x) * (zl - pl.x) * pl.y;
            ctx[0 * a.tmax + t] = (real)(c - zs);
            ctx[2 * a.tmax + t] = (real)wbar;
        }
    }
    // ---- phase 1 (ratio (n, t) paired with population latent t; nothing to do when the as-written branch above ran)
    const int n_phase1 = a.aw_d ? 0 : a.R * a.K * a.tmax;
    for (int j = tid; j < n_phase1; j += nthr) {
        const int t = j % a.tmax, k = (j / a.tmax) % a.K, r = j / (a.tmax * a.K);
        const int nt = a.nt[r];
        u_t[j] = 0.0; lp_t[j] = 0.0;
        if (t >= nt - 1) continue;
        const int is = a.sh0[r] + t, il = a.nst + a.sh0[r] + t;
        const double *S = sums + (size_t)(r * a.K + k) * NQ * a.tmax;
        real *ctx = a.ctx + (size_t)(r * a.K + k) * 3 * a.tmax;
        const double zs = z_t[(size_t)k * n2 + is], zl = z_t[(size_t)k * n2 + il];   // s-bar_t, log-sigma-bar_t
        const double c = log(S[Q_LAM * a.tmax + t + 1]) - log(S[Q_LAM * a.tmax + t]);
        const double wbar = exp(-2.0 * zl);
        const double av = zs - c;
        const double dn = S[Q_DN * a.tmax + t], d2n = S[Q_D2N * a.tmax + t];
        const double am = S[Q_A * a.tmax + t], wm = S[Q_W * a.tmax + t];
        const double qn = d2n + 2.0 * av * dn + a.n_neutral * av * av;   // sum_neutral res^2
        const double u = wbar * (dn + a.n_neutral * av) + (am + av * wm);   // sum_all w res
        const double2 ps = a.sh_pr[is], pl = a.sh_pr[il];
        scratch[(size_t)k * n2 + is] = -u - (zs - ps.x) * ps.y;
        scratch[(size_t)k * n2 + il] = wbar * qn - a.n_neutral - (zl - pl.x) * pl.y;
        u_t[j] = u;
        lp_t[j] = -a.n_neutral * zl - 0.5 * wbar * qn - 0.5 * (zs - ps.x) * (zs - ps.x) * ps.y -
                  0.5 * (zl - pl.x) * (zl - pl.x) * pl.y;
        ctx[0 * a.tmax + t] = (real)(c - zs);
        ctx[2 * a.tmax + t] = (real)wbar;
    }
    __syncthreads();
    // ---- phase 2
    for (int j = tid; j < a.R * a.K * a.tmax; j += nthr) {
        const int t = j % a.tmax, k = (j / a.tmax) % a.K, r = j / (a.tmax * a.K);
        if (t >= a.nt[r]) continue;
        const double *S = sums + (size_t)(r * a.K + k) * NQ * a.tmax;
        const double uprev = t > 0 ? u_t[j - 1] : 0.0;
        a.ctx[(size_t)(r * a.K + k) * 3 * a.tmax + 1 * a.tmax + t] = (real)((uprev - u_t[j]) / S[Q_LAM * a.tmax + t]);
    }
    // ---- phase 3
    const double invK = 1.0 / a.K;
    for (int i = tid; i < n2; i += nthr) {
        double2 th = a.sh_th[i];
        const double sigma = softplus_d(th.y);
        double sg = 0.0, sge = 0.0;
        for (int k = 0; k < a.K; ++k) {
            const double g = scratch[(size_t)k * n2 + i];
            sg += g; sge += g * eps_t[(size_t)k * n2 + i];
            if (a.dump) a.dump[(size_t)k * n2 + i] = g;
        }
        const double gm = sg * invK;
        const double go = (sge * invK + 1.0 / sigma) / (1.0 + exp(-th.y));
        if (a.opt.update) {
            double2 ac = a.sh_acc[i];
            double2 rg = make_double2(0.0, 0.0);
            const bool trunc = a.opt.kind == 0;
            if (trunc) rg = __ldcg(a.sh_ring_rd + i);                  // the slot evicted this step
            opt_apply<double>(a.opt, -gm, th.x, ac.x, rg.x);
            opt_apply<double>(a.opt, -go, th.y, ac.y, rg.y);
            a.sh_th[i] = th; a.sh_acc[i] = ac;
            if (trunc && a.ring_writer) a.sh_ring_wr[i] = rg;
        } else if (a.gout) {
            a.gout[i] = make_double2(gm, go);
        }
    }
    if (a.elbo_sh) {
        // fixed-order sums: deterministic ELBO terms of the shared latents (reported by rank 0 only)
        for (int k = tid; k <= a.K; k += nthr) {
            double s = 0.0;
            if (k < a.K) {
                for (int r = 0; r < a.R; ++r)
                    for (int t = 0; t < a.tmax; ++t) s += lp_t[(size_t)(r * a.K + k) * a.tmax + t];
            } else {
                for (int i = 0; i < n2; ++i) s += lsig_t[i];
            }
            a.elbo_sh[k] = a.leader ? s : 0.0;
        }
    }
}

template <typename real>
__global__ void __launch_bounds__(256) shared_kernel(const SharedArgs<real> a) {
    shared_body<real>(a, a.sums, a.scratch, true);
}

// noise of the shared latents for `nsteps` consecutive steps (what shared_phase0 would draw at each of them):
// out[(j * K + k) * n2 + i] for step step0 + j.  The persistent step kernel reads it instead of running the fp64
// Box-Muller on every CTA's critical path.
static __global__ void shared_noise_steps_kernel(const PhiloxKey key, uint32_t step0, int nsteps, int K, int n2, double *out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nsteps * K * n2) return;
    const int i = idx % n2, k = (idx / n2) % K, j = idx / (n2 * K);
    out[idx] = stream_normal<double>(STREAM_SHARED, (uint32_t)i, (uint32_t)k, step0 + (uint32_t)j, key);
}

// ------------------------------------------------------------------ merged step tail
// reduce (+ peer post) + shared latents in ONE launch: every CTA reduces its share of the partial sums;
// the CTA that finishes last (ticket) posts them to the peers (multi-GPU) and runs the shared-latent
// phases with its working arrays in shared memory.  Saves two launch boundaries per step and the
// global-memory round trips between the phases; the arithmetic and its order are those of
// reduce_kernel / xchg_post_kernel / shared_kernel, so the results are bitwise the same.
template <typename real>
__global__ void __launch_bounds__(256) tail_kernel(const ReduceArgs ra, int R, const XchgPostArgs xp,
                                                   const SharedArgs<real> sa, unsigned *ticket, int nsums) {
    extern __shared__ double tail_smem[];        // [nsums] totals | scratch
    __shared__ int is_last;
    asm volatile("griddepcontrol.launch_dependents;");      // let the next column kernel start its prologue
    // the extra CTA (index gridDim.x - 1, no reduction rows) draws the shared latents' noise beside the reduction
    if (blockIdx.x == gridDim.x - 1) shared_phase0<real>(sa, sa.scratch);
    else reduce_rows(ra, R);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = t == gridDim.x - 1 ? 1 : 0;
        if (is_last) *ticket = 0u;               // ready for the next launch (stream order)
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double *tot = tail_smem, *scratch = tail_smem + nsums;
    if (sa.xchg.buf) {
        // multi-GPU: this rank's sums go to every peer, the totals come back inside shared_body
        for (int r = 0; r < xp.world; ++r) {
            double *dst = xp.peer_buf[r];
            for (int i = threadIdx.x; i < xp.P; i += blockDim.x)
                dst[xchg_off(xp.P, xp.parity * xp.world + xp.rank, i)] = __ldcg(ra.sums + i);
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < xp.world) {
            unsigned long long *f = xp.peer_flag[threadIdx.x] + (xp.parity * xp.world + xp.rank);
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(xp.seq) : "memory");
        }
    } else {
        for (int i = threadIdx.x; i < nsums; i += blockDim.x) tot[i] = __ldcg(ra.sums + i);
        __syncthreads();
    }
    {
        // eps, z (contiguous) and log sigma of phase 0: global scratch -> shared-memory scratch
        const int n2 = 2 * sa.nst, kn = sa.K * n2;
        const size_t ls = (size_t)3 * kn + (size_t)2 * sa.R * sa.K * sa.tmax;
        for (int i = threadIdx.x; i < 2 * kn; i += blockDim.x) scratch[kn + i] = __ldcg(sa.scratch + kn + i);
        for (int i = threadIdx.x; i < n2; i += blockDim.x) scratch[ls + i] = __ldcg(sa.scratch + ls + i);
    }
    shared_body<real>(sa, tot, scratch, false);
}

// ------------------------------------------------------------------ hyper latents (theta of the hierarchical models)
template <typename real> struct HyperArgs {
    int H, K;
    uint32_t gid0;               // global hyper index of local hyper 0 (noise lattice slot)
    vec2<real> *hy_th, *hy_acc, *hy_ring;   // [H]; ring [n][H]
    const vec2<real> *hy_pr;     // (mean, 1/var) [H]
    vec2<real> *zeps;            // (z, eps) of hyper latent h, sample k at [k * hz_k + h * hz_h]
    int hz_k, hz_h;              // (H, 1): sample-major -- (1, K): a latent's K draws contiguous (scattered member columns)
    PhiloxKey key;
    uint32_t step;
    const real *eps_hy;          // supplied noise [K][H] or nullptr
    int z_direct;
    // update
    const int *csr_off;          // [H+1]
    const int *csr_mem;          // contribution slots (e * cpad + c)
    const vec2<real> *hcontrib;  // [E][cpad]
    vec2<real> *hsum;            // [H] member sums from hyper_gather_kernel, or nullptr (hyper_update gathers itself)
    const real *dump_hcontrib;   // [K][E * cpad] per-sample contributions or nullptr
    long long dump_stride;       // E * cpad
    real *dump;                  // [K][H] per-sample d log pi / d theta or nullptr
    vec2<real> *gout;            // [H]
    double *epart;               // [gridDim.x][K+1] or nullptr
    OptArgsT<real> opt;
    int prep_next;               // hyper_update_kernel: also draw (z, eps) of step + 1 from the updated theta (saves the
                                 // hyper_prep launch of the next step)
};

template <typename real>
__global__ void __launch_bounds__(BLOCK) hyper_prep_kernel(const HyperArgs<real> a) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= a.H) return;
    const vec2<real> th = a.hy_th[h];
    const real sigma = softplus_only<real>(th.y);
    for (int k = 0; k < a.K; ++k) {
        const real e = a.eps_hy ? a.eps_hy[(size_t)k * a.H + h]
                                : stream_normal<real>(STREAM_HYPER, a.gid0 + (uint32_t)h, (uint32_t)k, a.step,
                                                      a.key);
        const real z = a.z_direct ? e : fma(sigma, e, th.x);
        a.zeps[(size_t)k * a.hz_k + (size_t)h * a.hz_h] = mk2<real>(z, e);
    }
}

// Genotype-style hierarchies (few hyper latents, ~100 member columns each): one WARP per hyper latent gathers
// the member contributions -- lane l takes members l, l + 32, ... in order, then a fixed butterfly -- instead of
// one thread walking ~100 dependent (index, value) load pairs.  Deterministic; used when the mean member count
// is >= 8 (the replicate models, 2-3 members each, keep the in-thread loop of hyper_update_kernel).
template <typename real>
__global__ void __launch_bounds__(BLOCK) hyper_gather_kernel(const HyperArgs<real> a) {
    const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (h >= a.H) return;
    const int m0 = a.csr_off[h], m1 = a.csr_off[h + 1];
    real sg = real(0), sge = real(0);
    for (int m = m0 + lane; m < m1; m += 32) {
        const vec2<real> cb = a.hcontrib[a.csr_mem[m]];
        sg += cb.x; sge += cb.y;
    }
    sg = warp_sum<real>(sg);
    sge = warp_sum<real>(sge);
    if (lane == 0) a.hsum[h] = mk2<real>(sg, sge);
}

template <typename real>
__global__ void __launch_bounds__(BLOCK) hyper_update_kernel(const HyperArgs<real> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sel = reinterpret_cast<double *>(smem_raw);      // [K+1][BLOCK]
    const int tid = threadIdx.x;
    const int h = blockIdx.x * blockDim.x + tid;
    const bool want_elbo = a.epart != nullptr;
    if (want_elbo)
        for (int k = 0; k <= a.K; ++k) sel[k * BLOCK + tid] = 0.0;
    if (h < a.H) {
        vec2<real> th = a.hy_th[h];
        const real sigma = softplus_only<real>(th.y);
        const vec2<real> pr = a.hy_pr[h];
        real sg = real(0), sge = real(0);
        const int m0 = a.csr_off[h], m1 = a.csr_off[h + 1];
        if (a.hsum) {                                         // gathered by hyper_gather_kernel (many members)
            const vec2<real> hs = a.hsum[h];
            sg = hs.x; sge = hs.y;
        } else {
            for (int m = m0; m < m1; ++m) {                   // fixed member order -> deterministic
                const vec2<real> cb = a.hcontrib[a.csr_mem[m]];
                sg += cb.x; sge += cb.y;
            }
        }
        for (int k = 0; k < a.K; ++k) {
            const vec2<real> ze = a.zeps[(size_t)k * a.hz_k + (size_t)h * a.hz_h];
            const real dz = ze.x - pr.x;
            const real gp = -dz * pr.y;
            sg += gp; sge = fma(gp, ze.y, sge);
            if (want_elbo) sel[k * BLOCK + tid] += (double)(real(-0.5) * dz * dz * pr.y);
            if (a.dump) {
                real gk = gp;
                for (int m = m0; m < m1; ++m) gk += a.dump_hcontrib[(size_t)k * a.dump_stride + a.csr_mem[m]];
                a.dump[(size_t)k * a.H + h] = gk;
            }
        }
        if (want_elbo) sel[a.K * BLOCK + tid] = (double)bb_log(sigma);
        const real invK = real(1) / real(a.K);
        finish_latent<real>(a.opt, invK, sg, sge, th, a.hy_acc[h], a.hy_th + h, a.hy_acc + h,
                            a.hy_ring ? a.hy_ring + h : nullptr,
                            a.gout ? a.gout + h : nullptr);
        if (a.prep_next) {
            // what hyper_prep_kernel would do at the head of the next step, from the theta just written (this thread has
            // read all of zeps[.][h] above; nobody else touches column h)
            const vec2<real> tn = a.hy_th[h];
            const real sn = softplus_only<real>(tn.y);
            for (int k = 0; k < a.K; ++k) {
                const real e = stream_normal<real>(STREAM_HYPER, a.gid0 + (uint32_t)h, (uint32_t)k, a.step + 1u, a.key);
                a.zeps[(size_t)k * a.hz_k + (size_t)h * a.hz_h] = mk2<real>(fma(sn, e, tn.x), e);
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

// ------------------------------------------------------------------ ELBO assembly
// out[k] = sum of the variable log-density parts of sample k; out[K] = sum log sigma
static __global__ void elbo_sum_kernel(const double *p2part, int n2, const double *hypart, int nh, const double *elbo_sh,
                                int K, double *out) {
    const int k = blockIdx.x;        // one block per output row
    double s = 0.0;
    for (int b = threadIdx.x; b < n2; b += blockDim.x) s += p2part[(size_t)b * (K + 1) + k];
    for (int b = threadIdx.x; b < nh; b += blockDim.x) s += hypart[(size_t)b * (K + 1) + k];
    __shared__ double sm[128];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int j = 0; j < blockDim.x; ++j) t += sm[j];
        out[k] = t + (elbo_sh ? elbo_sh[k] : 0.0);
    }
}

// ------------------------------------------------------------------ layout permutations (map = reference index, -1 = padding)
template <typename real>
__global__ void scatter_pairs_kernel(vec2<real> *dst, const int *map, long long n, const double *x, const double *y,
                                     double fill_x, double fill_y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    dst[i] = m >= 0 ? mk2<real>((real)x[m], (real)y[m]) : mk2<real>((real)fill_x, (real)fill_y);
}

// TruncatedADAGrad, fp32: the running window sums rebuilt exactly from the window itself -- acc[i] = sum over the
// n_slots ring slots of (g_mu^2, g_omega^2), in double and in slot order.  The running form s <- max(s - evicted, 0) + g^2
// keeps, in fp32, a cancellation residue of size ulp(s) of every eviction; after the large squared gradients of the
// first steps have left the window that residue can exceed the true sum.  The engine runs this at
// window wraps (Engine::maybe_resum); upstream sums the window afresh at every step.
template <typename real>
__global__ void __launch_bounds__(256) ring_resum_kernel(const vec2<real> *ring, vec2<real> *acc, long long n_items,
                                                          int n_slots) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    double sx = 0.0, sy = 0.0;
    for (int s = 0; s < n_slots; ++s) {
        const vec2<real> v = ring[(size_t)s * (size_t)n_items + (size_t)i];
        sx += (double)v.x; sy += (double)v.y;
    }
    acc[i] = mk2<real>((real)sx, (real)sy);
}

template <typename real>
__global__ void gather_pairs_kernel(const vec2<real> *src, const int *map, long long n, double *x, double *y,
                                    int softplus_y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    if (m < 0) return;
    const vec2<real> v = src[i];
    if (x) x[m] = (double)v.x;
    if (y) y[m] = softplus_y ? softplus_d((double)v.y) : (double)v.y;
}

// K rows of scalars: dst[k][i] = src[k * D + map[i]]
template <typename real>
__global__ void scatter_rows_kernel(real *dst, const int *map, long long n, int K, const double *src, long long D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    for (int k = 0; k < K; ++k) dst[(size_t)k * n + i] = m >= 0 ? (real)src[(size_t)k * D + m] : real(0);
}

template <typename real>
__global__ void gather_rows_kernel(const real *src, const int *map, long long n, int K, double *dst, long long D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    if (m < 0) return;
    for (int k = 0; k < K; ++k) dst[(size_t)k * D + m] = (double)src[(size_t)k * n + i];
}

// mean-field initialisation: mu_j = n(INIT, j), omega_j = n(INIT, D + j) with j the reference index
template <typename real>
__global__ void init_params_kernel(vec2<real> *dst, const int *map, long long n, long long D, const PhiloxKey key) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    if (m < 0) { dst[i] = mk2<real>(0, 0); return; }
    const double mu = stream_normal<double>(STREAM_INIT, (uint32_t)m, 0u, 0u, key);
    const double om = stream_normal<double>(STREAM_INIT, (uint32_t)(D + m), 0u, 0u, key);
    dst[i] = mk2<real>((real)mu, (real)om);
}

template <typename T> __global__ void fill_kernel(T *p, long long n, T v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// the lattice's draws, reference order: eps[k * D + map[i]] for the column latents of one class row
template <typename real>
__global__ void noise_columns_kernel(const SegList segs, int cpad, int row, int is_bc, const uint32_t *col_id,
                                     const int *map_row, int K, long long D, uint32_t step, const PhiloxKey key,
                                     double *eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cpad) return;
    const int m = map_row[c];
    if (m < 0) return;
    int si = 0;
    for (int i = 1; i < segs.nseg; ++i) if (c >= segs.seg[i].col0) si = i;
    const Seg &sg = segs.seg[si];
    const uint32_t colid = col_id ? col_id[c] : sg.colid0 + (uint32_t)(c - sg.col0);
    const int slot = is_bc ? sg.nt + row : row;
    for (int k = 0; k < K; ++k) {
        real n[8];
        normals8<real>(colid, (STREAM_COLUMN << 24) | (uint32_t)(slot >> 3), (uint32_t)k, step, key, key.trig, n);
        real r = n[0];
#pragma unroll
        for (int l = 1; l < 8; ++l) r = (slot & 7) == l ? n[l] : r;
        eps[(size_t)k * D + m] = (double)r;
    }
}

template <typename real>
__global__ void noise_stream_kernel(uint32_t stream, uint32_t slot0, const int *map, int n, int K, long long D,
                                    uint32_t step, const PhiloxKey key, int as_double, double *eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    if (m < 0) return;
    for (int k = 0; k < K; ++k) {
        const double e = as_double ? stream_normal<double>(stream, slot0 + (uint32_t)i, (uint32_t)k, step, key)
                                   : (double)stream_normal<real>(stream, slot0 + (uint32_t)i, (uint32_t)k, step, key);
        eps[(size_t)k * D + m] = e;
    }
}

}  // namespace bb
