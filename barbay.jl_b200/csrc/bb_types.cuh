// Host/device shared argument structs of the column kernels.
//
// Device layout ("column space").  A *column* is one barcode time series of one
// replicate: T_r log-lambda latents plus, for mutant columns, the barcode-level
// latents (per environment e: (s, log-sigma), or (theta-tilde, log-tau, log-sigma)
// for hierarchical models).  Columns of one (replicate, population) pair form a
// *segment*; segments are laid back to back in a padded column index space of
// stride `cpad`, every segment start a multiple of 32.  All per-latent arrays are
// class-major SoA of (mu, omega) pairs: lam[t][c], bc[j][c] -- a warp reads 32
// consecutive pairs (256 B in fp32) per class.
#pragma once
#include "bb_device.cuh"

namespace bb {

struct Seg {
    int col0;        // first padded column index of the segment
    int ncol;        // valid columns (this shard)
    int rep;         // replicate index r
    int nt;          // time points T_r
    int neutral;     // 1 = neutral population
    uint32_t colid0; // canonical column id (noise lattice) of local column 0 when col_id == nullptr
    int blk0, blk1;  // CTAs [blk0, blk1) of the launch own this segment
    int sh0;         // offset of s-bar[r][0] inside the shared s-bar block
};

struct SegList {
    int nseg;
    Seg seg[MAX_SEG];
};

template <typename real> struct ColArrays {
    int cpad;                         // padded column count = row stride of every [class][c] array
    int tmax;                         // rows of the lam arrays
    int nj;                           // rows of the bc arrays = per * E
    vec2<real> *lam_th;               // [tmax][cpad] (mu, omega)
    vec2<real> *lam_acc;              // [tmax][cpad] optimiser accumulators (a_mu, a_omega)
    const int *cnt;                   // [tmax][cpad] barcode counts r_tb
    vec2<real> *bc_th;                // [nj][cpad]
    vec2<real> *bc_acc;               // [nj][cpad]
    const vec2<real> *lam_pr;         // per-latent prior (mean, 1/var) [tmax][cpad], or nullptr -> lam_pr_s
    const vec2<real> *bc_pr;          // [nj][cpad] or nullptr -> bc_pr_s[kind]
    vec2<real> lam_pr_s;
    vec2<real> bc_pr_s[3];            // by kind: direct {s, logsigma}; hier {theta-tilde, log-tau, log-sigma}
    const int *hgroup;                // [cpad] hyper-latent base index of the column (hier), else nullptr
    const uint32_t *col_id;           // [cpad] canonical column ids, or nullptr -> seg.colid0 + i
    vec2<real> *lam_ring;             // TruncatedADAGrad: the ring slot written this step, [tmax][cpad] of (g_mu^2, g_omega^2)
    vec2<real> *bc_ring;              // [nj][cpad] (the host offsets the [n][...] rings by slot)
};

template <typename real> struct OptArgsT {
    int kind;        // bb_opt_kind
    int update;      // 1: apply the optimiser; 0: only emit gradients
    real eta, tau, post;     // tau doubles as `pre` for DecayedADAGrad; typed by the kernel's arithmetic
};

// caller-supplied noise / per-sample dumps, device layout (SUPPLIED kernels only)
template <typename real> struct SupArgs {
    const real *eps_lam;   // [K][tmax][cpad]
    const real *eps_bc;    // [K][nj][cpad]
    int z_direct;          // 1: the supplied values are z themselves (mu, sigma ignored)
    real *dump_lam;        // d log pi / dz per sample [K][tmax][cpad], or nullptr
    real *dump_bc;         // [K][nj][cpad]
    real *dump_hcontrib;   // per-sample hyper contributions [K][E][cpad] (hier)
};

template <typename real> struct P1Args {
    SegList segs;
    ColArrays<real> cols;
    int K;
    int acc_slots;         // rows of BLOCK accumulators in shared memory (K-chunking budget)
    int ne;                // E
    int env_of_t[MAX_NT_DYN];   // 0-based environment of time point t
    PhiloxKey key;
    uint32_t step;
    const vec2<real> *hy_zeps;  // (z, eps) of hyper latent h, sample k at [k * hz_k + h * hz_h] (hier)
    int H, hz_k, hz_h;
    double *part;          // [gridDim.x][K][pv] block partial sums (double)
    int pv;                // slots per sample: nt + 2 (nt - 1)
    int nbuf;              // staging buffers (2, or 1 when shared memory is short)
    SupArgs<real> sup;
};

// As-written neutral pairing of the ragged replicate model (replicates.jl:599-605): arguments of aw_export_kernel
template <typename real> struct AwArgs {
    SegList segs;          // the neutral segments
    ColArrays<real> cols;
    int K, N;
    PhiloxKey key;
    uint32_t step;
    SupArgs<real> sup;
    real *d;               // [R][K][N][tmax-1] log-ratio differences d[t] = z[t+1] - z[t] per neutral column and sample
};

template <typename real> struct P2Args {
    SegList segs;
    ColArrays<real> cols;
    int K;
    int ne;
    int env_of_t[MAX_NT_DYN];
    PhiloxKey key;
    uint32_t step;
    const vec2<real> *hy_zeps;
    int H, hz_k, hz_h;
    const real *ctx;       // [R][K][3][tmax]: (c_t - sbar_t), G_Lambda_t, wbar_t
    int tmax_ctx;
    OptArgsT<real> opt;
    vec2<real> *gout_lam;  // (dELBO/dmu, dELBO/domega) [tmax][cpad] when !opt.update, else nullptr
    vec2<real> *gout_bc;   // [nj][cpad]
    vec2<real> *hcontrib;  // [E][cpad] (sum_k g_s, sum_k g_s eps_theta) (hier)
    double *epart;         // [gridDim.x][K+1] ELBO partials (log pi variable part per k; sum log sigma), or nullptr
    int stage_pr;          // 1: per-latent (matrix) priors are staged through shared memory
    int stage_ring;        // 1: TruncatedADAGrad update -> the evicted ring slot is staged too
    int stage_acc;         // 1: accumulators (and priors / ring) go through the stage; 0: read from global
    int l2_ring;           // 1: the evicted ring slot is NOT staged: prefetch it into L2 at tile start
    int nbuf;              // staging buffers: 2 = prefetch the next tile, 1 = no overlap (large T x E)
    double *part;          // fused step: block partial sums of the next step's pass 1, [K][pv][gridDim.x]
    int pv;                // row stride of part (3 nt - 2)
    int acc_slots;         // fused step: rows of BLOCK accumulators in shared memory
    const int *abort;      // multi-GPU: set by the tail kernel when the exchange of this step failed -> no update
    // as-written neutral pairing (replicates.jl:599-605): ratio (t, n) of neutral n is paired with population latent
    // p = (n (T_r - 1) + t) div N; aw_zs = this step's s-bar draws [R][K][tmax] (the context rows hold c_t - sbar_t and
    // wbar_t under the regular index); nullptr otherwise
    const real *aw_zs;
    int aw_N;
    SupArgs<real> sup;
};

}  // namespace bb
