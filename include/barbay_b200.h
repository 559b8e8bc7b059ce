/*
 * barbay_b200.h -- C ABI of the B200-native ADVI backend for BarBay.jl.
 *
 * Drop-in boundary: the one statement of the reference this library replaces is
 *     q = Turing.vi(bayes_model, advi; optimizer=opt)            (src/vi.jl:201)
 * together with the Turing model bodies it evaluates (src/model_*.jl).  The
 * caller (BarBay.vi.advi, src/vi.jl:86-235) packs its tidy DataFrame with
 * utils.data_to_arrays (src/utils.jl:996-1033), hands the packed arrays to
 * bb_create(), runs bb_step() for `max_iters` iterations and reads back
 * (q.dist.m, q.dist.sigma) with bb_get_posterior() for the unchanged
 * utils.advi_to_df (src/utils.jl:1409-1462).
 *
 * Conventions
 *  - plain C, no torch / CUDA types in any signature; every pointer is a HOST
 *    pointer owned by the caller unless stated otherwise; the library copies in
 *    bb_create and owns all device memory until bb_destroy.
 *  - latent vectors (mu, omega, m, sigma, z, eps, grad) are in the reference's
 *    VarInfo order, length D = bb_n_latent().
 *  - every function returns 0 on success, non-zero on error; bb_last_error()
 *    gives the message (the Julia glue raises it with error(msg), matching the
 *    reference's ErrorException convention, src/vi.jl:107,112,117).
 *  - one handle = one host thread (the reference call is synchronous and
 *    single-threaded).  Multi-GPU, two ways: (a) bb_desc.n_devices = N -- ONE
 *    handle, N GPUs of this process, every entry point one blocking call (the
 *    library runs one worker thread per device inside the call); (b) one process
 *    per GPU, each creating a handle with its (rank, world) and bb_comm_* wiring
 *    the exchange (torchrun-style launchers).
 */
#ifndef BARBAY_B200_H
#define BARBAY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB_ABI_VERSION 2   /* 2: bb_desc.n_devices (single-call multi-GPU) */

typedef struct bb_handle bb_handle;

/* Model variant <- name of the Turing model function (src/vi.jl:111-169 dispatches
 * on substrings of the function name). */
enum bb_model {
    BB_MODEL_FITNESS_NORMAL = 0,            /* src/model_fitness_normal.jl:120-272 */
    BB_MODEL_REPLICATE = 1,                 /* src/model_fitness_normal_hierarchical_replicates.jl:145-332, 407-638 */
    BB_MODEL_MULTIENV = 2,                  /* src/model_multienv_fitness_normal.jl:133-303 */
    BB_MODEL_GENOTYPE = 3,                  /* src/model_fitness_normal_hierarchical_genotypes.jl:151-330 */
    BB_MODEL_MULTIENV_REPLICATE = 4         /* src/model_multienv_fitness_normal_hierarchical_replicates.jl:158-363 */
};

enum bb_dtype { BB_F32 = 0, BB_F64 = 1 };

/* opt::Union{TruncatedADAGrad,DecayedADAGrad} (src/vi.jl:99) */
enum bb_opt_kind { BB_OPT_TRUNCATED_ADAGRAD = 0, BB_OPT_DECAYED_ADAGRAD = 1 };

/* One prior keyword (src/model_fitness_normal.jl:125-129): either the 2-vector
 * [mean, std] (is_matrix = 0, data[0..1]) or an n x 2 Julia Matrix{Float64} in
 * column-major memory order (is_matrix = 1: data[0..n) means, data[n..2n) stds). */
typedef struct bb_prior {
    const double *data;
    int64_t n;
    int32_t is_matrix;
} bb_prior;

typedef struct bb_opt {
    int32_t kind;        /* enum bb_opt_kind */
    double eta;          /* both: learning rate (default 0.1) */
    double tau;          /* Truncated: tau (1.0)   | Decayed: pre  (1.0) */
    double post;         /* Decayed: post (0.9)    | Truncated: unused */
    int32_t n;           /* Truncated: window length (100) */
} bb_opt;

/* Everything data_to_arrays produced (DataArrays, src/utils.jl:48-61) plus the
 * model keyword arguments and the ADVI settings. */
typedef struct bb_desc {
    int32_t abi_version;          /* BB_ABI_VERSION */
    int32_t model;                /* enum bb_model */
    int32_t dtype;                /* enum bb_dtype: arithmetic of the kernels */
    int32_t n_rep;                /* R; 1 for non-replicate models */
    const int32_t *n_time;        /* [n_rep] time points per replicate (DataArrays.n_time) */
    int32_t n_neutral;            /* N (DataArrays.n_neutral) */
    int32_t n_bc;                 /* M (DataArrays.n_bc) */
    /* DataArrays.bc_count in Julia memory order: Matrix T x B, Array{Int64,3} T x B x R,
     * or the R matrices T_r x B of a Vector{Matrix{Int64}} back to back; neutral columns first. */
    const int64_t *bc_count;
    int32_t n_env;                /* E; 1 unless multienv */
    /* 1-based indexin(envs, unique(envs)) (multienv.jl:151-155) or NULL: [n_time[0]] -- one list shared by all
     * replicates -- or, with env_per_rep = 1, the replicates' own lists back to back, [sum_r n_time[r]]
     * (indexin.(envs, Ref(unique(vcat(envs...)))), model_multienv_fitness_normal_hierarchical_replicates.jl:465-472) */
    const int32_t *env_idx;
    int32_t n_geno;               /* G; 0 unless genotype model */
    const int32_t *geno_idx;      /* [n_bc] 1-based indexin(genotypes, unique(genotypes)) (genotypes.jl:170-174) or NULL */
    bb_prior s_pop_prior, logsig_pop_prior, s_bc_prior, logsig_bc_prior, loglam_prior, logtau_prior;
    /* Replicate model with unequal T per replicate (the Vector{Matrix{Int64}} method): 1 = the neutral pairing as
     * written in replicates.jl:599-605 -- ratio k of vec(logGamma_n) against sbar[ceil(k / N)]; the reference's
     * behaviour; world must be 1 -- 0 = ratio (t, n) against sbar[t] like every other method of the reference. */
    int32_t ragged_as_written;
    int32_t n_samples;            /* K = advi.samples_per_step (src/vi.jl:98) */
    uint64_t seed;                /* key of the Philox noise lattice */
    int32_t device;               /* CUDA ordinal, -1 = current device */
    int32_t rank, world;          /* this handle owns shard `rank` of `world` of the barcode axis */
    /* ABI 2.  0 or 1: one GPU (`device`).  N > 1: ONE handle drives the N GPUs device .. device + N - 1 of this
     * process (rank / world must be 0 / 1): the barcode axis is sharded inside the library, every entry point
     * stays one blocking call, and the per-step exchange runs over NVLink peer memory -- what a single
     * BarBay.vi.advi() call (src/vi.jl:86-101) needs to use a whole box. */
    int32_t n_devices;
    /* 1: env_idx holds one environment list per replicate (the Vector{Matrix{Int64}} method of the multienv x
     * replicate model, ...replicates.jl:449-687; required when the replicates have unequal numbers of time points) */
    int32_t env_per_rep;
} bb_desc;

/* ---- lifecycle ---- */
int bb_create(const bb_desc *desc, bb_handle **out);
void bb_destroy(bb_handle *h);
const char *bb_last_error(const bb_handle *h);   /* h may be NULL: error of the last failed bb_create on this thread */
int64_t bb_n_latent(const bb_handle *h);          /* D */
int32_t bb_abi_version(void);

/* Host-only dry run of the shard layout (no CUDA call; usable without a GPU): which latents of the reference's
 * VarInfo order the shard (desc->rank, desc->world) owns.  owned[i] (i < D, caller-allocated, may be NULL) = number of
 * device slots that map to latent i on this shard: over all ranks every latent must be owned exactly once (the
 * replicated population latents are reported by rank 0).  info[8] = {D, first / end neutral column, first / end
 * mutant position (genotype models: in the genotype-sorted order), hyper latents owned, first global hyper index,
 * padded columns}.  Same validation and error messages as bb_create. */
int bb_layout_probe(const bb_desc *desc, int32_t *owned, int64_t info[8]);

/* ---- variational parameters theta = (mu, omega), sigma = softplus(omega) ---- */
int bb_init_params(bb_handle *h, uint64_t seed);                         /* meanfield(): mu, omega ~ N(0,1) */
int bb_set_params(bb_handle *h, const double *mu, const double *omega);  /* [D] each */
int bb_get_params(bb_handle *h, double *mu, double *omega);
int bb_get_posterior(bb_handle *h, double *m, double *sigma);            /* q.dist.m, q.dist.sigma (utils.jl:1060) */

/* ---- parity entry points (caller-supplied noise) ---- */
/* logp[k] = log pi(z_k), grad[k*D + i] = d log pi / d z_i (z_k); z_k = eps_is_noise ? mu + sigma.*x_k : x_k. */
int bb_logjoint_grad(bb_handle *h, const double *z_or_eps, int32_t n_samples, int32_t eps_is_noise,
                     double *logp, double *grad);
/* ELBO estimate and gradient of +ELBO w.r.t. (mu, omega) -> grad[2D]; eps == NULL -> Philox lattice at `step`. */
int bb_elbo_grad(bb_handle *h, const double *eps, int64_t step, double *elbo, double *grad);
/* The lattice's draws for `step`: eps[k*D + i]. */
int bb_get_noise(bb_handle *h, int64_t step, double *eps);

/* ---- optimisation: AdvancedVI.optimize! ---- */
int bb_set_optimizer(bb_handle *h, const bb_opt *opt);                   /* resets accumulators and the step counter */
int bb_step(bb_handle *h, int32_t n_steps, double *elbo_trace);          /* elbo_trace: [n_steps] or NULL */
int bb_step_with_noise(bb_handle *h, const double *eps);                 /* one step with caller noise eps[K*D] */
int64_t bb_step_count(const bb_handle *h);
/* ELBO-trace convergence stop -- an extension: the reference always runs max_iters steps (src/vi.jl:98, 201).
 * Steps run in blocks of `every` (every - 1 production steps + 1 step that also evaluates the ELBO); the run stops
 * when the mean of the last `window` ELBO estimates differs from the mean of the `window` estimates before them by
 * at most rel_tol * |mean|, or after max_iters steps.  n_done / converged are outputs; elbo_out (nullable) receives
 * the estimates, at most max_iters / every + 1 of them, *n_elbo their number. */
int bb_step_until(bb_handle *h, int32_t max_iters, int32_t every, int32_t window, double rel_tol, int32_t *n_done,
                  int32_t *converged, double *elbo_out, int32_t *n_elbo);

/* ---- state (checkpoint / resume; the reference keeps none) ----
 * [step_count, ring_slot, mu[D], omega[D], acc_mu[D], acc_omega[D]] and, for TruncatedADAGrad, the whole window
 * (n slots of squared gradients in latent order + the shared latents' ring): 2 + 4 D (+ 2 n D + 4 nst (n + 1))
 * doubles, bb_state_size() for the current optimiser.  A run restored on a fresh handle (same problem, same
 * optimiser) continues exactly like the uninterrupted one. */
int64_t bb_state_size(const bb_handle *h);                               /* doubles needed by get/set_state */
int bb_get_state(bb_handle *h, double *state);
int bb_set_state(bb_handle *h, const double *state);

/* ---- plumbing ---- */
int bb_set_stream(bb_handle *h, void *cuda_stream);     /* run on the caller's cudaStream_t (NULL -> library stream) */
int bb_sync(bb_handle *h);
int64_t bb_launch_count(const bb_handle *h);            /* kernels launched by this handle so far */
double bb_algorithmic_bytes_per_step(const bb_handle *h);   /* SURVEY §8d figure for this shard */
/* n_steps of bb_step timed with CUDA events on the launching stream: whole region, and the summed
 * durations of the pass-1 and pass-2 column kernels (the roofline numerators of bench.py). */
int bb_time_steps(bb_handle *h, int32_t n_steps, float *ms_total, float *ms_pass1, float *ms_pass2);

/* Persistent step kernel (several ADVI steps per launch, the tail of each step inside the kernel): SM cycles one
 * CTA of the mutant population spent in {column phase, arrival -> all ranks' sums (grid reduction + NVLink exchange),
 * sums -> next step's context}, summed over the in-kernel tails since the last call; out[3] = number of such
 * tails, out[4] = SM clock in kHz, out[5..6] = the last phase split into {completing the sums, shared-latent
 * phases}, out[7] = cycles the last-arriving CTA needed from the group sum to the flags posted to the peers,
 * out[8] = that CTA's column phase (the longest of the grid; minus out[0] = the tile-count imbalance), out[9..12] =
 * out[7] split into {group sum, tickets + fences, rank sum + stores to the peers, fence + flags}, out[13..15] = 0.
 * Resets the counters. */
int bb_persist_stats(bb_handle *h, double out[16]);

/* Derived `bc_fitness` rows of the hierarchical models: utils.advi_to_df -> process_hierarchical_samples!
 * (src/utils.jl:1284-1343) draws n_samples (default 10 000) of theta + exp(log-tau) * theta-tilde per
 * (barcode[, environment], replicate) from the fitted Normals and reports their median (`mean` column, :1315) and
 * sample standard deviation.  Computed on the device from the handle's current posterior; median / sd have
 * bb_n_derived() entries in the order of the log-tau rows (entries owned by other shards are 0). */
int bb_derived_fitness(bb_handle *h, int32_t n_samples, uint64_t seed, double *median, double *sd);
int64_t bb_n_derived(const bb_handle *h);

/* BarBay.stats.naive_prior (src/stats.jl:1175-1359) on the device, from the packed counts: the empirical priors the
 * documented workflow computes before every fit (docs/src/examples.md:121-140).  bc_count / n_rep / n_time /
 * n_neutral / n_bc as in bb_desc, the counts ALREADY incremented by the pseudocount (the reference adds it to the
 * caller's frame before packing, stats.jl:1185).  Outputs (caller-allocated): s_pop_prior and logsig_pop_prior,
 * sum_r (n_time[r] - 1) entries each, time fastest then replicate (-mean and -sd of the neutral log-frequency ratios
 * without their +-Inf entries, :1298, :1338); loglam_prior = log.(bc_count)[:], sum_r n_time[r] * (n_neutral + n_bc)
 * entries in the memory order of bc_count (:1345-1352).  device: CUDA ordinal, -1 = current.  No handle is needed;
 * errors are reported through bb_last_error(NULL). */
int bb_naive_prior(const int64_t *bc_count, int32_t n_rep, const int32_t *n_time, int32_t n_neutral, int32_t n_bc,
                   int32_t device, double *s_pop_prior, double *logsig_pop_prior, double *loglam_prior);

/* Which kernels / data plane this handle runs: out = {packed step kernel in use, steps per persistent launch
 * (0: one launch pair per step), NVLink peer-memory exchange on, NCCL communicator present, resident CTAs per SM of
 * the step kernel, its staging buffers (1 / 2), accumulators staged (0 / 1), its grid size}. */
int bb_data_plane(bb_handle *h, int32_t out[8]);

/* ---- multi-GPU, one process per GPU, WITHOUT NCCL: every rank exports the CUDA IPC handle (64 bytes) of its
 * exchange buffer, the caller gathers the `world` handles in rank order with whatever transport it has
 * (torch.distributed / MPI / Distributed.jl / a file) and attaches them.  The per-step exchange and the ELBO
 * reductions then run over NVLink peer memory only; no communicator is created (saves ~1-3 s of ncclCommInitRank). */
int bb_peer_handle(bb_handle *h, char out[64]);
int bb_peer_attach(bb_handle *h, const char *handles /* [world][64] */, int32_t n);

/* ---- multi-GPU: one process per GPU; id is an ncclUniqueId (128 bytes) ---- */
int bb_comm_unique_id(char id[128]);
int bb_comm_init(bb_handle *h, const char id[128]);

#ifdef __cplusplus
}
#endif
#endif /* BARBAY_B200_H */
