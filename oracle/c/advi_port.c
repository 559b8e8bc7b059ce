/*
 * advi_port.c -- compiled CPU restatement of the ADVI step of BarBay.model.fitness_normal.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: Julia is
 * not available here, so this is a *port* of the reference algorithm, validated against the
 * torch oracle (oracle/model_ref.py, oracle/advi_ref.py) in tests/test_oracle_cport.py.
 *
 * What it restates (paths relative to /root/reference):
 *   log-joint            src/model_fitness_normal.jl:131-271  (priors :137-203, frequencies and
 *                        log-ratios :209-219, Poisson x Multinomial counts :224-244 -- collapsed to
 *                        independent Poissons, exact because n_t = sum_b r_tb (src/utils.jl:431-432) --
 *                        neutral likelihood :251-257, mutant likelihood :262-270)
 *   ELBO / optimiser     AdvancedVI 0.2 as called from src/vi.jl:201 (see oracle/advi_ref.py header)
 * The gradient is analytic (the reference differentiates the same function with ReverseDiff /
 * ForwardDiff); fp64 throughout; OpenMP over barcodes.  This is the strong CPU baseline "B" of
 * Turing + ReverseDiff (single-threaded, taped) is slower than this.
 *
 * Latent order (VarInfo order): s_t[T-1], logsig_t[T-1], s_m[M], logsig_m[M], loglam[(b)T + t].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG2PI 1.8378770664093453

typedef struct {
    int T, N, M;            /* time points, neutrals, mutants */
    const int64_t *R;       /* counts, column-major T x B, neutrals first */
    double pri[5][2];       /* (mean, std): s_pop, logsig_pop, s_bc, logsig_bc, loglam (vector priors) */
} port_problem;

int port_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* the caller states the thread count (torchrun exports OMP_NUM_THREADS=1 to every rank) */
void port_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline double softplus(double w) { return (w > 0 ? w : 0) + log1p(exp(-fabs(w))); }
static inline double sigmoid(double w) { return 1.0 / (1.0 + exp(-w)); }

/* xoshiro256++ + Box-Muller: the reference draws eps with Julia's Xoshiro randn */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s, res = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static inline void rng_seed(rng_t *r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) {
        seed += 0x9E3779B97F4A7C15ULL;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        r->s[i] = z ^ (z >> 31);
    }
}
static inline void rng_normal2(rng_t *r, double *a, double *b) {
    double u = ((double)(rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double v = ((double)(rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double rad = sqrt(-2.0 * log(u));
    *a = rad * cos(6.283185307179586 * v); *b = rad * sin(6.283185307179586 * v);
}

/*
 * log pi(z) and d log pi / dz for one sample.  z, g: length D.  Returns log pi(z).
 * lgam_const = sum lgamma(r + 1) over all counts (data constant).
 */
static double logjoint_grad_one(const port_problem *p, const double *z, double *g, double lgam_const) {
    const int T = p->T, N = p->N, M = p->M, B = N + M;
    const double *st = z, *lst = z + (T - 1), *sm = z + 2 * (T - 1), *lsm = sm + M, *ll = lsm + M;
    double *g_st = g, *g_lst = g + (T - 1), *g_sm = g + 2 * (T - 1), *g_lsm = g_sm + M, *g_ll = g_lsm + M;
    double Lam[64], c[64], U[64], Qn[64], lp = 0.0;
    /* pass 1: Lambda_t (model_fitness_normal.jl:209-212) */
    for (int t = 0; t < T; ++t) Lam[t] = 0.0;
#pragma omp parallel
    {
        double loc[64];
        for (int t = 0; t < T; ++t) loc[t] = 0.0;
#pragma omp for schedule(static) nowait
        for (int b = 0; b < B; ++b)
            for (int t = 0; t < T; ++t) loc[t] += exp(ll[(size_t)b * T + t]);
#pragma omp critical
        for (int t = 0; t < T; ++t) Lam[t] += loc[t];
    }
    for (int t = 0; t < T - 1; ++t) { c[t] = log(Lam[t + 1]) - log(Lam[t]); U[t] = 0.0; Qn[t] = 0.0; }
    /* pass 2: likelihood terms and per-barcode gradients */
    double lp_par = 0.0;
#pragma omp parallel reduction(+ : lp_par)
    {
        double uloc[64], qloc[64];
        for (int t = 0; t < T - 1; ++t) { uloc[t] = 0.0; qloc[t] = 0.0; }
#pragma omp for schedule(static) nowait
        for (int b = 0; b < B; ++b) {
            const double *zb = ll + (size_t)b * T;
            double *gb = g_ll + (size_t)b * T;
            const int mut = b >= N, m = b - N;
            const double w_m = mut ? exp(-2.0 * lsm[m]) : 0.0;
            double uprev = 0.0, gs = 0.0, gq = 0.0;
            for (int t = 0; t < T; ++t) {
                const double lam = exp(zb[t]);
                const double r = (double)p->R[(size_t)b * T + t];
                const double dz = (zb[t] - p->pri[4][0]) / p->pri[4][1];
                gb[t] = (r - lam) - dz / p->pri[4][1];                         /* Poisson + prior :194-203 */
                lp_par += r * zb[t] - lam - 0.5 * dz * dz - log(p->pri[4][1]) - 0.5 * LOG2PI;
            }
            for (int t = 0; t < T - 1; ++t) {
                const double gamma = (zb[t + 1] - zb[t]) - c[t];               /* log(f_{t+1}/f_t) :215 */
                double w, res;
                if (mut) { w = w_m; res = gamma - (sm[m] - st[t]); }           /* :262-270 */
                else { w = exp(-2.0 * lst[t]); res = gamma + st[t]; }           /* :251-257 */
                const double u = w * res;
                uloc[t] += u;
                if (!mut) qloc[t] += u * res;
                gs += u; gq += u * res;
                gb[t] += u - uprev;
                uprev = u;
                lp_par += -0.5 * LOG2PI - (mut ? lsm[m] : lst[t]) - 0.5 * u * res;
            }
            gb[T - 1] -= uprev;
            if (mut) {
                const double d1 = (sm[m] - p->pri[2][0]) / p->pri[2][1], d2 = (lsm[m] - p->pri[3][0]) / p->pri[3][1];
                g_sm[m] = gs - d1 / p->pri[2][1];
                g_lsm[m] = gq - (T - 1) - d2 / p->pri[3][1];
                lp_par += -0.5 * d1 * d1 - log(p->pri[2][1]) - 0.5 * d2 * d2 - log(p->pri[3][1]) - LOG2PI;
            }
        }
#pragma omp critical
        for (int t = 0; t < T - 1; ++t) { U[t] += uloc[t]; Qn[t] += qloc[t]; }
    }
    lp += lp_par - lgam_const;
    /* shared latents */
    for (int t = 0; t < T - 1; ++t) {
        const double d1 = (st[t] - p->pri[0][0]) / p->pri[0][1], d2 = (lst[t] - p->pri[1][0]) / p->pri[1][1];
        g_st[t] = -U[t] - d1 / p->pri[0][1];
        g_lst[t] = Qn[t] - N - d2 / p->pri[1][1];
        lp += -0.5 * d1 * d1 - log(p->pri[0][1]) - 0.5 * d2 * d2 - log(p->pri[1][1]) - LOG2PI;
    }
    /* coupling through log Lambda_t: d/d loglam_tb += f_tb (U_{t-1} - U_t) */
    double GL[64];
    for (int t = 0; t < T; ++t) GL[t] = ((t > 0 ? U[t - 1] : 0.0) - (t < T - 1 ? U[t] : 0.0)) / Lam[t];
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t) g_ll[(size_t)b * T + t] += exp(ll[(size_t)b * T + t]) * GL[t];
    return lp;
}

static double lgamma_const(const port_problem *p) {
    const size_t n = (size_t)p->T * (p->N + p->M);
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t i = 0; i < n; ++i) s += lgamma((double)p->R[i] + 1.0);
    return s;
}

long port_n_latent(const port_problem *p) { return 2L * (p->T - 1) + 2L * p->M + (long)p->T * (p->N + p->M); }

/* ELBO and gradient of +ELBO wrt (mu, omega) for caller-supplied eps[K][D]; logp[K] optional. */
double port_elbo_grad(const port_problem *p, const double *mu, const double *omega, const double *eps, int K,
                      double *grad, double *logp) {
    const long D = port_n_latent(p);
    double *z = (double *)malloc(sizeof(double) * D), *g = (double *)malloc(sizeof(double) * D);
    double *sig = (double *)malloc(sizeof(double) * D);
    const double lg = lgamma_const(p);
    double elbo = 0.0, ent = 0.5 * D * (1.0 + LOG2PI);
    for (long i = 0; i < D; ++i) { sig[i] = softplus(omega[i]); ent += log(sig[i]); grad[i] = 0.0; grad[D + i] = 0.0; }
    for (int k = 0; k < K; ++k) {
        const double *e = eps + (size_t)k * D;
        for (long i = 0; i < D; ++i) z[i] = mu[i] + sig[i] * e[i];
        const double lp = logjoint_grad_one(p, z, g, lg);
        if (logp) logp[k] = lp;
        elbo += lp / K;
        for (long i = 0; i < D; ++i) { grad[i] += g[i] / K; grad[D + i] += g[i] * e[i] / K; }
    }
    for (long i = 0; i < D; ++i) grad[D + i] = (grad[D + i] + 1.0 / sig[i]) * sigmoid(omega[i]);
    free(z); free(g); free(sig);
    return elbo + ent;
}

/*
 * n_steps of AdvancedVI.optimize! with fresh eps per step.  theta = [mu, omega] (2D), acc (2D) is the
 * optimiser accumulator (DecayedADAGrad, kind 1: acc0 = 1e-8) or running sum (TruncatedADAGrad, kind 0,
 * ring = n x 2D zero-initialised, may be NULL for kind 1).  Returns the last ELBO estimate.
 */
double port_advi_steps(const port_problem *p, double *theta, double *acc, double *ring, int kind, double eta,
                       double tau_or_pre, double post, int n_ring, int K, int n_steps, uint64_t seed,
                       long first_step) {
    const long D = port_n_latent(p);
    double *z = (double *)malloc(sizeof(double) * D), *g = (double *)malloc(sizeof(double) * D);
    double *e = (double *)malloc(sizeof(double) * D), *sig = (double *)malloc(sizeof(double) * D);
    double *gm = (double *)malloc(sizeof(double) * 2 * D);
    const double lg = lgamma_const(p);
    double elbo = 0.0;
    for (int it = 0; it < n_steps; ++it) {
        const long step = first_step + it;
        double ent = 0.5 * D * (1.0 + LOG2PI);
#pragma omp parallel for reduction(+ : ent) schedule(static)
        for (long i = 0; i < D; ++i) { sig[i] = softplus(theta[D + i]); ent += log(sig[i]); gm[i] = 0.0; gm[D + i] = 0.0; }
        elbo = ent;
        for (int k = 0; k < K; ++k) {
#pragma omp parallel
            {
                rng_t r;
#ifdef _OPENMP
                const int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
                const int tid = 0, nth = 1;
#endif
                rng_seed(&r, seed ^ (uint64_t)(step * 1000003 + k) * 0x9E3779B97F4A7C15ULL ^ ((uint64_t)tid << 48));
                const long lo = D * tid / nth, hi = D * (tid + 1) / nth;
                for (long i = lo; i < hi; i += 2) {
                    double a, b;
                    rng_normal2(&r, &a, &b);
                    e[i] = a; z[i] = theta[i] + sig[i] * a;
                    if (i + 1 < hi) { e[i + 1] = b; z[i + 1] = theta[i + 1] + sig[i + 1] * b; }
                }
            }
            elbo += logjoint_grad_one(p, z, g, lg) / K;
#pragma omp parallel for schedule(static)
            for (long i = 0; i < D; ++i) { gm[i] += g[i] / K; gm[D + i] += g[i] * e[i] / K; }
        }
        const int slot = (int)(step % (n_ring > 0 ? n_ring : 1));
#pragma omp parallel for schedule(static)
        for (long j = 0; j < 2 * D; ++j) {
            double gr = j < D ? gm[j] : (gm[j] + 1.0 / sig[j - D]) * sigmoid(theta[j]);
            gr = -gr;                                        /* objective is -ELBO */
            double denom;
            if (kind == 1) {
                acc[j] = post * acc[j] + tau_or_pre * gr * gr;
                denom = sqrt(acc[j]) + 1e-8;
            } else {
                double *rs = ring + (size_t)slot * 2 * D + j;
                acc[j] = fmax(acc[j] - *rs + gr * gr, 0.0);
                *rs = gr * gr;
                denom = tau_or_pre + sqrt(acc[j]) + 1e-8;
            }
            theta[j] -= eta * gr / denom;
        }
    }
    free(z); free(g); free(e); free(sig); free(gm);
    return elbo;
}
