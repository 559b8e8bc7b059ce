/*
 * advi_port_models.c -- compiled CPU restatement of the ADVI step for ALL model families of the path
 * (fitness_normal, replicate, multienv, genotype, multienv x replicate; equal time points per replicate,
 * vector priors).
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: Julia is not available
 * here, so this is a *port* of the reference algorithm, validated against the torch oracle
 * (oracle/model_ref.py) in tests/test_oracle_cport.py.  It exists so that the BASELINE configurations 3-5
 * have a full-size parity check (10^6 barcodes in ~0.1 s per evaluation) and a CPU baseline.
 *
 * What it restates (paths relative to /root/reference):
 *   src/model_fitness_normal.jl:131-271                         (single condition)
 *   src/model_fitness_normal_hierarchical_replicates.jl:157-331 (s_mr = theta_m + exp(logtau_mr) thetatilde_mr, :216)
 *   src/model_multienv_fitness_normal.jl:145-302                (ratio t -> t+1 uses the environment of t+1, :293-301)
 *   src/model_fitness_normal_hierarchical_genotypes.jl:164-329  (s_m = theta_g(m) + exp(logtau_m) thetatilde_m, :230)
 *   src/model_multienv_fitness_normal_hierarchical_replicates.jl:171-362
 * Counts: Poisson(n_t | Lambda_t) x Multinomial(r_t | n_t, f_t) collapsed to independent Poissons (exact because
 * n_t = sum_b r_tb, src/utils.jl:431-432).  Gradient: analytic.  fp64, OpenMP over columns.
 *
 * Latent order (VarInfo order):
 *   s_t[(r)(T-1)+t], logsig_t[...], [theta[H]], kind_0[e + E (m + M r)], kind_1[...], [kind_2[...]],
 *   loglam[((r) B + b) T + t]
 * with kinds (s, logsig) or -- hierarchical -- (thetatilde, logtau, logsig); H = G (genotype: theta[g]) or E M
 * (replicates: theta[e + E m]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG2PI 1.8378770664093453
#define MAXT 64
#define MAXE 8

typedef struct {
    int T, N, M, R, E, G;        /* G = 0 unless genotype model */
    int hier;                    /* 1: replicate / genotype / multienv x replicate */
    const int32_t *env_of_t;     /* [T] 0-based environment of each time point (E == 1: may be NULL) */
    const int32_t *geno;         /* [M] 0-based genotype of each mutant (G > 0) */
    const int64_t *counts;       /* Julia memory order of T x B x R */
    double pri[6][2];            /* (mean, std): s_pop, logsig_pop, s_bc, logsig_bc, loglam, logtau */
} gport_problem;

void gport_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int gport_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline double softplus(double w) { return (w > 0 ? w : 0) + log1p(exp(-fabs(w))); }
static inline double sigmoid(double w) { return 1.0 / (1.0 + exp(-w)); }

typedef struct {
    long nst, H, blk, per, o_st, o_lst, o_h, o_k[3], o_ll, D;
} gport_layout;

static gport_layout layout_of(const gport_problem *p) {
    gport_layout L;
    L.nst = (long)p->R * (p->T - 1);
    L.H = p->hier ? (p->G > 0 ? p->G : (long)p->E * p->M) : 0;
    L.blk = (long)p->E * p->M * p->R;
    L.per = p->hier ? 3 : 2;
    long o = 0;
    L.o_st = o; o += L.nst;
    L.o_lst = o; o += L.nst;
    L.o_h = o; o += L.H;
    for (int k = 0; k < 3; ++k) { L.o_k[k] = o; if (k < L.per) o += L.blk; }
    L.o_ll = o; o += (long)p->T * (p->N + p->M) * p->R;
    L.D = o;
    return L;
}

long gport_n_latent(const gport_problem *p) { return layout_of(p).D; }

static inline double npdf(double x, double m, double s) {      /* log Normal(x | m, s) */
    const double d = (x - m) / s;
    return -0.5 * LOG2PI - log(s) - 0.5 * d * d;
}

/* log pi(z) and d log pi / dz for one sample; lgam_const = sum lgamma(r + 1) over all counts */
static double logjoint_grad_one(const gport_problem *p, const double *z, double *g, double lgam_const) {
    const int T = p->T, N = p->N, M = p->M, B = N + M, R = p->R, E = p->E;
    const gport_layout L = layout_of(p);
    const double *st = z + L.o_st, *lst = z + L.o_lst, *th = z + L.o_h, *ll = z + L.o_ll;
    const double *k0 = z + L.o_k[0], *k1 = z + L.o_k[1], *k2 = z + L.o_k[2];
    double *g_st = g + L.o_st, *g_lst = g + L.o_lst, *g_th = g + L.o_h, *g_ll = g + L.o_ll;
    double *g0 = g + L.o_k[0], *g1 = g + L.o_k[1], *g2 = g + L.o_k[2];
    int env[MAXT], n_of_e[MAXE];
    for (int e = 0; e < E; ++e) n_of_e[e] = 0;
    for (int t = 0; t < T; ++t) env[t] = (p->env_of_t && E > 1) ? p->env_of_t[t] : 0;
    for (int t = 1; t < T; ++t) n_of_e[env[t]] += 1;
    double lp = -lgam_const;
    for (long h = 0; h < L.H; ++h) g_th[h] = 0.0;

    for (int r = 0; r < R; ++r) {
        const double *llr = ll + (size_t)r * B * T;
        double *g_llr = g_ll + (size_t)r * B * T;
        const double *str = st + (size_t)r * (T - 1), *lstr = lst + (size_t)r * (T - 1);
        double Lam[MAXT], c[MAXT], U[MAXT], Qn[MAXT];
        for (int t = 0; t < T; ++t) Lam[t] = 0.0;
#pragma omp parallel
        {
            double loc[MAXT];
            for (int t = 0; t < T; ++t) loc[t] = 0.0;
#pragma omp for schedule(static) nowait
            for (int b = 0; b < B; ++b)
                for (int t = 0; t < T; ++t) loc[t] += exp(llr[(size_t)b * T + t]);
#pragma omp critical
            for (int t = 0; t < T; ++t) Lam[t] += loc[t];
        }
        for (int t = 0; t < T - 1; ++t) { c[t] = log(Lam[t + 1]) - log(Lam[t]); U[t] = 0.0; Qn[t] = 0.0; }
        double lp_par = 0.0;
#pragma omp parallel reduction(+ : lp_par)
        {
            double uloc[MAXT], qloc[MAXT];
            for (int t = 0; t < T - 1; ++t) { uloc[t] = 0.0; qloc[t] = 0.0; }
#pragma omp for schedule(static) nowait
            for (int b = 0; b < B; ++b) {
                const double *zb = llr + (size_t)b * T;
                double *gb = g_llr + (size_t)b * T;
                const int mut = b >= N, m = b - N;
                double s_e[MAXE], w_e[MAXE], gs[MAXE], gq[MAXE], extau[MAXE];
                if (mut) {
                    for (int e = 0; e < E; ++e) {
                        const size_t i = (size_t)e + (size_t)E * ((size_t)m + (size_t)M * r);
                        if (p->hier) {
                            const long h = p->G > 0 ? p->geno[m] : (long)m * E + e;
                            extau[e] = exp(k1[i]);
                            s_e[e] = th[h] + extau[e] * k0[i];
                            w_e[e] = exp(-2.0 * k2[i]);
                        } else {
                            s_e[e] = k0[i];
                            w_e[e] = exp(-2.0 * k1[i]);
                        }
                        gs[e] = 0.0; gq[e] = 0.0;
                    }
                }
                for (int t = 0; t < T; ++t) {
                    const double lam = exp(zb[t]);
                    const double rr = (double)p->counts[((size_t)r * B + b) * T + t];
                    const double dz = (zb[t] - p->pri[4][0]) / p->pri[4][1];
                    gb[t] = (rr - lam) - dz / p->pri[4][1];
                    lp_par += rr * zb[t] - lam - 0.5 * dz * dz - log(p->pri[4][1]) - 0.5 * LOG2PI;
                }
                double uprev = 0.0;
                for (int t = 0; t < T - 1; ++t) {
                    const double gamma = (zb[t + 1] - zb[t]) - c[t];
                    const int e = env[t + 1];
                    double w, res, lsig;
                    if (mut) {
                        const size_t i = (size_t)e + (size_t)E * ((size_t)m + (size_t)M * r);
                        w = w_e[e]; res = gamma - (s_e[e] - str[t]); lsig = p->hier ? k2[i] : k1[i];
                    } else { w = exp(-2.0 * lstr[t]); res = gamma + str[t]; lsig = lstr[t]; }
                    const double u = w * res;
                    uloc[t] += u;
                    if (!mut) qloc[t] += u * res;
                    else { gs[e] += u; gq[e] += u * res; }
                    gb[t] += u - uprev;
                    uprev = u;
                    lp_par += -0.5 * LOG2PI - lsig - 0.5 * u * res;
                }
                gb[T - 1] -= uprev;
                if (mut) {
                    for (int e = 0; e < E; ++e) {
                        const size_t i = (size_t)e + (size_t)E * ((size_t)m + (size_t)M * r);
                        if (p->hier) {
                            const long h = p->G > 0 ? p->geno[m] : (long)m * E + e;
                            g0[i] = gs[e] * extau[e] - k0[i];                                  /* thetatilde ~ N(0, 1) */
                            g1[i] = gs[e] * extau[e] * k0[i] - (k1[i] - p->pri[5][0]) / (p->pri[5][1] * p->pri[5][1]);
                            g2[i] = gq[e] - n_of_e[e] - (k2[i] - p->pri[3][0]) / (p->pri[3][1] * p->pri[3][1]);
#pragma omp atomic
                            g_th[h] += gs[e];
                            lp_par += npdf(k0[i], 0.0, 1.0) + npdf(k1[i], p->pri[5][0], p->pri[5][1]) +
                                      npdf(k2[i], p->pri[3][0], p->pri[3][1]);
                        } else {
                            g0[i] = gs[e] - (k0[i] - p->pri[2][0]) / (p->pri[2][1] * p->pri[2][1]);
                            g1[i] = gq[e] - n_of_e[e] - (k1[i] - p->pri[3][0]) / (p->pri[3][1] * p->pri[3][1]);
                            lp_par += npdf(k0[i], p->pri[2][0], p->pri[2][1]) + npdf(k1[i], p->pri[3][0], p->pri[3][1]);
                        }
                    }
                }
            }
#pragma omp critical
            for (int t = 0; t < T - 1; ++t) { U[t] += uloc[t]; Qn[t] += qloc[t]; }
        }
        lp += lp_par;
        for (int t = 0; t < T - 1; ++t) {
            g_st[(size_t)r * (T - 1) + t] = -U[t] - (str[t] - p->pri[0][0]) / (p->pri[0][1] * p->pri[0][1]);
            g_lst[(size_t)r * (T - 1) + t] = Qn[t] - N - (lstr[t] - p->pri[1][0]) / (p->pri[1][1] * p->pri[1][1]);
            lp += npdf(str[t], p->pri[0][0], p->pri[0][1]) + npdf(lstr[t], p->pri[1][0], p->pri[1][1]);
        }
        double GL[MAXT];
        for (int t = 0; t < T; ++t) GL[t] = ((t > 0 ? U[t - 1] : 0.0) - (t < T - 1 ? U[t] : 0.0)) / Lam[t];
#pragma omp parallel for schedule(static)
        for (int b = 0; b < B; ++b)
            for (int t = 0; t < T; ++t) g_llr[(size_t)b * T + t] += exp(llr[(size_t)b * T + t]) * GL[t];
    }
    /* hyper prior (s_bc_prior on theta) */
    for (long h = 0; h < L.H; ++h) {
        g_th[h] -= (th[h] - p->pri[2][0]) / (p->pri[2][1] * p->pri[2][1]);
        lp += npdf(th[h], p->pri[2][0], p->pri[2][1]);
    }
    return lp;
}

static double lgamma_const(const gport_problem *p) {
    const size_t n = (size_t)p->T * (p->N + p->M) * p->R;
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t i = 0; i < n; ++i) s += lgamma((double)p->counts[i] + 1.0);
    return s;
}

/* log pi(z_k) and gradient for K caller-supplied latent vectors z[K][D] */
void gport_logjoint_grad(const gport_problem *p, const double *z, int K, double *logp, double *grad) {
    const long D = gport_n_latent(p);
    const double lg = lgamma_const(p);
    for (int k = 0; k < K; ++k) logp[k] = logjoint_grad_one(p, z + (size_t)k * D, grad + (size_t)k * D, lg);
}

/* ELBO and gradient of +ELBO wrt (mu, omega) for caller-supplied eps[K][D]; logp[K] optional */
double gport_elbo_grad(const gport_problem *p, const double *mu, const double *omega, const double *eps, int K,
                       double *grad, double *logp) {
    const long D = gport_n_latent(p);
    double *z = (double *)malloc(sizeof(double) * D), *g = (double *)malloc(sizeof(double) * D);
    double *sig = (double *)malloc(sizeof(double) * D);
    const double lg = lgamma_const(p);
    double elbo = 0.0, ent = 0.5 * D * (1.0 + LOG2PI);
    for (long i = 0; i < D; ++i) { sig[i] = softplus(omega[i]); ent += log(sig[i]); grad[i] = 0.0; grad[D + i] = 0.0; }
    for (int k = 0; k < K; ++k) {
        const double *e = eps + (size_t)k * D;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < D; ++i) z[i] = mu[i] + sig[i] * e[i];
        const double lp = logjoint_grad_one(p, z, g, lg);
        if (logp) logp[k] = lp;
        elbo += lp / K;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < D; ++i) { grad[i] += g[i] / K; grad[D + i] += g[i] * e[i] / K; }
    }
    for (long i = 0; i < D; ++i) grad[D + i] = (grad[D + i] + 1.0 / sig[i]) * sigmoid(omega[i]);
    free(z); free(g); free(sig);
    return elbo + ent;
}

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

/* n_steps of AdvancedVI.optimize! (DecayedADAGrad) with fresh eps per step: the timed CPU baseline */
double gport_advi_steps(const gport_problem *p, double *theta, double *acc, double eta, double pre, double post,
                        int K, int n_steps, uint64_t seed, long first_step) {
    const long D = gport_n_latent(p);
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
#pragma omp parallel for schedule(static)
        for (long j = 0; j < 2 * D; ++j) {
            double gr = j < D ? gm[j] : (gm[j] + 1.0 / sig[j - D]) * sigmoid(theta[j]);
            gr = -gr;
            acc[j] = post * acc[j] + pre * gr * gr;
            theta[j] -= eta * gr / (sqrt(acc[j]) + 1e-8);
        }
    }
    free(z); free(g); free(e); free(sig); free(gm);
    return elbo;
}
