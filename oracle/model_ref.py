"""Oracle log-joints: torch.float64 restatement of the BarBay.jl model bodies.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

Every function here follows one ``Turing.@model`` body of the reference *as
written* -- same latent order, same ``reshape``/``vec``/``repeat`` orderings
(Julia is column-major and 1-based; the helpers below reproduce that), the
Poisson x Multinomial observation terms kept un-collapsed -- so that autograd
gives an independent gradient to hold the analytic CUDA gradient against.

Reference files restated (paths relative to /root/reference):
  src/model_fitness_normal.jl:120-272                         fitness_normal
  src/model_fitness_normal_hierarchical_replicates.jl:145-332 replicate_fitness_normal (Array{Int64,3})
  src/model_fitness_normal_hierarchical_replicates.jl:407-638 replicate_fitness_normal (Vector{Matrix})
  src/model_multienv_fitness_normal.jl:133-303                multienv_fitness_normal
  src/model_fitness_normal_hierarchical_genotypes.jl:151-330  genotype_fitness_normal
  src/model_multienv_fitness_normal_hierarchical_replicates.jl:158-363
                                                              multienv_replicate_fitness_normal (Array{Int64,3})
Densities follow Distributions.jl 0.25 (not vendored): MvNormal with diagonal
covariance, Poisson, Multinomial(check_args=false).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

F64 = torch.float64
LOG2PI = math.log(2.0 * math.pi)

DEFAULT_PRIORS = {
    "s_pop_prior": [0.0, 2.0],
    "logσ_pop_prior": [0.0, 1.0],
    "s_bc_prior": [0.0, 2.0],
    "logσ_bc_prior": [0.0, 1.0],
    "logλ_prior": [3.0, 3.0],
    "logτ_prior": [-2.0, 1.0],
}


# --------------------------------------------------------------------------
# Julia array-semantics helpers (column-major)
# --------------------------------------------------------------------------
def jl_reshape(v: torch.Tensor, *dims: int) -> torch.Tensor:
    """``reshape(v, dims...)`` of a Julia vector: first index fastest."""
    return v.reshape(*reversed(dims)).permute(*reversed(range(len(dims))))


def jl_vec(a: torch.Tensor) -> torch.Tensor:
    """``vec(A)`` / ``A[:]``: column-major flattening."""
    return a.permute(*reversed(range(a.dim()))).reshape(-1)


def jl_repeat(v: torch.Tensor, outer: int = 1, inner: int = 1) -> torch.Tensor:
    """``repeat(v, outer)`` / ``repeat(v, inner=n)`` for vectors."""
    return v.repeat_interleave(inner).repeat(outer)


def _indexin_unique(labels: Sequence) -> tuple[list, np.ndarray]:
    """``unique(x)`` (first appearance) and 0-based ``indexin(x, unique(x))``."""
    uniq: list = []
    pos: dict = {}
    idx = np.empty(len(labels), dtype=np.int64)
    for i, lab in enumerate(labels):
        if lab not in pos:
            pos[lab] = len(uniq)
            uniq.append(lab)
        idx[i] = pos[lab]
    return uniq, idx


# --------------------------------------------------------------------------
# Densities (Distributions.jl semantics)
# --------------------------------------------------------------------------
def mvnormal_diag_logpdf(x: torch.Tensor, mean: torch.Tensor, var: torch.Tensor) -> torch.Tensor:
    """logpdf(MvNormal(mean, Diagonal(var)), x)."""
    n = x.numel()
    return -0.5 * (n * LOG2PI + torch.log(var).sum()) - 0.5 * (((x - mean) ** 2) / var).sum()


def poisson_logpdf(n: torch.Tensor, lam: torch.Tensor) -> torch.Tensor:
    """logpdf(Poisson(lam), n) = xlogy(n, lam) - lam - loggamma(n+1)."""
    nf = n.to(F64)
    return torch.xlogy(nf, lam) - lam - torch.lgamma(nf + 1.0)


def multinomial_logpdf(r: torch.Tensor, n: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """logpdf(Multinomial(n, p; check_args=false), r), one row."""
    rf = r.to(F64)
    if int(r.sum()) != int(n):
        return torch.tensor(-math.inf, dtype=F64)
    return torch.lgamma(n.to(F64) + 1.0) - torch.lgamma(rf + 1.0).sum() + torch.xlogy(rf, p).sum()


def _prior_mean_var(prior, n: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Vector prior [mean, std] -> repeated; Matrix prior n x 2 -> columns."""
    p = np.asarray(prior, dtype=np.float64)
    if p.ndim == 1:
        mean = torch.full((n,), float(p[0]), dtype=F64)
        var = torch.full((n,), float(p[1]) ** 2, dtype=F64)
    else:
        if p.shape != (n, 2):
            raise ValueError(f"matrix prior must be {n}x2, got {p.shape}")
        mean = torch.from_numpy(p[:, 0].copy())
        var = torch.from_numpy(p[:, 1].copy()) ** 2
    return mean, var


def _prior(z: torch.Tensor, prior, n: int) -> torch.Tensor:
    mean, var = _prior_mean_var(prior, n)
    return mvnormal_diag_logpdf(z, mean, var)


class _Cursor:
    """Walks the flat latent vector in VarInfo order."""

    def __init__(self, z: torch.Tensor):
        self.z = z
        self.pos = 0
        self.ranges: list[tuple[str, int, int]] = []

    def take(self, name: str, n: int) -> torch.Tensor:
        out = self.z[self.pos:self.pos + n]
        self.ranges.append((name, self.pos, self.pos + n))
        self.pos += n
        return out

    def done(self):
        if self.pos != self.z.numel():
            raise ValueError(f"latent vector has {self.z.numel()} entries, model consumed {self.pos}")


def _pri(priors: dict | None) -> dict:
    out = dict(DEFAULT_PRIORS)
    if priors:
        out.update(priors)
    return out


def _count_terms_matrix(Lam: torch.Tensor, R: torch.Tensor, nt: torch.Tensor) -> torch.Tensor:
    """Poisson(n_t | sum Lam_t) + sum_t Multinomial(R_t | n_t, F_t) for a T x B block."""
    F = Lam / Lam.sum(dim=1, keepdim=True)
    lp = poisson_logpdf(nt, Lam.sum(dim=1)).sum()
    for t in range(R.shape[0]):
        lp = lp + multinomial_logpdf(R[t], nt[t], F[t])
    return lp


# --------------------------------------------------------------------------
# M1  fitness_normal  (model_fitness_normal.jl:120-272)
# --------------------------------------------------------------------------
def logjoint_fitness_normal(z, R, nt, n_neutral, n_bc, priors=None):
    pr = _pri(priors)
    R = torch.as_tensor(np.asarray(R), dtype=torch.int64)
    nt = torch.as_tensor(np.asarray(nt), dtype=torch.int64)
    T, B = R.shape
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", T - 1)                       # :137-146
    lsig_t = cur.take("logσ̲ₜ", T - 1)                  # :149-160
    s_m = cur.take("s̲⁽ᵐ⁾", n_bc)                      # :165-174
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_bc)                 # :178-189
    logLam = cur.take("logΛ̲̲", T * B)                   # :194-203
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], T - 1)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], T - 1)
    lp = lp + _prior(s_m, pr["s_bc_prior"], n_bc)
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_bc)
    lp = lp + _prior(logLam, pr["logλ_prior"], T * B)

    Lam = jl_reshape(torch.exp(logLam), T, B)           # :209
    F = Lam / Lam.sum(dim=1, keepdim=True)              # :212
    logG = torch.log(F[1:, :] / F[:-1, :])              # :215
    logG_n = jl_vec(logG[:, :n_neutral])                # :218
    logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc])  # :219
    lp = lp + _count_terms_matrix(Lam, R, nt)           # :224-244
    # neutrals :251-257
    lp = lp + mvnormal_diag_logpdf(
        logG_n, jl_repeat(-s_t, outer=n_neutral), jl_repeat(torch.exp(lsig_t) ** 2, outer=n_neutral))
    # mutants :262-270
    lp = lp + mvnormal_diag_logpdf(
        logG_m,
        jl_repeat(s_m, inner=T - 1) - jl_repeat(s_t, outer=n_bc),
        jl_repeat(torch.exp(lsig_m) ** 2, inner=T - 1))
    return lp


# --------------------------------------------------------------------------
# M2  replicate_fitness_normal, Array{Int64,3} method (replicates.jl:145-332)
# --------------------------------------------------------------------------
def logjoint_replicate_fitness_normal(z, R, nt, n_neutral, n_bc, priors=None):
    pr = _pri(priors)
    R = torch.as_tensor(np.asarray(R), dtype=torch.int64)        # T x B x Rep
    nt = torch.as_tensor(np.asarray(nt), dtype=torch.int64)      # T x Rep
    T, B, n_rep = R.shape
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", (T - 1) * n_rep)
    lsig_t = cur.take("logσ̲ₜ", (T - 1) * n_rep)
    theta = cur.take("θ̲⁽ᵐ⁾", n_bc)
    theta_tilde = cur.take("θ̲̃⁽ᵐ⁾", n_bc * n_rep)
    ltau = cur.take("logτ̲⁽ᵐ⁾", n_bc * n_rep)
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_bc * n_rep)
    logLam = cur.take("logΛ̲̲", T * B * n_rep)
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], (T - 1) * n_rep)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], (T - 1) * n_rep)
    lp = lp + _prior(theta, pr["s_bc_prior"], n_bc)
    lp = lp + _prior(theta_tilde, [0.0, 1.0], n_bc * n_rep)                 # :205-207
    lp = lp + _prior(ltau, pr["logτ_prior"], n_bc * n_rep)                  # :210-213
    s_m = jl_repeat(theta, outer=n_rep) + torch.exp(ltau) * theta_tilde     # :216
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_bc * n_rep)
    lp = lp + _prior(logLam, pr["logλ_prior"], T * B * n_rep)

    Lam = jl_reshape(torch.exp(logLam), T, B, n_rep)                        # :248
    F = Lam / Lam.sum(dim=1, keepdim=True)                                  # :251
    logG = torch.log(F[1:, :, :] / F[:-1, :, :])                            # :254
    logG_n = jl_vec(logG[:, :n_neutral, :])                                 # :258
    logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc, :])                 # :259
    for r in range(n_rep):                                                  # :264-291
        lp = lp + _count_terms_matrix(Lam[:, :, r], R[:, :, r], nt[:, r])
    s_t2 = jl_reshape(s_t, T - 1, n_rep)                                    # :296
    lsig_t2 = jl_reshape(lsig_t, T - 1, n_rep)                              # :297
    # neutrals :304-316  -vec(repeat(s_t, inner=(1, n_neutral)))
    lp = lp + mvnormal_diag_logpdf(
        logG_n,
        -jl_vec(s_t2.repeat_interleave(n_neutral, dim=1)),
        jl_vec((torch.exp(lsig_t2) ** 2).repeat_interleave(n_neutral, dim=1)))
    # mutants :321-330
    lp = lp + mvnormal_diag_logpdf(
        logG_m,
        jl_repeat(s_m, inner=T - 1) - jl_vec(s_t2.repeat_interleave(n_bc, dim=1)),
        jl_repeat(torch.exp(lsig_m) ** 2, inner=T - 1))
    return lp


# --------------------------------------------------------------------------
# M2v replicate_fitness_normal, Vector{Matrix} method (replicates.jl:407-638)
# --------------------------------------------------------------------------
def logjoint_replicate_fitness_normal_ragged(z, R_list, nt_list, n_neutral, n_bc, priors=None,
                                             corrected: bool = False):
    """Unequal number of time points per replicate.

    ``corrected=False`` reproduces the reference as written: the neutral mean and
    variance are built with ``repeat(..., inner=n_neutral)`` (:599, :603-605)
    while ``vec(logΓⁿ)`` is time-fastest (:549), so the k-th neutral ratio is
    paired with s̄[ceil(k / N)].  ``corrected=True`` pairs
    ratio (t, n) with s̄[t] like the equal-T method does.
    """
    pr = _pri(priors)
    Rs = [torch.as_tensor(np.asarray(r), dtype=torch.int64) for r in R_list]
    nts = [torch.as_tensor(np.asarray(n), dtype=torch.int64) for n in nt_list]
    n_rep = len(Rs)
    n_time = [int(r.shape[0]) for r in Rs]
    B = n_neutral + n_bc
    n_st = sum(t - 1 for t in n_time)
    n_lam = sum(int(r.numel()) for r in Rs)
    # rep_ranges / time_ranges :426-447
    rep_ranges, time_ranges = [], []
    a = b = 0
    for r in range(n_rep):
        rep_ranges.append((a, a + Rs[r].numel()))
        a += Rs[r].numel()
        time_ranges.append((b, b + n_time[r] - 1))
        b += n_time[r] - 1
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", n_st)
    lsig_t = cur.take("logσ̲ₜ", n_st)
    theta = cur.take("θ̲⁽ᵐ⁾", n_bc)
    theta_tilde = cur.take("θ̲̃⁽ᵐ⁾", n_bc * n_rep)
    ltau = cur.take("logτ̲⁽ᵐ⁾", n_bc * n_rep)
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_bc * n_rep)
    logLam = cur.take("logΛ̲̲", n_lam)
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], n_st)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], n_st)
    lp = lp + _prior(theta, pr["s_bc_prior"], n_bc)
    lp = lp + _prior(theta_tilde, [0.0, 1.0], n_bc * n_rep)
    lp = lp + _prior(ltau, pr["logτ_prior"], n_bc * n_rep)
    s_m = jl_repeat(theta, outer=n_rep) + torch.exp(ltau) * theta_tilde       # :498
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_bc * n_rep)
    lp = lp + _prior(logLam, pr["logλ_prior"], n_lam)

    Lam = [jl_reshape(torch.exp(logLam)[lo:hi], n_time[r], B)                 # :533-536
           for r, (lo, hi) in enumerate(rep_ranges)]
    s_m2 = jl_reshape(s_m, n_bc, n_rep)                                       # :553
    lsig_m2 = jl_reshape(lsig_m, n_bc, n_rep)                                 # :554
    for r in range(n_rep):
        lp = lp + _count_terms_matrix(Lam[r], Rs[r], nts[r])                  # :559-585
    for r in range(n_rep):
        F = Lam[r] / Lam[r].sum(dim=1, keepdim=True)                          # :539
        logG = torch.log(F[1:, :] / F[:-1, :])                                # :542
        logG_n = jl_vec(logG[:, :n_neutral])                                  # :549
        logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc])                  # :550
        lo, hi = time_ranges[r]
        st_r, ls_r = s_t[lo:hi], lsig_t[lo:hi]
        if corrected:
            mean_n = -jl_repeat(st_r, outer=n_neutral)
            var_n = jl_repeat(torch.exp(ls_r) ** 2, outer=n_neutral)
        else:
            mean_n = -jl_repeat(st_r, inner=n_neutral)                        # :599
            var_n = jl_repeat(torch.exp(ls_r) ** 2, inner=n_neutral)          # :603-605
        lp = lp + mvnormal_diag_logpdf(logG_n, mean_n, var_n)
        # mutants :615-634: reduce(vcat, [s - s_t[range] for s in s_m[:, rep]])
        mean_m = jl_repeat(s_m2[:, r], inner=n_time[r] - 1) - jl_repeat(st_r, outer=n_bc)
        var_m = jl_repeat(torch.exp(lsig_m2[:, r]) ** 2, inner=n_time[r] - 1)
        lp = lp + mvnormal_diag_logpdf(logG_m, mean_m, var_m)
    return lp


# --------------------------------------------------------------------------
# M3  multienv_fitness_normal (model_multienv_fitness_normal.jl:133-303)
# --------------------------------------------------------------------------
def logjoint_multienv_fitness_normal(z, R, nt, n_neutral, n_bc, envs, priors=None):
    pr = _pri(priors)
    R = torch.as_tensor(np.asarray(R), dtype=torch.int64)
    nt = torch.as_tensor(np.asarray(nt), dtype=torch.int64)
    T, B = R.shape
    if T != len(envs):
        raise ValueError("Number of time points must match list of of environments")  # :146-148
    env_unique, env_idx = _indexin_unique(list(envs))                                    # :151-155
    n_env = len(env_unique)
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", T - 1)
    lsig_t = cur.take("logσ̲ₜ", T - 1)
    s_m = cur.take("s̲⁽ᵐ⁾", n_bc * n_env)
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_bc * n_env)
    logLam = cur.take("logΛ̲̲", T * B)
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], T - 1)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], T - 1)
    lp = lp + _prior(s_m, pr["s_bc_prior"], n_bc * n_env)
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_bc * n_env)
    lp = lp + _prior(logLam, pr["logλ_prior"], T * B)
    Lam = jl_reshape(torch.exp(logLam), T, B)
    F = Lam / Lam.sum(dim=1, keepdim=True)
    logG = torch.log(F[1:, :] / F[:-1, :])
    logG_n = jl_vec(logG[:, :n_neutral])
    logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc])
    lp = lp + _count_terms_matrix(Lam, R, nt)
    s_m2 = jl_reshape(s_m, n_env, n_bc)                 # :271
    lsig_m2 = jl_reshape(lsig_m, n_env, n_bc)           # :272
    lp = lp + mvnormal_diag_logpdf(                     # :279-285
        logG_n, jl_repeat(-s_t, outer=n_neutral), jl_repeat(torch.exp(lsig_t) ** 2, outer=n_neutral))
    rows = torch.as_tensor(env_idx[1:], dtype=torch.int64)
    lp = lp + mvnormal_diag_logpdf(                     # :293-301
        logG_m,
        jl_vec(s_m2[rows, :]) - jl_repeat(s_t, outer=n_bc),
        jl_vec(torch.exp(lsig_m2[rows, :]) ** 2))
    return lp


# --------------------------------------------------------------------------
# M4  genotype_fitness_normal (model_fitness_normal_hierarchical_genotypes.jl:151-330)
# --------------------------------------------------------------------------
def logjoint_genotype_fitness_normal(z, R, nt, n_neutral, n_bc, genotypes, priors=None):
    pr = _pri(priors)
    R = torch.as_tensor(np.asarray(R), dtype=torch.int64)
    nt = torch.as_tensor(np.asarray(nt), dtype=torch.int64)
    T, B = R.shape
    if n_bc != len(genotypes):
        raise ValueError("List of genotypes must match number of barcodes")  # :165-167
    geno_unique, geno_idx = _indexin_unique(list(genotypes))                   # :170-174
    n_geno = len(geno_unique)
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", T - 1)
    lsig_t = cur.take("logσ̲ₜ", T - 1)
    theta = cur.take("θ̲⁽ᵐ⁾", n_geno)
    theta_tilde = cur.take("θ̲̃⁽ᵐ⁾", n_bc)
    ltau = cur.take("logτ̲⁽ᵐ⁾", n_bc)
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_bc)
    logLam = cur.take("logΛ̲̲", T * B)
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], T - 1)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], T - 1)
    lp = lp + _prior(theta, pr["s_bc_prior"], n_geno)
    lp = lp + _prior(theta_tilde, [0.0, 1.0], n_bc)                            # :221
    lp = lp + _prior(ltau, pr["logτ_prior"], n_bc)                             # :224-227
    gi = torch.as_tensor(geno_idx, dtype=torch.int64)
    s_m = theta[gi] + torch.exp(ltau) * theta_tilde                            # :230
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_bc)
    lp = lp + _prior(logLam, pr["logλ_prior"], T * B)
    Lam = jl_reshape(torch.exp(logLam), T, B)
    F = Lam / Lam.sum(dim=1, keepdim=True)
    logG = torch.log(F[1:, :] / F[:-1, :])
    logG_n = jl_vec(logG[:, :n_neutral])
    logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc])
    lp = lp + _count_terms_matrix(Lam, R, nt)
    lp = lp + mvnormal_diag_logpdf(                                            # :305-311
        logG_n, jl_repeat(-s_t, outer=n_neutral), jl_repeat(torch.exp(lsig_t) ** 2, outer=n_neutral))
    lp = lp + mvnormal_diag_logpdf(                                            # :316-328
        logG_m,
        jl_repeat(s_m, inner=T - 1) - jl_repeat(s_t, outer=n_bc),
        jl_repeat(torch.exp(lsig_m) ** 2, inner=T - 1))
    return lp


# --------------------------------------------------------------------------
# M5  multienv_replicate_fitness_normal, Array{Int64,3} method
#     (model_multienv_fitness_normal_hierarchical_replicates.jl:158-363)
# --------------------------------------------------------------------------
def logjoint_multienv_replicate_fitness_normal(z, R, nt, n_neutral, n_bc, envs, priors=None):
    pr = _pri(priors)
    R = torch.as_tensor(np.asarray(R), dtype=torch.int64)       # T x B x Rep
    nt = torch.as_tensor(np.asarray(nt), dtype=torch.int64)     # T x Rep
    T, B, n_rep = R.shape
    if T != len(envs):
        raise ValueError("Number of time points must match list of of environments")  # :172-174
    env_unique, env_idx = _indexin_unique(list(envs))
    n_env = len(env_unique)
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", (T - 1) * n_rep)
    lsig_t = cur.take("logσ̲ₜ", (T - 1) * n_rep)
    theta = cur.take("θ̲⁽ᵐ⁾", n_env * n_bc)
    theta_tilde = cur.take("θ̲̃⁽ᵐ⁾", n_env * n_bc * n_rep)
    ltau = cur.take("logτ̲⁽ᵐ⁾", n_env * n_bc * n_rep)
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_env * n_bc * n_rep)
    logLam = cur.take("logΛ̲̲", T * B * n_rep)
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], (T - 1) * n_rep)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], (T - 1) * n_rep)
    lp = lp + _prior(theta, pr["s_bc_prior"], n_env * n_bc)
    lp = lp + _prior(theta_tilde, [0.0, 1.0], n_env * n_bc * n_rep)
    lp = lp + _prior(ltau, pr["logτ_prior"], n_env * n_bc * n_rep)
    s_m = jl_repeat(theta, outer=n_rep) + torch.exp(ltau) * theta_tilde        # :248
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_env * n_bc * n_rep)
    lp = lp + _prior(logLam, pr["logλ_prior"], T * B * n_rep)
    Lam = jl_reshape(torch.exp(logLam), T, B, n_rep)
    F = Lam / Lam.sum(dim=1, keepdim=True)
    logG = torch.log(F[1:, :, :] / F[:-1, :, :])
    logG_n = jl_vec(logG[:, :n_neutral, :])
    logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc, :])
    for r in range(n_rep):
        lp = lp + _count_terms_matrix(Lam[:, :, r], R[:, :, r], nt[:, r])
    s_t2 = jl_reshape(s_t, T - 1, n_rep)                                       # :320
    lsig_t2 = jl_reshape(lsig_t, T - 1, n_rep)                                 # :321
    s_m3 = jl_reshape(s_m, n_env, n_bc, n_rep)                                 # :325
    lsig_m3 = jl_reshape(lsig_m, n_env, n_bc, n_rep)                           # :326
    # neutrals :333-343  reduce(vcat, repeat.(eachcol(s_t), n_neutral))
    mean_n = -torch.cat([jl_repeat(s_t2[:, r], outer=n_neutral) for r in range(n_rep)])
    var_n = torch.cat([jl_repeat(torch.exp(lsig_t2[:, r]) ** 2, outer=n_neutral) for r in range(n_rep)])
    lp = lp + mvnormal_diag_logpdf(logG_n, mean_n, var_n)
    rows = torch.as_tensor(env_idx[1:], dtype=torch.int64)
    mean_m = jl_vec(s_m3[rows, :, :]) - torch.cat(
        [jl_repeat(s_t2[:, r], outer=n_bc) for r in range(n_rep)])             # :349-354
    var_m = jl_vec(torch.exp(lsig_m3[rows, :, :])) ** 2                        # :355-357
    lp = lp + mvnormal_diag_logpdf(logG_m, mean_m, var_m)
    return lp


# --------------------------------------------------------------------------
# M5v multienv_replicate_fitness_normal, Vector{Matrix{Int64}} method: unequal T
#     per replicate, one environment list per replicate
#     (model_multienv_fitness_normal_hierarchical_replicates.jl:449-687)
# --------------------------------------------------------------------------
def logjoint_multienv_replicate_fitness_normal_ragged(z, R_list, nt_list, n_neutral, n_bc, envs, priors=None):
    pr = _pri(priors)
    Rs = [torch.as_tensor(np.asarray(r), dtype=torch.int64) for r in R_list]
    nts = [torch.as_tensor(np.asarray(n), dtype=torch.int64) for n in nt_list]
    n_rep = len(Rs)
    n_time = [int(r.shape[0]) for r in Rs]
    if any(n_time[r] != len(envs[r]) for r in range(n_rep)):                             # :463-465
        raise ValueError("Number of time points must match list of of environments for all replicates")
    env_unique = []                                                                      # unique(vcat(envs...)) :468
    for es in envs:
        for e in es:
            if e not in env_unique:
                env_unique.append(e)
    n_env = len(env_unique)
    env_idx = [np.asarray([env_unique.index(e) for e in es], dtype=np.int64) for es in envs]   # :472 (0-based here)
    B = n_neutral + n_bc
    n_st = sum(t - 1 for t in n_time)
    n_lam = sum(int(r.numel()) for r in Rs)
    rep_ranges, time_ranges = [], []                                                     # :478-495
    a = b = 0
    for r in range(n_rep):
        rep_ranges.append((a, a + Rs[r].numel()))
        a += Rs[r].numel()
        time_ranges.append((b, b + n_time[r] - 1))
        b += n_time[r] - 1
    cur = _Cursor(z)
    s_t = cur.take("s̲ₜ", n_st)
    lsig_t = cur.take("logσ̲ₜ", n_st)
    theta = cur.take("θ̲⁽ᵐ⁾", n_env * n_bc)
    theta_tilde = cur.take("θ̲̃⁽ᵐ⁾", n_env * n_bc * n_rep)
    ltau = cur.take("logτ̲⁽ᵐ⁾", n_env * n_bc * n_rep)
    lsig_m = cur.take("logσ̲⁽ᵐ⁾", n_env * n_bc * n_rep)
    logLam = cur.take("logΛ̲̲", n_lam)
    cur.done()
    lp = _prior(s_t, pr["s_pop_prior"], n_st)
    lp = lp + _prior(lsig_t, pr["logσ_pop_prior"], n_st)
    lp = lp + _prior(theta, pr["s_bc_prior"], n_env * n_bc)
    lp = lp + _prior(theta_tilde, [0.0, 1.0], n_env * n_bc * n_rep)
    lp = lp + _prior(ltau, pr["logτ_prior"], n_env * n_bc * n_rep)
    s_m = jl_repeat(theta, outer=n_rep) + torch.exp(ltau) * theta_tilde                  # :557
    lp = lp + _prior(lsig_m, pr["logσ_bc_prior"], n_env * n_bc * n_rep)
    lp = lp + _prior(logLam, pr["logλ_prior"], n_lam)
    Lam = [jl_reshape(torch.exp(logLam)[lo:hi], n_time[r], B) for r, (lo, hi) in enumerate(rep_ranges)]   # :592-595
    s_m3 = jl_reshape(s_m, n_env, n_bc, n_rep)                                           # :609
    lsig_m3 = jl_reshape(lsig_m, n_env, n_bc, n_rep)                                     # :610
    for r in range(n_rep):
        lp = lp + _count_terms_matrix(Lam[r], Rs[r], nts[r])                             # :615-640
    for r in range(n_rep):
        F = Lam[r] / Lam[r].sum(dim=1, keepdim=True)                                     # :598
        logG = torch.log(F[1:, :] / F[:-1, :])                                           # :601
        logG_n = jl_vec(logG[:, :n_neutral])                                             # :605
        logG_m = jl_vec(logG[:, n_neutral:n_neutral + n_bc])                             # :606
        lo, hi = time_ranges[r]
        st_r, ls_r = s_t[lo:hi], lsig_t[lo:hi]
        # neutrals :650-668: repeat(s_t[range], n_neutral) -- outer repeat, pairs ratio (t, n) with s_t[t]
        lp = lp + mvnormal_diag_logpdf(logG_n, -jl_repeat(st_r, outer=n_neutral),
                                       jl_repeat(torch.exp(ls_r) ** 2, outer=n_neutral))
        # mutants :670-684: s_m[env_idx[rep][2:end], :, rep][:] .- repeat(s_t[range], n_bc)
        rows = torch.as_tensor(env_idx[r][1:], dtype=torch.int64)
        mean_m = jl_vec(s_m3[rows, :, r]) - jl_repeat(st_r, outer=n_bc)
        var_m = jl_vec(torch.exp(lsig_m3[rows, :, r])) ** 2
        lp = lp + mvnormal_diag_logpdf(logG_m, mean_m, var_m)
    return lp


# --------------------------------------------------------------------------
# Uniform entry: log-joint value and autograd gradient for a problem dict
# --------------------------------------------------------------------------
MODELS = (
    "fitness_normal",
    "replicate_fitness_normal",
    "multienv_fitness_normal",
    "genotype_fitness_normal",
    "multienv_replicate_fitness_normal",
)


def logjoint(model: str, z: torch.Tensor, prob: dict) -> torch.Tensor:
    """Dispatch on the reference model name.  ``prob`` carries bc_count, bc_total,
    n_neutral, n_bc (+ envs / genotypes) and ``priors``."""
    R, nt = prob["bc_count"], prob["bc_total"]
    N, M = prob["n_neutral"], prob["n_bc"]
    pri = prob.get("priors")
    if model == "fitness_normal":
        return logjoint_fitness_normal(z, R, nt, N, M, pri)
    if model == "replicate_fitness_normal":
        if isinstance(R, (list, tuple)):
            return logjoint_replicate_fitness_normal_ragged(
                z, R, nt, N, M, pri, corrected=bool(prob.get("corrected", False)))
        return logjoint_replicate_fitness_normal(z, R, nt, N, M, pri)
    if model == "multienv_fitness_normal":
        return logjoint_multienv_fitness_normal(z, R, nt, N, M, prob["envs"], pri)
    if model == "genotype_fitness_normal":
        return logjoint_genotype_fitness_normal(z, R, nt, N, M, prob["genotypes"], pri)
    if model == "multienv_replicate_fitness_normal":
        if isinstance(R, (list, tuple)):
            envs = prob["envs"]
            if not isinstance(envs[0], (list, tuple)):
                envs = [list(envs)] * len(R)
            return logjoint_multienv_replicate_fitness_normal_ragged(z, R, nt, N, M, envs, pri)
        return logjoint_multienv_replicate_fitness_normal(z, R, nt, N, M, prob["envs"], pri)
    raise ValueError(f"unknown model {model!r}")


def n_latent(model: str, prob: dict) -> int:
    R = prob["bc_count"]
    N, M = prob["n_neutral"], prob["n_bc"]
    B = N + M
    if isinstance(R, (list, tuple)):
        n_rep = len(R)
        n_st = sum(int(np.asarray(r).shape[0]) - 1 for r in R)
        n_lam = sum(int(np.asarray(r).size) for r in R)
        if model == "multienv_replicate_fitness_normal":
            envs = prob["envs"]
            flat = [e for es in envs for e in es] if isinstance(envs[0], (list, tuple)) else list(envs)
            E = len(_indexin_unique(flat)[0])
            return 2 * n_st + E * M + 3 * E * M * n_rep + n_lam
        return 2 * n_st + M + 3 * M * n_rep + n_lam
    R = np.asarray(R)
    T = R.shape[0]
    n_rep = R.shape[2] if R.ndim == 3 else 1
    if model == "fitness_normal":
        return 2 * (T - 1) + 2 * M + T * B
    if model == "replicate_fitness_normal":
        return 2 * (T - 1) * n_rep + M + 3 * M * n_rep + T * B * n_rep
    if model == "multienv_fitness_normal":
        E = len(_indexin_unique(list(prob["envs"]))[0])
        return 2 * (T - 1) + 2 * M * E + T * B
    if model == "genotype_fitness_normal":
        G = len(_indexin_unique(list(prob["genotypes"]))[0])
        return 2 * (T - 1) + G + 3 * M + T * B
    if model == "multienv_replicate_fitness_normal":
        E = len(_indexin_unique(list(prob["envs"]))[0])
        return 2 * (T - 1) * n_rep + E * M + 3 * E * M * n_rep + T * B * n_rep
    raise ValueError(model)


def logjoint_and_grad(model: str, z: np.ndarray, prob: dict) -> tuple[float, np.ndarray]:
    """log pi(z) and d log pi / dz by reverse-mode autograd (fp64)."""
    zt = torch.tensor(np.asarray(z, dtype=np.float64), dtype=F64, requires_grad=True)
    lp = logjoint(model, zt, prob)
    (g,) = torch.autograd.grad(lp, zt)
    return float(lp.detach()), g.numpy().copy()
