"""CPU: the oracle against independent evaluations (PARITY UNPINNED by the reference -- these are the
cross-checks that stand in for golden vectors)."""
import math

import numpy as np
import pytest
import scipy.stats as st
import torch

from helpers import FIXTURES, load_fixture, oracle_problem, plausible_latents, plausible_theta, uneven_replicates
from oracle import advi_ref, model_ref


def _setup(bb, model, uneven=False):
    df, cols = load_fixture(model)
    if uneven:
        df = uneven_replicates(df)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc, da.n_env,
                              da.n_geno)
    return da, lay


def _scipy_fitness_normal(z, da):
    """Second, independent evaluation of model_fitness_normal.jl:131-271 with scipy.stats densities
    and explicit (t, b) loops -- no vectorised repeat/vec bookkeeping shared with the oracle."""
    R = np.asarray(da.bc_count)
    T, B = R.shape
    N, M = da.n_neutral, da.n_bc
    s_t, ls_t = z[:T - 1], z[T - 1:2 * (T - 1)]
    s_m, ls_m = z[2 * (T - 1):2 * (T - 1) + M], z[2 * (T - 1) + M:2 * (T - 1) + 2 * M]
    ll = z[2 * (T - 1) + 2 * M:]
    lp = st.norm(0, 2).logpdf(s_t).sum() + st.norm(0, 1).logpdf(ls_t).sum()
    lp += st.norm(0, 2).logpdf(s_m).sum() + st.norm(0, 1).logpdf(ls_m).sum() + st.norm(3, 3).logpdf(ll).sum()
    Lam = np.exp(ll).reshape(B, T).T
    F = Lam / Lam.sum(axis=1, keepdims=True)
    for t in range(T):
        lp += st.poisson(Lam[t].sum()).logpmf(R[t].sum())
        lp += st.multinomial(int(R[t].sum()), F[t]).logpmf(R[t])
    for b in range(B):
        for t in range(T - 1):
            gamma = math.log(F[t + 1, b] / F[t, b])
            if b < N:
                lp += st.norm(-s_t[t], math.exp(ls_t[t])).logpdf(gamma)
            else:
                lp += st.norm(s_m[b - N] - s_t[t], math.exp(ls_m[b - N])).logpdf(gamma)
    return lp


def test_fitness_normal_against_scipy(bb):
    da, lay = _setup(bb, "fitness_normal")
    rng = np.random.default_rng(0)
    z = plausible_latents(lay, da, rng, 2)
    prob = oracle_problem(da, "fitness_normal")
    for k in range(2):
        lp = float(model_ref.logjoint("fitness_normal", torch.tensor(z[k]), prob))
        ref = _scipy_fitness_normal(z[k], da)
        assert abs(lp - ref) <= 1e-9 * abs(ref)


@pytest.mark.parametrize("model", list(FIXTURES))
def test_collapsed_count_likelihood_identity(bb, model):
    """Poisson(n_t|Lambda_t) x Multinomial(r_t|n_t, f_t) == prod_b Poisson(r_tb|lambda_tb)
    exactly when n_t = sum_b r_tb (guaranteed by the packer, utils.jl:431-432)."""
    da, lay = _setup(bb, model)
    rng = np.random.default_rng(1)
    R = np.asarray(da.bc_count)
    blocks = [R] if R.ndim == 2 else [R[:, :, r] for r in range(R.shape[2])]
    for blk in blocks:
        T, B = blk.shape
        ll = torch.tensor(np.log(blk + 1.0) + 0.2 * rng.standard_normal((T, B)))
        Lam = torch.exp(ll)
        a = float(model_ref._count_terms_matrix(Lam, torch.tensor(blk), torch.tensor(blk.sum(axis=1))))
        b = float((torch.tensor(blk, dtype=torch.float64) * ll - Lam
                   - torch.lgamma(torch.tensor(blk, dtype=torch.float64) + 1)).sum())
        assert abs(a - b) <= 1e-10 * abs(b)


@pytest.mark.parametrize("model", list(FIXTURES))
def test_autograd_gradient_against_finite_differences(bb, model):
    da, lay = _setup(bb, model)
    rng = np.random.default_rng(2)
    z = plausible_latents(lay, da, rng, 1)[0]
    prob = oracle_problem(da, model)
    lp, g = model_ref.logjoint_and_grad(model, z, prob)
    idx = rng.choice(lay.n_latent, 12, replace=False)
    for i in idx:
        h = 1e-5 * max(1.0, abs(z[i]))
        zp, zm = z.copy(), z.copy()
        zp[i] += h
        zm[i] -= h
        fd = (model_ref.logjoint_and_grad(model, zp, prob)[0] - model_ref.logjoint_and_grad(model, zm, prob)[0]) / (2 * h)
        # central differences of a value ~1e6 in fp64: cancellation floor ~ |lp| * 1e-16 / h
        assert abs(fd - g[i]) <= 2e-5 * max(1.0, abs(g[i])) + 4 * abs(lp) * 2.2e-16 / h, (i, fd, g[i])


def test_latent_counts_match_survey_table(bb):
    """D = 103 / 236 / 177 / 114 on the four fixtures."""
    expect = {"fitness_normal": 103, "replicate_fitness_normal": 236, "multienv_fitness_normal": 177,
              "genotype_fitness_normal": 114}
    for model, D in expect.items():
        da, lay = _setup(bb, model)
        assert lay.n_latent == D == model_ref.n_latent(model, oracle_problem(da, model))


def test_ragged_replicates_quirk_and_correction_differ(bb):
    """M2v (replicates.jl:599-605): the as-written neutral pairing differs from the corrected one."""
    da, lay = _setup(bb, "replicate_fitness_normal", uneven=True)
    rng = np.random.default_rng(3)
    z = torch.tensor(plausible_latents(lay, da, rng, 1)[0])
    a = float(model_ref.logjoint("replicate_fitness_normal", z, oracle_problem(da, "replicate_fitness_normal", corrected=False)))
    b = float(model_ref.logjoint("replicate_fitness_normal", z, oracle_problem(da, "replicate_fitness_normal", corrected=True)))
    assert np.isfinite(a) and np.isfinite(b) and a != b


def test_multienv_replicate_reduces_to_replicate_when_one_env(bb):
    """M5 with a single environment is M2 (same terms, index maps composed)."""
    da, lay = _setup(bb, "replicate_fitness_normal")
    rng = np.random.default_rng(4)
    z = torch.tensor(plausible_latents(lay, da, rng, 1)[0])
    prob = oracle_problem(da, "replicate_fitness_normal")
    prob5 = dict(prob, envs=["e"] * np.asarray(da.bc_count).shape[0])
    a = float(model_ref.logjoint("replicate_fitness_normal", z, prob))
    b = float(model_ref.logjoint("multienv_replicate_fitness_normal", z, prob5))
    assert abs(a - b) <= 1e-10 * abs(a)


def _ragged_multienv_arrays(bb):
    """Replicate fixture, last time point of the last replicate dropped (test/vi_tests.jl:102-105), with an
    environment column: one environment list per replicate, of unequal length."""
    df, _ = load_fixture("replicate_fitness_normal")
    df = uneven_replicates(df.assign(env=df.time.map({1: "A", 2: "A", 3: "B", 4: "C", 5: "B"})))
    return bb.utils.data_to_arrays(df, rep_col="rep", env_col="env")


def test_ragged_multienv_replicate_model(bb):
    """M5, Vector{Matrix{Int64}} method (…hierarchical_replicates.jl:449-687): with equal T and one shared
    environment list it is the Array{Int64,3} method (:158-363) term for term."""
    da = _ragged_multienv_arrays(bb)
    assert isinstance(da.bc_count, list) and da.n_time == [5, 4]
    assert da.envs == [["A", "A", "B", "C", "B"], ["A", "A", "B", "C"]] and da.n_env == 3
    model = "multienv_replicate_fitness_normal"
    lay = bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc, da.n_env, 0)
    prob = oracle_problem(da, model)
    assert lay.n_latent == model_ref.n_latent(model, prob)
    rng = np.random.default_rng(8)
    z = torch.tensor(plausible_latents(lay, da, rng, 1)[0])
    assert np.isfinite(float(model_ref.logjoint(model, z, prob)))
    # equal T: the list-of-matrices method == the 3-D method
    df, _ = load_fixture("replicate_fitness_normal")
    da3 = bb.utils.data_to_arrays(df.assign(env=df.time.map({1: "A", 2: "A", 3: "B", 4: "C", 5: "B"})), rep_col="rep", env_col="env")
    R3 = np.asarray(da3.bc_count)
    lay3 = bb.model.var_groups(bb.model.resolve(model), da3.n_time, da3.n_rep, da3.n_neutral, da3.n_bc, da3.n_env, 0)
    z3 = torch.tensor(plausible_latents(lay3, da3, rng, 1)[0])
    prob3 = oracle_problem(da3, model)
    probv = dict(prob3, bc_count=[R3[:, :, r] for r in range(R3.shape[2])],
                 bc_total=[np.asarray(da3.bc_total)[:, r] for r in range(R3.shape[2])],
                 envs=[list(da3.envs)] * R3.shape[2])
    a, b = float(model_ref.logjoint(model, z3, prob3)), float(model_ref.logjoint(model, z3, probv))
    assert abs(a - b) <= 1e-12 * abs(a)


def test_elbo_gradient_formula(bb):
    """g_mu = mean_k dlogpi/dz, g_omega = (mean_k dlogpi/dz * eps + 1/sigma) * sigmoid(omega)
    equals autograd through the whole ELBO."""
    model = "fitness_normal"
    da, lay = _setup(bb, model)
    rng = np.random.default_rng(5)
    mu, om = plausible_theta(lay, da, rng)
    K = 3
    eps = rng.standard_normal((K, lay.n_latent))
    prob = oracle_problem(da, model)
    elbo, g_mu, g_om, logps = advi_ref.elbo_value_and_grad(model, prob, mu, om, eps)
    sig = advi_ref.softplus(om)
    gz = np.stack([model_ref.logjoint_and_grad(model, mu + sig * eps[k], prob)[1] for k in range(K)])
    assert np.allclose(g_mu, gz.mean(axis=0), rtol=1e-10, atol=1e-8)
    assert np.allclose(g_om, ((gz * eps).mean(axis=0) + 1 / sig) * advi_ref.sigmoid(om), rtol=1e-10, atol=1e-8)
    assert abs(elbo - (logps.mean() + advi_ref.entropy_diag_normal(sig))) <= 1e-9 * abs(elbo)


def test_optimizers_follow_advancedvi_rules():
    g = np.array([1.0, -2.0, 0.5])
    d = advi_ref.DecayedADAGrad(0.1, 1.0, 0.9)
    acc = 0.9 * 1e-8 + g ** 2
    assert np.allclose(d.apply(g), 0.1 * g / (np.sqrt(acc) + 1e-8))
    t = advi_ref.TruncatedADAGrad(0.1, 1.0, 2)
    d1 = t.apply(g)
    assert np.allclose(d1, 0.1 * g / (1.0 + np.abs(g) + 1e-8))
    t.apply(2 * g)
    d3 = t.apply(3 * g)                       # window of 2: slots hold (2g)^2 and (3g)^2
    assert np.allclose(d3, 0.1 * 3 * g / (1.0 + np.sqrt(4 * g ** 2 + 9 * g ** 2) + 1e-8))
