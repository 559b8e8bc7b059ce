"""CPU: the compiled C/OpenMP port (cpu_baseline of bench.py) against the torch oracle."""
import numpy as np

from helpers import load_fixture, oracle_problem, plausible_theta, rel_err
from oracle import advi_ref, cport


def test_c_port_matches_torch_oracle(bb):
    model = "fitness_normal"
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc)
    rng = np.random.default_rng(1)
    mu, om = plausible_theta(lay, da, rng)
    eps = rng.standard_normal((3, lay.n_latent))
    e_ref, gm, go, lps = advi_ref.elbo_value_and_grad(model, oracle_problem(da, model), mu, om, eps)
    pp = cport.PortProblem(da.bc_count, da.n_neutral, da.n_bc)
    e, gm2, go2, lp2 = pp.elbo_grad(mu, om, eps)
    assert abs(e - e_ref) <= 1e-11 * abs(e_ref)
    assert rel_err(gm2, gm) < 1e-11 and rel_err(go2, go) < 1e-11 and rel_err(lp2, lps) < 1e-11


def test_c_port_advi_steps_improve_elbo(bb):
    _, da, _ = bb.synth.config(2, scale=0.002)
    pp = cport.PortProblem(da.bc_count, da.n_neutral, da.n_bc)
    rng = np.random.default_rng(0)
    theta = np.concatenate([rng.standard_normal(pp.D), rng.standard_normal(pp.D)])
    acc = np.full(2 * pp.D, 1e-8)
    e0 = pp.advi_steps(theta, acc, 1, 4)
    e1 = pp.advi_steps(theta, acc, 200, 4, first_step=1)
    assert np.isfinite(e1) and e1 > e0


import pytest

ALL_MODELS = ["fitness_normal", "replicate_fitness_normal", "multienv_fitness_normal", "genotype_fitness_normal"]


def _model_port(da, model, priors=None):
    return cport.ModelPort(model, da.bc_count, da.n_neutral, da.n_bc, envs=da.envs, genotypes=da.genotypes,
                           priors=priors)


@pytest.mark.parametrize("model", ALL_MODELS)
def test_model_port_matches_torch_oracle_on_fixtures(bb, model):
    """The all-families C port (full-size parity checker and CPU baseline of cfg3-5) against the torch
    transliteration: ELBO, its gradient and the per-sample log-joints, reference fixtures."""
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc,
                              da.n_env, da.n_geno)
    rng = np.random.default_rng(2)
    mu, om = plausible_theta(lay, da, rng)
    eps = rng.standard_normal((3, lay.n_latent))
    e_ref, gm, go, lps = advi_ref.elbo_value_and_grad(model, oracle_problem(da, model), mu, om, eps)
    pp = _model_port(da, model)
    assert pp.D == lay.n_latent
    e, gm2, go2, lp2 = pp.elbo_grad(mu, om, eps)
    assert abs(e - e_ref) <= 1e-11 * abs(e_ref)
    assert rel_err(lp2, lps) < 1e-11
    assert rel_err(gm2, gm) < 1e-10 and rel_err(go2, go) < 1e-10


def test_model_port_multienv_replicate_and_uneven_groups(bb):
    """Fifth family (multienv x replicate) and genotype groups of uneven size, synthetic, non-default priors."""
    from oracle import model_ref
    rng = np.random.default_rng(5)
    priors = {"s_pop_prior": [0.1, 1.5], "logσ_pop_prior": [-0.5, 0.7], "s_bc_prior": [0.05, 1.2],
              "logσ_bc_prior": [-0.3, 0.8], "logλ_prior": [2.5, 2.0], "logτ_prior": [-1.5, 0.6]}
    cases = [("multienv_replicate_fitness_normal", dict(n_neutral=6, n_bc=17, n_time=6, n_rep=2, envs=[1, 1, 2, 3, 2, 3])),
             ("genotype_fitness_normal", dict(n_neutral=5, n_bc=23, n_time=5, n_geno=4))]
    for model, spec in cases:
        da, _ = bb.synth.simulate(model, seed=11, **spec)
        prob = oracle_problem(da, model, priors)
        D = model_ref.n_latent(model, prob)
        pp = _model_port(da, model, priors)
        assert pp.D == D
        z = 0.3 * rng.standard_normal((2, D))
        z[:, D - np.asarray(da.bc_count).size:] += 4.0
        logp, grad = pp.logjoint_grad(z)
        for k in range(2):
            lp_ref, g_ref = model_ref.logjoint_and_grad(model, z[k], prob)
            assert abs(logp[k] - lp_ref) <= 1e-11 * abs(lp_ref), model
            assert rel_err(grad[k], g_ref) < 1e-10, model
