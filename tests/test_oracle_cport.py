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
