"""GPU end to end: advi() -- the mirror of BarBay.vi.advi -- on the reference's fixtures: the output
contracts of test/vi_tests.jl, convergence against the restated CPU ADVI, and recovery of the
simulator ground truth shipped in the CSVs."""
import os

import numpy as np
import pandas as pd
import pytest

from helpers import FIXTURES, load_fixture, oracle_problem, uneven_replicates

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", list(FIXTURES))
def test_advi_one_step_output_contract(bb, model):
    """test/vi_tests.jl:23-174: ADVI(1, 1) returns a DataFrame with the documented columns / vartypes."""
    df, cols = load_fixture(model)
    out = bb.advi(data=df, model=getattr(bb.model, model), advi=bb.ADVI(1, 1), verbose=False,
                  n_posterior_samples=200, **cols)
    assert isinstance(out, pd.DataFrame)
    for c in ("mean", "std", "vartype", "varname", "id"):
        assert c in out.columns
    need = {"pop_mean_fitness", "pop_std", "bc_fitness", "bc_std", "log_poisson"}
    if getattr(bb.model, model).hier:
        need |= {"bc_hyperfitness", "bc_noncenter", "bc_deviations"}
    assert need <= set(out.vartype)
    assert np.isfinite(out["mean"]).all() and (out["std"] > 0).all()
    for c in cols.values():
        if c != "genotype":
            assert c in out.columns


def test_advi_uneven_replicates(bb):
    df, cols = load_fixture("replicate_fitness_normal")
    out = bb.advi(data=uneven_replicates(df), model=bb.model.replicate_fitness_normal, advi=bb.ADVI(1, 1),
                  verbose=False, n_posterior_samples=100, **cols)
    assert isinstance(out, pd.DataFrame)


def test_advi_csv_output(bb, tmp_path):
    """test/vi_tests.jl:213-236: with outputname the call returns nothing and writes <name>.csv."""
    df, _ = load_fixture("fitness_normal")
    name = str(tmp_path / "out")
    res = bb.advi(data=df, model=bb.model.fitness_normal, outputname=name, advi=bb.ADVI(1, 1), verbose=False)
    assert res is None and os.path.isfile(name + ".csv")
    back = pd.read_csv(name + ".csv")
    assert list(back.columns) == ["mean", "std", "varname", "vartype", "id"]
    with pytest.raises(bb.BarBayError, match="already processed"):
        bb.advi(data=df, model=bb.model.fitness_normal, outputname=name, advi=bb.ADVI(1, 1), verbose=False)


def test_posterior_matches_restated_cpu_advi(bb):
    """Converged posterior mean / sd vs the oracle's ADVI (different noise streams): agreement within
    Monte-Carlo error.  Both run 1500 steps of DecayedADAGrad with 4 samples per step."""
    from oracle import advi_ref, philox_ref
    model = "fitness_normal"
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df)
    priors = {"logλ_prior": np.column_stack([np.log(np.asarray(da.bc_count).T.reshape(-1) + 1.0),
                                             np.full(np.asarray(da.bc_count).size, 3.0)])}
    eng = bb.Engine(da, model, dict(priors), n_samples=4, dtype="f64", seed=17)
    eng.init_params(3)
    mu0, om0 = eng.get_params()
    eng.set_optimizer("decayed", eta=0.05)
    eng.step(1500)
    m_gpu, s_gpu = eng.get_posterior()
    eng.close()
    prob = oracle_problem(da, model, priors=priors)
    tr = advi_ref.advi_run(model, prob, 1500, 4, advi_ref.DecayedADAGrad(0.05), mu0, om0, seed=99)
    m_cpu, s_cpu = tr.mu, tr.sigma
    lay = bb.model.var_groups(bb.model.fitness_normal, da.n_time, 1, da.n_neutral, da.n_bc)
    g = {x.name: x for x in lay.groups}
    sl = slice(g[bb.model.V_S_BC].start, g[bb.model.V_S_BC].start + da.n_bc)
    # stated bound: |Δmean| <= 4 posterior sd (both chains still jitter with the AdaGrad step) and sd within 50 %
    assert np.all(np.abs(m_gpu[sl] - m_cpu[sl]) <= 4 * np.maximum(s_gpu[sl], s_cpu[sl]) + 0.02)
    assert np.all(np.abs(np.log(s_gpu[sl] / s_cpu[sl])) < 0.7)
    pop = slice(g[bb.model.V_S_POP].start, g[bb.model.V_S_POP].start + g[bb.model.V_S_POP].length)
    assert np.all(np.abs(m_gpu[pop] - m_cpu[pop]) <= 4 * np.maximum(s_gpu[pop], s_cpu[pop]) + 0.02)


def test_recovers_simulator_ground_truth(bb):
    """data001 ships the simulator's `fitness` per barcode; after convergence the
    posterior means of s^(m) track it (reference docs workflow: informative logλ prior)."""
    df, _ = load_fixture("fitness_normal")
    da = bb.utils.data_to_arrays(df)
    truth = df.groupby("barcode", sort=False)["fitness"].first()
    R = np.asarray(da.bc_count)
    priors = {"logλ_prior": np.column_stack([np.log(R.T.reshape(-1) + 1.0), np.full(R.size, 3.0)])}
    out = bb.advi(data=df, model=bb.model.fitness_normal, model_kwargs=priors,
                  advi=bb.ADVI(4, 4000), opt=bb.DecayedADAGrad(0.05), verbose=False, seed=5)
    fit = out[out.vartype == "bc_fitness"].set_index("id")["mean"]
    t = truth.loc[fit.index].to_numpy()
    f = fit.to_numpy()
    sd = out[out.vartype == "bc_fitness"].set_index("id")["std"].loc[fit.index].to_numpy()
    # 5 time points and 5 neutrals: posterior sds are 0.1-0.4, so recovery is judged against them
    assert np.corrcoef(t, f)[0, 1] > 0.75, np.corrcoef(t, f)[0, 1]
    assert np.all(np.abs(f - t) <= 4.0 * sd), np.max(np.abs(f - t) / sd)
    assert abs(np.mean(f - t)) < 0.15, np.mean(f - t)
    pop = out[out.vartype == "pop_mean_fitness"]["mean"].to_numpy()
    F = (R + 1.0) / (R + 1.0).sum(axis=1, keepdims=True)
    naive = -np.log(F[1:, :da.n_neutral] / F[:-1, :da.n_neutral]).mean(axis=1)   # stats.naive_fitness idea
    assert np.max(np.abs(pop - naive)) < 0.25, np.max(np.abs(pop - naive))


def test_documented_workflow_with_naive_priors(bb):
    """docs/src/examples.md:121-160: naive_prior -> matrix priors -> advi(TruncatedADAGrad).  The
    population mean fitness posterior must stay near its informative prior and the output keep the
    reference's columns."""
    df, _ = load_fixture("fitness_normal")
    pri = bb.stats.prior_matrices(bb.stats.naive_prior(df.copy()))
    out = bb.advi(data=df, model=bb.model.fitness_normal, model_kwargs=pri, advi=bb.ADVI(1, 3000),
                  opt=bb.TruncatedADAGrad(), verbose=False, seed=11, dtype="f32")
    pop = out[out.vartype == "pop_mean_fitness"]
    assert np.all(np.abs(pop["mean"].to_numpy() - pri["s_pop_prior"][:, 0]) < 0.2)
    assert list(out.columns) == ["mean", "std", "varname", "vartype", "id"]
    assert np.isfinite(out["mean"]).all() and (out["std"] > 0).all()
