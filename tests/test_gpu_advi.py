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


def test_advi_uneven_replicates_pairings_and_multienv(bb):
    """Unequal T per replicate through the public call: the reference's neutral pairing as written (default) and the
    corrected one give different fits; the multienv x replicate model takes the per-replicate environment lists that
    data_to_arrays builds, and the frame labels the population rows with every replicate's own envs[2:end]."""
    df, cols = load_fixture("replicate_fitness_normal")
    df = uneven_replicates(df)
    kw = dict(model=bb.model.replicate_fitness_normal, advi=bb.ADVI(2, 300), opt=bb.DecayedADAGrad(), verbose=False,
              n_posterior_samples=100, seed=5, **cols)
    a = bb.advi(data=df, **kw)
    b = bb.advi(data=df, corrected_ragged=True, **kw)
    assert a.shape == b.shape and np.isfinite(a["mean"]).all() and np.isfinite(b["mean"]).all()
    pa, pb = a[a.vartype == "pop_mean_fitness"]["mean"].to_numpy(), b[b.vartype == "pop_mean_fitness"]["mean"].to_numpy()
    assert np.max(np.abs(pa - pb)) > 1e-6          # same data, same noise lattice: only the pairing differs
    with pytest.raises(bb.BarBayError, match="single shard"):
        bb.Engine(bb.utils.data_to_arrays(df, **cols), "replicate_fitness_normal", rank=0, world=2)
    dfe = df.assign(env=df.time.map({1: "A", 2: "A", 3: "B", 4: "C", 5: "B"}))
    out = bb.advi(data=dfe, model=bb.model.multienv_replicate_fitness_normal, advi=bb.ADVI(1, 50), verbose=False,
                  n_posterior_samples=100, env_col="env", **cols)
    pop = out[out.vartype == "pop_mean_fitness"]
    assert list(pop["env"]) == ["A", "B", "C", "B", "A", "B", "C"]       # R1: envs[2:5], R2 (4 time points): envs[2:4]
    assert set(out["vartype"]) >= {"bc_hyperfitness", "bc_fitness", "pop_std", "log_poisson"}
    assert np.isfinite(out["mean"]).all() and (out["std"] > 0).all()


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


@pytest.mark.parametrize("model", list(FIXTURES))
def test_averaged_posterior_matches_restated_cpu_advi(bb, model):
    """Converged posterior vs the restated CPU ADVI (oracle C port: same optimiser, its own xoshiro noise stream),
    all four model families.  Both chains run 4000 burn-in steps of DecayedADAGrad(0.03) with 4 samples per step and
    then average 300 iterates taken every 10 steps, which removes the optimiser's jitter: two independent CPU chains
    agree to 0.05 sd / 3 % this way, so the gate is |mean_gpu - mean_cpu| <= 0.25 sd and sd within 10 % for every
    latent above the log-Poisson block (population, hyper and barcode latents), and 0.05 absolute / 35 % for the
    log-Poisson latents (sd ~ 1e-3 for large counts: their iterates never settle below the step size)."""
    from oracle import cport
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    burn, blocks, every, K, eta = 4000, 300, 10, 4, 0.03
    pp = cport.ModelPort(model, da.bc_count, da.n_neutral, da.n_bc, envs=da.envs, genotypes=da.genotypes)
    rng = np.random.default_rng(1)
    D = pp.D
    theta = np.concatenate([rng.standard_normal(D), rng.standard_normal(D)])
    acc = np.full(2 * D, 1e-8)
    pp.advi_steps(theta, acc, burn, K, eta=eta, seed=1)
    m_cpu, s_cpu = np.zeros(D), np.zeros(D)
    for b in range(blocks):
        pp.advi_steps(theta, acc, every, K, eta=eta, seed=1, first_step=burn + b * every)
        m_cpu += theta[:D] / blocks
        s_cpu += np.log1p(np.exp(theta[D:])) / blocks

    eng = bb.Engine(da, model, n_samples=K, dtype="f64", seed=17)
    assert eng.D == D
    eng.init_params(3)
    eng.set_optimizer("decayed", eta=eta)
    eng.step(burn)
    m_gpu, s_gpu = np.zeros(D), np.zeros(D)
    for b in range(blocks):
        eng.step(every)
        m, s = eng.get_posterior()
        m_gpu += m / blocks
        s_gpu += s / blocks
    eng.close()
    n_lam = int(np.sum([np.asarray(c).size for c in da.bc_count])) if isinstance(da.bc_count, list) else np.asarray(da.bc_count).size
    head, lam = slice(0, D - n_lam), slice(D - n_lam, D)
    sd = np.maximum(s_gpu, s_cpu)
    z = np.abs(m_gpu - m_cpu)[head] / sd[head]
    assert z.max() <= 0.25, (model, z.max())
    assert np.abs(s_gpu / s_cpu - 1.0)[head].max() <= 0.10, (model, np.abs(s_gpu / s_cpu - 1.0)[head].max())
    assert np.abs(m_gpu - m_cpu)[lam].max() <= 0.05, (model, np.abs(m_gpu - m_cpu)[lam].max())
    assert np.abs(s_gpu / s_cpu - 1.0)[lam].max() <= 0.35, (model, np.abs(s_gpu / s_cpu - 1.0)[lam].max())


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


def test_elbo_trace_convergence_stop(bb):
    """bb_step_until (extension; the reference runs a fixed max_iters, src/vi.jl:98): stops once the windowed ELBO
    mean stops moving, reports the estimates it used, and degenerates to a plain max_iters run with rel_tol = 0."""
    df, _ = load_fixture("fitness_normal")
    da = bb.utils.data_to_arrays(df)
    eng = bb.Engine(da, "fitness_normal", n_samples=4, dtype="f64", seed=2)
    eng.init_params(1)
    eng.set_optimizer("decayed", eta=0.05)
    n_done, conv, est = eng.step_until(20000, every=50, window=4, rel_tol=2e-3)
    assert conv and n_done < 20000 and n_done % 50 == 0 and len(est) == n_done // 50
    assert eng.step_count == n_done
    assert est[-1] > est[0]                                    # the ELBO went up on the way
    n2, conv2, est2 = eng.step_until(130, every=50, window=2, rel_tol=0.0)
    assert n2 == 130 and not conv2 and len(est2) == 3 and eng.step_count == n_done + 130
    eng.close()
    out = bb.advi(data=df, model=bb.model.fitness_normal, advi=bb.ADVI(4, 20000), opt=bb.DecayedADAGrad(0.05),
                  verbose=False, seed=2, elbo_rel_tol=2e-3, elbo_every=50, elbo_window=4, n_posterior_samples=100)
    assert np.isfinite(out["mean"]).all()


@pytest.mark.parametrize("model", ["replicate_fitness_normal", "genotype_fitness_normal"])
def test_device_derived_fitness_rows_match_host_sampler(bb, model):
    """The derived bc_fitness rows (utils.jl:1284-1343: median and sd of 10^4 draws of θ + exp(logτ) θ̃) sampled on the
    device agree with the host sampler (numpy, its own stream) within Monte-Carlo error: the median's standard error
    is 1.25 sd / sqrt(n) = 0.0125 sd, the sd's 0.7 % (x ~1.5 for the exp(logτ) tail)."""
    df, cols = load_fixture(model)
    kw = dict(data=df, model=getattr(bb.model, model), advi=bb.ADVI(2, 300), opt=bb.DecayedADAGrad(0.05),
              verbose=False, seed=4, **cols)
    dev = bb.advi(**kw, device_derived_rows=True)
    host = bb.advi(**kw, device_derived_rows=False)
    assert list(dev.columns) == list(host.columns) and len(dev) == len(host)
    a = dev[dev.vartype == "bc_fitness"].reset_index(drop=True)
    b = host[host.vartype == "bc_fitness"].reset_index(drop=True)
    assert len(a) > 0 and (a["varname"] == b["varname"]).all() and (a["id"] == b["id"]).all()
    sd = b["std"].to_numpy()
    assert np.max(np.abs(a["mean"].to_numpy() - b["mean"].to_numpy()) / sd) < 0.08
    # exp(logτ) makes the draws heavy-tailed where the fit is still wide and the sample sd itself noisy: its relative
    # standard error is sqrt((kurtosis - 1) / 4n), kurtosis <= 3 exp(4 sd(logτ)^2) for a lognormal-scaled normal
    # (0.7 % for a normal, 6 % at sd(logτ) = 1).  Two independent samplers, four standard errors, never below 15 %:
    sd_tau = host.loc[host.vartype == "bc_deviations", "std"].to_numpy()
    assert sd_tau.shape == sd.shape
    se = np.sqrt((3.0 * np.exp(np.minimum(4.0 * sd_tau ** 2, 20.0)) - 1.0) / 4.0e4)
    dev_sd = np.abs(a["std"].to_numpy() / sd - 1.0)
    assert (dev_sd < np.maximum(0.15, 4.0 * np.sqrt(2.0) * se)).all(), (dev_sd, se)
    assert np.median(dev_sd) < 0.03
    # the fitted rows themselves are untouched by the choice
    fa, fb = dev[dev.vartype != "bc_fitness"], host[host.vartype != "bc_fitness"]
    assert np.array_equal(fa["mean"].to_numpy(), fb["mean"].to_numpy())
