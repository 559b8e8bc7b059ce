"""CPU: naive_prior / naive_fitness (host mirror of BarBay.stats) against the loop-level oracle and the
reference's own test assertions (test/stats_tests.jl:137-224)."""
import numpy as np
import pytest

from helpers import load_fixture, uneven_replicates
from oracle import stats_ref


@pytest.mark.parametrize("case", ["single", "replicates", "uneven", "multienv"])
def test_naive_prior_matches_oracle(bb, case):
    model = {"single": "fitness_normal", "replicates": "replicate_fitness_normal", "uneven": "replicate_fitness_normal",
             "multienv": "multienv_fitness_normal"}[case]
    df, _ = load_fixture(model)
    if case == "uneven":
        df = uneven_replicates(df)
    rep = "rep" if case in ("replicates", "uneven") else None
    ref = stats_ref.naive_prior_ref(df, rep_col=rep)
    before = df["count"].copy()
    got = bb.stats.naive_prior(df, rep_col=rep)
    assert (df["count"] == before + 1).all()                                # the reference mutates the caller's frame
    assert set(got) == {"s_pop_prior", "logσ_pop_prior", "logλ_prior"}
    for k in got:
        assert got[k].shape == ref[k].shape and not np.isnan(got[k]).any()
        assert np.allclose(got[k], ref[k], rtol=1e-12, atol=1e-14), k
    n_rows_unique = df.drop_duplicates(["barcode", "time"] + ([rep] if rep else [])).shape[0]
    assert got["logλ_prior"].size == n_rows_unique                          # stats_tests.jl:180-183
    assert (got["logσ_pop_prior"] < 0).all()                                 # -std: the reference's quirk 6
    again = bb.stats.naive_prior(df, rep_col=rep, mutate=False)
    assert (df["count"] == before + 1).all() and again["logλ_prior"].size == got["logλ_prior"].size


def test_naive_prior_lengths_match_reference_tests(bb):
    df, _ = load_fixture("replicate_fitness_normal")
    r = bb.stats.naive_prior(df.copy(), rep_col="rep")
    assert r["s_pop_prior"].size == (5 - 1) * 2 == r["logσ_pop_prior"].size     # stats_tests.jl:203-206
    u = bb.stats.naive_prior(uneven_replicates(df).copy(), rep_col="rep")
    assert u["s_pop_prior"].size == 4 + 3                                       # stats_tests.jl:218-224


def test_naive_fitness_matches_oracle(bb):
    df, _ = load_fixture("fitness_normal")
    got = bb.stats.naive_fitness(df)
    ref = stats_ref.naive_fitness_ref(df)
    assert list(got["barcode"]) == list(ref["barcode"]) and len(got) == 10
    assert np.allclose(got["fitness"], ref["fitness"], rtol=1e-12)
    truth = df.groupby("barcode", sort=False)["fitness"].first().loc[got["barcode"]].to_numpy()
    assert np.corrcoef(truth, got["fitness"])[0, 1] > 0.8                       # tracks the simulator's ground truth


def test_prior_matrices_feed_the_engine_layout(bb):
    df, _ = load_fixture("fitness_normal")
    pri = bb.stats.prior_matrices(bb.stats.naive_prior(df.copy()))
    da = bb.utils.data_to_arrays(df)
    lay = bb.model.var_groups(bb.model.fitness_normal, da.n_time, 1, da.n_neutral, da.n_bc)
    size = {g.name: g.length for g in lay.groups}
    assert pri["s_pop_prior"].shape == (size[bb.model.V_S_POP], 2)
    assert pri["logλ_prior"].shape == (size[bb.model.V_LOGLAM], 2)


# ------------------------------------------------------------------ posterior predictive checks (stats.jl:35-996)
def _draws(n=4000, seed=0):
    import pandas as pd
    r = np.random.default_rng(seed)
    return pd.DataFrame({"s̲ₜ[1]": 0.3 + 0.01 * r.standard_normal(n), "s̲ₜ[2]": 0.5 + 0.01 * r.standard_normal(n),
                         "s⁽ᵐ⁾": 1.0 + 0.01 * r.standard_normal(n), "σ⁽ᵐ⁾": np.full(n, np.log(0.2)),
                         "f̲⁽ᵐ⁾[1]": np.full(n, 1e-3)})


def test_logfreq_ratio_bc_ppc_shapes_and_moments(bb):
    df = _draws()
    flat = bb.stats.logfreq_ratio_bc_ppc(df, 3, rng=np.random.default_rng(1))
    raw = bb.stats.logfreq_ratio_bc_ppc(df, 3, flatten=False, rng=np.random.default_rng(1))
    assert flat.shape == (3 * len(df), 2) and raw.shape == (len(df), 2, 3)
    assert np.array_equal(flat[len(df):2 * len(df)], raw[:, :, 1])           # vcat of the n_ppc slices
    # N(s - sbar_t, exp(sigma)): mean 0.7 / 0.5, sd 0.2 (+ the 0.014 spread of the draws)
    assert np.allclose(flat.mean(axis=0), [0.7, 0.5], atol=0.01) and np.allclose(flat.std(axis=0), 0.2, atol=0.01)


def test_popmean_and_multienv_ppc(bb):
    import pandas as pd
    n = 3000
    df = pd.DataFrame({"sₜ[1]": np.full(n, 0.4), "sₜ[2]": np.full(n, 0.1), "σₜ[1]": np.full(n, np.log(0.1)),
                       "σₜ[2]": np.full(n, np.log(0.3))})
    pm = bb.stats.logfreq_ratio_popmean_ppc(df, 2, rng=np.random.default_rng(2))
    assert pm.shape == (2 * n, 2)
    assert np.allclose(pm.mean(axis=0), [-0.4, -0.1], atol=0.02) and np.allclose(pm.std(axis=0), [0.1, 0.3], atol=0.02)
    with pytest.raises(bb.BarBayError, match="does not match"):
        bb.stats.logfreq_ratio_popmean_ppc(df.drop(columns=["σₜ[2]"]), 2)
    me = pd.DataFrame({"s̲ₜ[1]": np.zeros(n), "s̲ₜ[2]": np.zeros(n), "s̲⁽ᵐ⁾[1]": np.full(n, 1.0), "s̲⁽ᵐ⁾[2]": np.full(n, -1.0),
                       "σ̲⁽ᵐ⁾[1]": np.full(n, np.log(0.05)), "σ̲⁽ᵐ⁾[2]": np.full(n, np.log(0.05))})
    out = bb.stats.logfreq_ratio_multienv_ppc(me, 1, ["a", "a", "b"], rng=np.random.default_rng(3))
    assert np.allclose(out.mean(axis=0), [1.0, -1.0], atol=0.01)             # ratio t -> t+1 uses the env of t+1
    with pytest.raises(bb.BarBayError, match="environments"):
        bb.stats.logfreq_ratio_multienv_ppc(me, 1, ["a", "b"])
    with pytest.raises(bb.BarBayError, match="# of mutant-related"):
        bb.stats.logfreq_ratio_multienv_ppc(me, 1, ["a", "b", "c"])


def test_freq_bc_ppc_and_quantile_range(bb):
    df = _draws(2000)
    df["σ⁽ᵐ⁾"] = 0.2                                                         # :lognormal takes sigma as is
    f = bb.stats.freq_bc_ppc(df, 2, flatten=False, rng=np.random.default_rng(4))
    assert f.shape == (len(df), 3, 2) and np.all(f[:, 0, :] == 1e-3)
    assert np.allclose(np.log(f[:, 1, :] / f[:, 0, :]).mean(), 0.7, atol=0.02)
    fn = bb.stats.freq_bc_ppc(df.assign(**{"σ⁽ᵐ⁾": np.log(0.2)}), 2, model="normal", flatten=False,
                              rng=np.random.default_rng(4))
    assert np.allclose(fn, f)                                                # exp(log 0.2) == 0.2: same draws
    with pytest.raises(bb.BarBayError, match="model must be"):
        bb.stats.freq_bc_ppc(df, 1, model="poisson")
    m = np.random.default_rng(5).standard_normal((500, 4))
    q = bb.stats.matrix_quantile_range([0.95, 0.5], m, dims=2)
    assert q.shape == (4, 2, 2)
    assert np.allclose(q[:, 0, 0], np.quantile(m, 0.025, axis=0)) and np.allclose(q[:, 1, 1], np.quantile(m, 0.75, axis=0))
    with pytest.raises(bb.BarBayError, match="between zero and one"):
        bb.stats.matrix_quantile_range([1.5], m)
