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
