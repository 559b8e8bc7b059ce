"""GPU: stats.naive_prior on the device (bb_naive_prior, csrc/bb_naive.cuh; src/stats.jl:1175-1359) against the
loop-level oracle on the reference's fixtures and against the numpy statement of the same formulas at the
BASELINE sizes.  Tolerance: fp64 everywhere; the device `log` and glibc's differ by at most an ulp or two, the
block-tree sums by rounding -> rel 1e-12 on every entry (totals are exact integer sums)."""
import numpy as np
import pytest

from helpers import load_fixture, uneven_replicates
from oracle import stats_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["single", "replicates", "uneven", "multienv"])
def test_device_naive_prior_matches_oracle(bb, case):
    model = {"single": "fitness_normal", "replicates": "replicate_fitness_normal", "uneven": "replicate_fitness_normal",
             "multienv": "multienv_fitness_normal"}[case]
    df, _ = load_fixture(model)
    if case == "uneven":
        df = uneven_replicates(df)
    rep = "rep" if case in ("replicates", "uneven") else None
    ref = stats_ref.naive_prior_ref(df, rep_col=rep)
    host = bb.stats.naive_prior(df, rep_col=rep, mutate=False)
    got = bb.stats.naive_prior(df, rep_col=rep, mutate=False, device=0)
    assert set(got) == {"s_pop_prior", "logσ_pop_prior", "logλ_prior"}
    for k in got:
        assert got[k].shape == ref[k].shape and not np.isnan(got[k]).any()
        assert np.allclose(got[k], ref[k], rtol=1e-12, atol=1e-14), k
        assert np.allclose(got[k], host[k], rtol=1e-12, atol=1e-14), k


def _numpy_statement(blocks, N):
    """the reference's formulas on (T x B counts) blocks, vectorised (test infrastructure)"""
    s_pop, lsig, logl = [], [], []
    for R in blocks:
        tot = R.sum(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            f = R / tot[:, None]
            lr = np.log(f[1:, :N] / f[:-1, :N])
            logl.append(np.log(R.astype(np.float64)).T.reshape(-1))
        for row in lr:
            v = row[~np.isinf(row)]
            s_pop.append(-v.mean()); lsig.append(-v.std(ddof=1))
    return np.asarray(s_pop), np.asarray(lsig), np.concatenate(logl)


@pytest.mark.parametrize("cfg", [2, 3])
def test_device_naive_prior_at_baseline_size(bb, cfg):
    """cfg2 (10^6 barcodes x 5) and cfg3 (3 replicates x 2 10^5 x 5): pseudocount 0, so the zero counts of the
    simulation exercise the +-Inf exclusion (stats.jl:1260-1262) and log(0) = -Inf in the log-lambda prior."""
    _, da, _ = bb.synth.config(cfg)
    R = np.asarray(da.bc_count)
    blocks = [R] if R.ndim == 2 else [R[:, :, r] for r in range(R.shape[2])]
    s_ref, l_ref, ll_ref = _numpy_statement(blocks, da.n_neutral)
    got = bb.stats.naive_prior_packed(da, device=0)
    # a neutral barcode with zeros at consecutive times gives 0 / 0 = NaN, which the reference's isinf filter keeps
    # (cfg3 has one): the device result carries the same NaN
    assert np.allclose(got["s_pop_prior"], s_ref, rtol=1e-12, atol=1e-15, equal_nan=True)
    assert np.allclose(got["logσ_pop_prior"], l_ref, rtol=1e-12, atol=1e-15, equal_nan=True)
    assert np.isfinite(s_ref).sum() >= s_ref.size - 1
    fin = np.isfinite(ll_ref)
    assert (np.isneginf(got["logλ_prior"]) == ~fin).all()
    assert np.allclose(got["logλ_prior"][fin], ll_ref[fin], rtol=1e-14, atol=0)
    again = bb.stats.naive_prior_packed(da, device=0)                       # deterministic: fixed-order sums
    for k in got:
        assert np.array_equal(got[k], again[k], equal_nan=True)


def test_device_naive_prior_zero_neutral_counts(bb):
    """neutral ratios with a zero on either side are +-Inf and are left out of mean and sd"""
    rng = np.random.default_rng(5)
    T, N, M = 6, 300, 2000
    R = rng.poisson(3.0, size=(T, N + M)).astype(np.int64)
    R[:, :N][rng.random((T, N)) < 0.2] = 0
    R[R.sum(axis=1) == 0, -1] = 1
    # a column with zeros at consecutive times gives 0 / 0 = NaN, which the reference does NOT drop: avoid it here
    both = (R[1:, :N] == 0) & (R[:-1, :N] == 0)
    R[1:, :N][both] = 1
    da = type("DA", (), {"bc_count": R, "n_neutral": N, "n_bc": M})()
    s_ref, l_ref, _ = _numpy_statement([R], N)
    got = bb.stats.naive_prior_packed(da, device=0)
    assert np.isfinite(got["s_pop_prior"]).all() and np.isfinite(got["logσ_pop_prior"]).all()
    assert np.allclose(got["s_pop_prior"], s_ref, rtol=1e-12, atol=1e-15)
    assert np.allclose(got["logσ_pop_prior"], l_ref, rtol=1e-12, atol=1e-15)


def test_device_naive_prior_rejects_bad_arguments(bb):
    R = np.ones((1, 8), dtype=np.int64)
    da = type("DA", (), {"bc_count": R, "n_neutral": 3, "n_bc": 5})()
    with pytest.raises(bb.BarBayError, match="time points"):
        bb.stats.naive_prior_packed(da, device=0)
    da2 = type("DA", (), {"bc_count": np.ones((4, 8), dtype=np.int64), "n_neutral": 3, "n_bc": 5})()
    with pytest.raises(bb.BarBayError, match="no such device"):
        bb.stats.naive_prior_packed(da2, device=4096)
