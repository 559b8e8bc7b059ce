"""CPU: data_to_arrays (O(rows) product packer) is bit-exact against the oracle's naive transliteration
of utils.jl on the reference's fixtures and on shuffled / relabelled / integer-id variants."""
import numpy as np
import pandas as pd
import pytest

from helpers import FIXTURES, load_fixture, uneven_replicates
from oracle.packer_ref import data_to_arrays_ref


def _same(a, b):
    if isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
        return
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == np.int64 and b.dtype == np.int64 and a.shape == b.shape
    assert np.array_equal(a, b)


def _compare(bb, df, cols):
    got = bb.utils.data_to_arrays(df, **cols)
    ref = data_to_arrays_ref(df, **cols)
    _same(got.bc_count, ref.bc_count)
    _same(got.bc_total, ref.bc_total)
    assert (got.n_neutral, got.n_bc, got.n_env, got.n_rep, got.n_geno) == \
           (ref.n_neutral, ref.n_bc, ref.n_env, ref.n_rep, ref.n_geno)
    assert list(got.bc_ids) == list(ref.bc_ids) and list(got.neutral_ids) == list(ref.neutral_ids)
    assert got.n_time == ref.n_time and got.envs == ref.envs and got.genotypes == ref.genotypes
    return got


@pytest.mark.parametrize("model", list(FIXTURES))
def test_fixtures_bit_exact(bb, model):
    df, cols = load_fixture(model)
    da = _compare(bb, df, cols)
    # neutrals first, totals are the row sums (utils.jl:428-432)
    R = np.asarray(da.bc_count)
    assert np.array_equal(np.asarray(da.bc_total), R.sum(axis=1))
    assert da.n_neutral == 5 and da.n_bc == 10


@pytest.mark.parametrize("model", list(FIXTURES))
@pytest.mark.parametrize("seed", [0, 1])
def test_shuffled_rows_bit_exact(bb, model, seed):
    df, cols = load_fixture(model)
    df = df.sample(frac=1.0, random_state=seed).reset_index(drop=True)
    _compare(bb, df, cols)


def test_uneven_replicates_bit_exact(bb):
    df, cols = load_fixture("replicate_fitness_normal")
    da = _compare(bb, uneven_replicates(df), cols)
    assert isinstance(da.bc_count, list) and [m.shape for m in da.bc_count] == [(5, 15), (4, 15)]


def test_integer_ids_follow_dataframes_group_order(bb):
    """Narrow-range Integer keys are grouped in value order by DataFrames.jl, strings by first appearance."""
    df, cols = load_fixture("fitness_normal")
    ids = {b: i for i, b in enumerate(sorted(df.barcode.unique(), reverse=True))}   # reversed integer labels
    df2 = df.assign(barcode=df.barcode.map(ids).astype(np.int64)).sample(frac=1.0, random_state=3)
    da = _compare(bb, df2.reset_index(drop=True), cols)
    assert da.bc_ids == sorted(da.bc_ids) and da.neutral_ids == sorted(da.neutral_ids)


def test_string_time_labels_sort_lexicographically(bb):
    df, cols = load_fixture("fitness_normal")
    df2 = df.assign(time=df.time.map(lambda t: f"t{t:02d}"))
    a = bb.utils.data_to_arrays(df2, **cols)
    b = bb.utils.data_to_arrays(df, **cols)
    assert np.array_equal(a.bc_count, b.bc_count)


def test_multienv_replicate_packing(bb):
    df, _ = load_fixture("replicate_fitness_normal")
    df = df.assign(env=df.time.map({1: "A", 2: "A", 3: "B", 4: "C", 5: "B"}))
    da = _compare(bb, df, {"rep_col": "rep", "env_col": "env"})
    assert da.envs == ["A", "A", "B", "C", "B"] and da.n_env == 3 and np.asarray(da.bc_count).shape == (5, 15, 2)


def test_packer_errors_match_reference(bb):
    """utils.jl:1007-1018 and :113-121; test/utils_tests.jl:128-136, 277-287, 386-392."""
    df, _ = load_fixture("fitness_normal")
    with pytest.raises(bb.BarBayError, match="does not exist"):
        bb.utils.data_to_arrays(df, id_col="nope")
    with pytest.raises(bb.BarBayError, match="does not exist"):
        bb.utils.data_to_arrays(df, rep_col="rep")
    with pytest.raises(bb.BarBayError, match="must be of type Bool"):
        bb.utils.data_to_arrays(df.assign(neutral=df.neutral.astype(int)))
    with pytest.raises(bb.BarBayError, match="Not all neutral barcodes"):
        bb.utils.data_to_arrays(df.iloc[1:])                       # first neutral loses a time point
    mut_row = df.index[~df.neutral][0]
    with pytest.raises(bb.BarBayError, match="Not all mutant barcodes"):
        bb.utils.data_to_arrays(df.drop(index=mut_row))


def test_large_frame_is_fast_and_consistent(bb):
    """O(rows) packing: 2e4 barcodes x 5 time points x 3 replicates in well under the reference's
    O(ids * reps * rows) cost, and equal to the arrays the frame was built from."""
    import time
    model, da, _ = bb.synth.config(3, scale=0.1)
    R = np.asarray(da.bc_count)
    T, B, nrep = R.shape
    t_idx, b_idx, r_idx = np.meshgrid(np.arange(T), np.arange(B), np.arange(nrep), indexing="ij")
    ids = np.asarray(list(da.neutral_ids) + list(da.bc_ids), dtype=object)
    df = pd.DataFrame({"time": t_idx.ravel() + 1, "barcode": ids[b_idx.ravel()], "count": R.ravel(),
                       "neutral": b_idx.ravel() < da.n_neutral,
                       "rep": np.asarray([f"R{r + 1}" for r in range(nrep)], dtype=object)[r_idx.ravel()]})
    df = df.sample(frac=1.0, random_state=0).reset_index(drop=True)
    t0 = time.time()
    out = bb.utils.data_to_arrays(df, rep_col="rep")
    assert time.time() - t0 < 30
    # neutrals: first appearance order (shuffled) ; mutants: sorted -> compare as id-keyed dictionaries
    pos = {b: i for i, b in enumerate(ids)}
    order = [pos[b] for b in list(out.neutral_ids) + list(out.bc_ids)]
    reps_first = pd.unique(df.loc[df.neutral, "rep"])              # neutral replicate order (utils.jl:200)
    rn = [int(r[1:]) - 1 for r in reps_first]
    assert np.array_equal(np.asarray(out.bc_count)[:, :da.n_neutral, :], R[:, order[:da.n_neutral], :][:, :, rn])
    assert np.array_equal(np.asarray(out.bc_count)[:, da.n_neutral:, :], R[:, order[da.n_neutral:], :])
