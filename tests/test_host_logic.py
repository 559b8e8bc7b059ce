"""CPU: host-side mirror of the reference interface -- model layouts, variable names, advi_to_df,
the advi() argument checks -- and that the C-ABI library loads and exports every declared symbol
(no compute calls: there is no GPU here)."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from helpers import FIXTURES, ROOT, load_fixture, uneven_replicates


def _fake_q(bb, layout, rng):
    m = rng.standard_normal(layout.n_latent)
    s = np.abs(rng.standard_normal(layout.n_latent)) + 0.1
    return bb.utils.MeanFieldPosterior.build(m, s, layout.ranges_out)


def _layout(bb, model, da):
    return bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc, da.n_env,
                               da.n_geno)


def test_variable_names_match_reference_spelling(bb):
    M = bb.model
    assert [ord(c) for c in M.V_S_POP] == [0x73, 0x332, 0x209c]
    assert [ord(c) for c in M.V_LOGLAM] == [0x6c, 0x6f, 0x67, 0x39b, 0x332, 0x332]
    assert [ord(c) for c in M.V_THETA_TILDE] == [0x3b8, 0x332, 0x303, 0x207d, 0x1d50, 0x207e]
    ref = "/root/reference/src/utils.jl"
    if os.path.exists(ref):                      # only in the build container
        src = open(ref, encoding="utf8").read()
        block = re.search(r"const varname_to_vartype = Dict\((.*?)\n\)", src, re.S).group(1)
        pairs = dict(re.findall(r'"([^"]*)"\s*=>\s*"([^"]*)"', block))
        assert pairs == M.VARNAME_TO_VARTYPE


@pytest.mark.parametrize("model", list(FIXTURES))
def test_advi_to_df_columns_and_vartypes(bb, model):
    """test/vi_tests.jl:33-47, 78-89, 133-134, 169-174: column names and vartype strings."""
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = _layout(bb, model, da)
    q = _fake_q(bb, lay, np.random.default_rng(0))
    out = bb.utils.advi_to_df(df, q, lay.var_names, n_samples=500, seed=1, **cols)
    for c in ("mean", "std", "varname", "vartype", "id"):
        assert c in out.columns
    base = out.iloc[:lay.n_latent]
    assert np.array_equal(base["mean"].to_numpy(), q.dist.m) and np.array_equal(base["std"].to_numpy(), q.dist.σ)
    assert list(base["varname"]) == lay.var_names
    assert "tmp" not in set(out["vartype"])
    need = {"pop_mean_fitness", "pop_std", "bc_fitness", "bc_std", "log_poisson"}
    if bb.model.resolve(model).hier:
        need |= {"bc_hyperfitness", "bc_noncenter", "bc_deviations"}
        extra = out.iloc[lay.n_latent:]
        assert len(extra) == da.n_bc * da.n_rep and set(extra["vartype"]) == {"bc_fitness"}
        assert all(v.startswith("s") for v in extra["varname"])        # "logτ" -> "s" (utils.jl:1322)
        assert np.isfinite(extra["mean"]).all() and (extra["std"] > 0).all()
    assert need <= set(out["vartype"])
    if "rep_col" in cols:
        assert set(out["rep"]) == {"R1", "R2", "N/A"}
    if "env_col" in cols:
        pop = out[out.vartype == "pop_mean_fitness"]
        assert list(pop["env"]) == da.envs[1:]
    lam = out[out.vartype == "log_poisson"]
    assert list(lam["id"].iloc[:np.asarray(da.bc_count).shape[0]]) == [da.neutral_ids[0]] * np.asarray(da.bc_count).shape[0]
    assert set(out.loc[out.vartype == "pop_std", "id"]) == {"N/A"}


def test_advi_to_df_uneven_replicates(bb):
    df, cols = load_fixture("replicate_fitness_normal")
    df = uneven_replicates(df)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = _layout(bb, "replicate_fitness_normal", da)
    assert lay.n_latent == 2 * 7 + 10 + 3 * 20 + 15 * 9
    out = bb.utils.advi_to_df(df, _fake_q(bb, lay, np.random.default_rng(1)), lay.var_names, n_samples=200, **cols)
    pop = out[out.vartype == "pop_mean_fitness"]
    assert list(pop["rep"]) == ["R1"] * 4 + ["R2"] * 3


def test_model_dispatch_and_kwargs(bb):
    M = bb.model
    assert str(M.replicate_fitness_normal) == "replicate_fitness_normal" and "multienv" in str(M.multienv_fitness_normal)
    kw = M.normalise_kwargs(M.fitness_normal, {"loglam_prior": [1.0, 2.0]})
    assert kw["logλ_prior"] == [1.0, 2.0] and kw["s_pop_prior"] == [0.0, 2.0]
    with pytest.raises(TypeError):
        M.normalise_kwargs(M.fitness_normal, {"logτ_prior": [0.0, 1.0]})
    uniq, idx = M.indexin_unique(["b", "a", "b", "c"])
    assert uniq == ["b", "a", "c"] and idx.tolist() == [1, 2, 1, 3]


def test_advi_argument_errors_match_reference(bb, tmp_path):
    """src/vi.jl:103-118; test/vi_tests.jl:196-207 -- raised before any GPU work."""
    df, _ = load_fixture("fitness_normal")
    with pytest.raises(bb.BarBayError, match="require argument `:rep_col`"):
        bb.advi(data=df, model=bb.model.replicate_fitness_normal, verbose=False)
    with pytest.raises(bb.BarBayError, match="require argument `:env_col`"):
        bb.advi(data=df, model=bb.model.multienv_fitness_normal, verbose=False)
    out = tmp_path / "done"
    (tmp_path / "done.csv").write_text("x")
    with pytest.raises(bb.BarBayError, match="was already processed"):
        bb.advi(data=df, model=bb.model.fitness_normal, outputname=str(out), verbose=False)


def test_no_cpu_fallback(bb):
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    df, cols = load_fixture("fitness_normal")
    with pytest.raises(bb.BarBayError, match="no CPU fallback"):
        bb.advi(data=df, model=bb.model.fitness_normal, advi=bb.ADVI(1, 1), verbose=False)


def test_c_abi_exports_every_declared_symbol(bb):
    """Every function declared in include/barbay_b200.h is exported by libbarbay_b200.so and typed."""
    hdr = open(os.path.join(ROOT, "include", "barbay_b200.h"), encoding="utf8").read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 26
    lib = bb.load_library()
    bound = {name for name, _, _ in bb._lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)
    assert lib.bb_abi_version() == 2
    assert lib.bb_n_latent(None) == -1 and lib.bb_last_error(None) is not None


def test_device_naive_prior_fails_loudly_without_a_gpu(bb):
    """stats.naive_prior(device=...) is the CUDA path (bb_naive_prior): no silent numpy fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine WITHOUT a GPU")
    df, _ = load_fixture("fitness_normal")
    with pytest.raises(bb.BarBayError, match="no CPU fallback"):
        bb.stats.naive_prior(df, mutate=False, device=0)
    assert bb.stats.naive_prior(df, mutate=False)["s_pop_prior"].size == 4     # the host mirror is a separate, explicit call


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "barbay.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                src = open(os.path.join(dirpath, f), encoding="utf8").read()
                assert "from oracle" not in src and "import oracle" not in src, f
                assert "advi_port" not in src and "oracle/c" not in src, f      # nor link / execute it


def test_synthetic_configs_shapes(bb):
    for cfg, (model, dims) in {2: ("fitness_normal", 2), 3: ("replicate_fitness_normal", 3),
                               4: ("multienv_fitness_normal", 2), 5: ("genotype_fitness_normal", 2)}.items():
        m, da, truth = bb.synth.config(cfg, scale=0.001)
        assert m == model and np.asarray(da.bc_count).ndim == dims
        R = np.asarray(da.bc_count)
        assert R.dtype == np.int64 and (R >= 0).all() and np.array_equal(np.asarray(da.bc_total), R.sum(axis=1))
    _, da, _ = bb.synth.config(4, scale=0.001)
    assert da.envs == [1, 1, 2, 3, 4, 2, 3, 4] and da.n_env == 4
    df = bb.synth.to_tidy(bb.synth.config(2, scale=0.0005)[1])
    back = bb.utils.data_to_arrays(df)
    assert np.array_equal(back.bc_count, bb.synth.config(2, scale=0.0005)[1].bc_count)


@pytest.mark.parametrize("model", list(FIXTURES))
def test_lazy_variable_names_equal_the_list_the_reference_builds(bb, model):
    """`layout.var_names` is a lazy sequence (7 * 10^6 Python strings at BASELINE size cost more than the fit): it must
    behave like the list of src/vi.jl:184-198, and `advi_to_df` must produce the same frame -- values and dtypes --
    from it (vectorised Arrow path) as from the plain list (generic path); integer ids keep their type."""
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = _layout(bb, model, da)
    names = lay.var_names
    plain = [f"{g.name}[{i}]" for g in lay.groups for i in range(1, g.length + 1)]
    assert len(names) == len(plain) == lay.n_latent and names == plain and list(names) == plain
    assert names[0] == plain[0] and names[-1] == plain[-1] and names[3:9] == plain[3:9]
    assert names.groups == [g.name for g in lay.groups]
    assert list(names.to_pandas()) == plain
    with pytest.raises(IndexError):
        names[len(plain)]
    q = _fake_q(bb, lay, np.random.default_rng(2))
    fast = bb.utils.advi_to_df(df, q, names, n_samples=100, seed=3, **cols)
    slow = bb.utils.advi_to_df(df, q, plain, n_samples=100, seed=3, **cols)
    assert fast.equals(slow) and list(fast.dtypes) == list(slow.dtypes)
    # integer barcode ids: the id column keeps the objects as they are
    df_int = df.copy()
    codes = {b: i for i, b in enumerate(sorted(df_int["barcode"].unique()))}
    df_int["barcode"] = df_int["barcode"].map(codes)
    da_i = bb.utils.data_to_arrays(df_int, **cols)
    out_i = bb.utils.advi_to_df(df_int, q, names, n_samples=50, seed=3, output=da_i, **cols)
    ids = [x for x in out_i["id"] if not (isinstance(x, str) and x == "N/A")]
    assert any(isinstance(x, (int, np.integer)) for x in ids)                       # barcode rows: still integers
    assert not any(isinstance(x, str) and x.isdigit() for x in ids)                 # ... none turned into a string
