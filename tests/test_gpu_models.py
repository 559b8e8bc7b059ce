"""GPU parity beyond the four fixtures: matrix priors, the multienv x replicate model (M5), the
runtime-T / runtime-E fallback kernels, a 10^4-barcode synthetic, and size-independent properties
at the full BASELINE size."""
import numpy as np
import pytest

from helpers import load_fixture, oracle_problem, plausible_latents, plausible_theta, rel_err

pytestmark = pytest.mark.gpu


def _check_logjoint(bb, da, model, kwargs, priors, dtype="f64", K=2, seed=0, tol=1e-9):
    from oracle import model_ref
    eng = bb.Engine(da, model, kwargs, n_samples=K, dtype=dtype, seed=1)
    rng = np.random.default_rng(seed)
    z = plausible_latents(eng.layout, da, rng, K)
    logp, grad = eng.logjoint_grad(z)
    prob = oracle_problem(da, model, priors=priors)
    for k in range(K):
        lp_ref, g_ref = model_ref.logjoint_and_grad(model, z[k], prob)
        assert abs(logp[k] - lp_ref) <= tol * abs(lp_ref), (logp[k], lp_ref)
        assert rel_err(grad[k], g_ref) <= tol, rel_err(grad[k], g_ref)
    eng.close()


@pytest.mark.parametrize("model", ["fitness_normal", "replicate_fitness_normal", "multienv_fitness_normal",
                                   "genotype_fitness_normal"])
def test_matrix_priors(bb, model):
    """Per-element n x 2 priors (model_fitness_normal.jl:137-203), the documented workflow with
    stats.naive_prior (docs/src/examples.md:121-140)."""
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc, da.n_env, da.n_geno)
    rng = np.random.default_rng(4)
    size = {g.name: g.length for g in lay.groups}
    M = bb.model

    def mat(n, mean, sd):
        return np.column_stack([mean + 0.3 * rng.standard_normal(n), sd * (0.5 + rng.random(n))])

    R = np.asarray(da.bc_count)
    flat = R.T.reshape(-1) if R.ndim == 2 else R.transpose(2, 1, 0).reshape(-1)
    priors = {"s_pop_prior": mat(size[M.V_S_POP], 0.0, 0.5), "logσ_pop_prior": mat(size[M.V_LOGSIG_POP], -1.0, 0.5),
              "logλ_prior": np.column_stack([np.log(flat + 1.0), np.full(flat.size, 3.0)]),
              "logσ_bc_prior": mat(size[M.V_LOGSIG_BC], -1.0, 0.7)}
    bc_name = M.V_THETA if bb.model.resolve(model).hier else M.V_S_BC
    priors["s_bc_prior"] = mat(size[bc_name], 0.0, 1.0)
    _check_logjoint(bb, da, model, dict(priors), priors)


def test_multienv_replicate_model(bb):
    """M5: src/model_multienv_fitness_normal_hierarchical_replicates.jl:158-363."""
    df, _ = load_fixture("replicate_fitness_normal")
    df = df.assign(env=df.time.map({1: "A", 2: "A", 3: "B", 4: "C", 5: "B"}))
    da = bb.utils.data_to_arrays(df, rep_col="rep", env_col="env")
    _check_logjoint(bb, da, "multienv_replicate_fitness_normal", {"envs": da.envs}, None)
    _check_logjoint(bb, da, "multienv_replicate_fitness_normal", {"envs": da.envs}, None, dtype="f32", tol=2e-3)


def test_ragged_multienv_replicate_model(bb):
    """M5, Vector{Matrix{Int64}} method (…hierarchical_replicates.jl:449-687): unequal T per replicate, one
    environment list per replicate (two replicates share T = 6 but not the list -> three launch groups)."""
    from helpers import uneven_replicates
    df, _ = load_fixture("replicate_fitness_normal")
    df = uneven_replicates(df.assign(env=df.time.map({1: "A", 2: "A", 3: "B", 4: "C", 5: "B"})))
    da = bb.utils.data_to_arrays(df, rep_col="rep", env_col="env")
    assert isinstance(da.bc_count, list) and isinstance(da.envs[0], list)
    model = "multienv_replicate_fitness_normal"
    _check_logjoint(bb, da, model, {"envs": da.envs}, None)
    _check_logjoint(bb, da, model, {"envs": da.envs}, None, dtype="f32", tol=2e-3)
    da, _ = bb.synth.simulate(model, 9, 130, [6, 4, 6], envs=[[1, 2, 3, 1, 2, 3], [1, 1, 2, 3], [1, 3, 2, 1, 3, 2]], seed=5)
    _check_logjoint(bb, da, model, {"envs": da.envs}, None)
    with pytest.raises(bb.BarBayError, match="one environment list per replicate"):
        bb.Engine(da, model, {"envs": [1, 2, 3, 1, 2, 3]})


def test_ragged_multienv_replicate_advi_trajectory(bb):
    """A few optimiser steps of the ragged M5 model with supplied noise == the restated AdvancedVI loop (fp64)."""
    from oracle import advi_ref
    model, K, n_steps = "multienv_replicate_fitness_normal", 2, 4
    da, _ = bb.synth.simulate(model, 6, 40, [5, 4], envs=[["a", "b", "b", "c", "a"], ["a", "c", "b", "b"]], seed=11)
    eng = bb.Engine(da, model, {"envs": da.envs}, n_samples=K, dtype="f64", seed=2)
    rng = np.random.default_rng(5)
    mu, omega = plausible_theta(eng.layout, da, rng)
    noise = rng.standard_normal((n_steps, K, eng.D))
    eng.set_params(mu, omega)
    eng.set_optimizer("truncated", eta=0.1, tau=1.0, n=3)
    for i in range(n_steps):
        eng.step_with_noise(noise[i])
    mu_g, om_g = eng.get_params()
    tr = advi_ref.advi_run(model, oracle_problem(da, model), n_steps, K, advi_ref.TruncatedADAGrad(0.1, 1.0, 3), mu, omega,
                           eps_fn=lambda s: noise[s])
    assert rel_err(mu_g, tr.mu) < 1e-8 and rel_err(om_g, tr.omega) < 1e-8
    # in-kernel Philox steps run and improve the ELBO
    eng.init_params(1); eng.set_optimizer("decayed")
    trace = eng.step(300, elbo_trace=True)
    assert np.all(np.isfinite(trace)) and np.mean(trace[-20:]) > np.mean(trace[:20])
    eng.close()


@pytest.mark.parametrize("n_time", [3, 6, 11])
def test_runtime_time_points_fallback(bb, n_time):
    """T without a compiled specialisation runs on the runtime-size kernels."""
    da, _ = bb.synth.simulate("fitness_normal", 6, 40, n_time, seed=3)
    _check_logjoint(bb, da, "fitness_normal", {}, None)


def test_many_environments_fallback(bb):
    da, _ = bb.synth.simulate("multienv_fitness_normal", 5, 30, 9, envs=[1, 2, 3, 4, 5, 6, 1, 2, 3], seed=5)
    _check_logjoint(bb, da, "multienv_fitness_normal", {"envs": da.envs}, None)


def test_synthetic_ten_thousand_barcodes(bb):
    """SURVEY §8d gate (2): log-joint / gradient parity on a 10^4-barcode synthetic, fp64 and fp32."""
    da, _ = bb.synth.simulate("fitness_normal", 100, 9_900, 5, seed=9)
    _check_logjoint(bb, da, "fitness_normal", {}, None, K=1)
    _check_logjoint(bb, da, "fitness_normal", {}, None, K=1, dtype="f32", tol=2e-3)


def test_genotype_groups_of_uneven_size(bb):
    da, _ = bb.synth.simulate("genotype_fitness_normal", 8, 200, 5, n_geno=7, seed=2)
    _check_logjoint(bb, da, "genotype_fitness_normal", {"genotypes": da.genotypes}, None)


def test_elbo_gradient_many_samples_chunked(bb):
    """K large enough that pass 1 sweeps the samples in chunks (shared-memory budget)."""
    from oracle import advi_ref
    model, K = "fitness_normal", 24
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    eng = bb.Engine(da, model, n_samples=K, dtype="f64", seed=2)
    rng = np.random.default_rng(1)
    mu, om = plausible_theta(eng.layout, da, rng)
    eps = rng.standard_normal((K, eng.D))
    eng.set_params(mu, om)
    elbo, g_mu, g_om = eng.elbo_grad(eps)
    e_ref, gm, go, _ = advi_ref.elbo_value_and_grad(model, oracle_problem(da, model), mu, om, eps)
    assert abs(elbo - e_ref) <= 1e-9 * abs(e_ref) and rel_err(g_mu, gm) < 1e-9 and rel_err(g_om, go) < 1e-9
    eng.close()


def test_full_size_properties(bb):
    """BASELINE configs[1] at full size (10^6 barcodes): properties that need no oracle run --
    determinism (same seed => bitwise same theta), fp32 vs fp64 agreement on the same noise lattice,
    ELBO increases, state round trip."""
    model, da, _ = bb.synth.config(2)
    out = {}
    for dtype in ("f32", "f64"):
        eng = bb.Engine(da, model, n_samples=2, dtype=dtype, seed=3)
        eng.init_params(1)
        eng.set_optimizer("decayed")
        e_before = eng.elbo_grad(step=1000)[0]            # fixed lattice step: same noise before / after
        tr = eng.step(5, elbo_trace=True)
        e_after = eng.elbo_grad(step=1000)[0]
        assert np.isfinite(tr).all() and e_after > e_before
        out[dtype] = (tr, *eng.get_posterior())
        if dtype == "f32":
            eng2 = bb.Engine(da, model, n_samples=2, dtype=dtype, seed=3)
            eng2.init_params(1)
            eng2.set_optimizer("decayed")
            eng2.step(5)
            m2, s2 = eng2.get_posterior()
            eng3 = bb.Engine(da, model, n_samples=2, dtype=dtype, seed=3)
            eng3.init_params(1)
            eng3.set_optimizer("decayed")
            eng3.step(5)
            m3b, s3b = eng3.get_posterior()
            eng3.close()
            assert np.array_equal(m2, m3b) and np.array_equal(s2, s3b)        # same calls => bitwise same theta
            # the traced (two-kernel, ELBO) path sums the partials in another order: equal up to fp32 rounding
            assert np.mean(np.abs(m2 - out[dtype][1]) > 1e-3) < 1e-3
            state = eng2.get_state()
            eng2.step(2)
            eng2.set_state(state)
            m3, _ = eng2.get_posterior()
            assert np.array_equal(m3, m2) and eng2.step_count == 5
            eng2.close()
        eng.close()
    tr32, m32, s32 = out["f32"]
    tr64, m64, s64 = out["f64"]
    assert np.all(np.abs(tr32 - tr64) <= 1e-4 * np.abs(tr64))
    # AdaGrad's first steps are sign-like (delta ~ eta * sign(g)): a latent whose gradient is within fp32
    # rounding of zero can step the other way, so agreement is required of all but a sliver of latents
    dm, ds = np.abs(m32 - m64), np.abs(s32 - s64) / s64
    assert np.median(dm) < 1e-5 and np.mean(dm > 1e-3) < 1e-3
    assert np.median(ds) < 1e-5 and np.mean(ds > 1e-3) < 1e-3


@pytest.mark.parametrize("n_neutral,n_bc,n_time", [(0, 6, 4), (5, 0, 4), (1, 1, 2), (3, 2, 2), (33, 95, 5)])
def test_edge_shapes(bb, n_neutral, n_bc, n_time):
    """Empty populations, the minimal two time points, sizes around the 32-column padding."""
    da, _ = bb.synth.simulate("fitness_normal", max(n_neutral, 1), max(n_bc, 1), n_time, seed=1)
    R = np.asarray(da.bc_count)
    # carve the requested (possibly empty) populations out of the simulated block
    cols = list(range(n_neutral)) + list(range(da.n_neutral, da.n_neutral + n_bc))
    da.bc_count = np.ascontiguousarray(R[:, cols])
    da.bc_total = da.bc_count.sum(axis=1)
    da.n_neutral, da.n_bc = n_neutral, n_bc
    da.neutral_ids, da.bc_ids = da.neutral_ids[:n_neutral], da.bc_ids[:n_bc]
    _check_logjoint(bb, da, "fitness_normal", {}, None)
    eng = bb.Engine(da, "fitness_normal", n_samples=3, dtype="f32", seed=2)
    eng.init_params(1)
    eng.step(3)
    m, s = eng.get_posterior()
    assert np.isfinite(m).all() and (s > 0).all()
    eng.close()


def test_zero_counts_and_large_counts(bb):
    da, _ = bb.synth.simulate("fitness_normal", 4, 12, 5, seed=6, mean_reads=0.7)       # many zero counts
    assert (np.asarray(da.bc_count) == 0).any()
    _check_logjoint(bb, da, "fitness_normal", {}, None)
    da2, _ = bb.synth.simulate("fitness_normal", 4, 12, 5, seed=6, mean_reads=2e7)      # counts ~ 10^8
    assert 5e7 < np.asarray(da2.bc_count).max() < 2 ** 31
    _check_logjoint(bb, da2, "fitness_normal", {}, None)


def test_c_abi_rejects_bad_arguments(bb):
    """Errors come back as BarBayError with the library's message (the Julia glue would raise error(msg))."""
    da, _ = bb.synth.simulate("fitness_normal", 4, 12, 5, seed=1)
    with pytest.raises(bb.BarBayError, match="samples_per_step"):
        bb.Engine(da, "fitness_normal", n_samples=0)
    with pytest.raises(bb.BarBayError, match="rows"):
        bb.Engine(da, "fitness_normal", {"s_bc_prior": np.ones((5, 2))})
    with pytest.raises(bb.BarBayError, match="> 0"):
        bb.Engine(da, "fitness_normal", {"s_pop_prior": [0.0, -1.0]})
    with pytest.raises(bb.BarBayError, match="single replicate"):
        bb.Engine(bb.synth.simulate("replicate_fitness_normal", 4, 8, 5, n_rep=2, seed=1)[0], "fitness_normal")
    big = bb.synth.simulate("fitness_normal", 2, 4, 5, seed=1)[0]
    big.bc_count = big.bc_count.copy()
    big.bc_count[0, 0] = 2 ** 31
    with pytest.raises(bb.BarBayError, match="2\\^31"):
        bb.Engine(big, "fitness_normal")
    eng = bb.Engine(da, "fitness_normal", n_samples=2)
    with pytest.raises(bb.BarBayError, match="length D"):
        eng.set_params(np.zeros(3), np.zeros(3))
    with pytest.raises(bb.BarBayError, match="TruncatedADAGrad or DecayedADAGrad"):
        eng.set_optimizer("adam")
    eng.close()


@pytest.mark.parametrize("persist", ["0", None])
@pytest.mark.parametrize("model", ["fitness_normal", "replicate_fitness_normal"])
@pytest.mark.parametrize("opt", ["truncated", "decayed"])
def test_checkpoint_resume_on_a_fresh_handle_equals_uninterrupted_run(bb, model, opt, persist, monkeypatch):
    """bb_get_state / bb_set_state carry theta, the accumulators AND the TruncatedADAGrad window (the reference's
    default optimiser): a run restored on a FRESH handle continues like the uninterrupted one, past the point where
    the restored window starts evicting (n = 4, 5 + 6 steps) -- bitwise with one launch pair per step
    (BB_PERSIST=0), to rounding with the default persistent step kernel (the first step of every launch takes its
    context from the tail kernel, the others from the in-kernel phases: a resumed run cuts the launches elsewhere)."""
    from helpers import load_fixture, rel_err
    if persist is None:
        monkeypatch.delenv("BB_PERSIST", raising=False)
    else:
        monkeypatch.setenv("BB_PERSIST", persist)
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    kw = dict(eta=0.1, tau=1.0, n=4) if opt == "truncated" else dict(eta=0.1, pre=1.0, post=0.9)

    def fresh():
        e = bb.Engine(da, model, n_samples=2, dtype="f64", seed=8)
        e.init_params(2)
        e.set_optimizer(opt, **kw)
        return e

    a = fresh()
    a.step(11)
    ref = a.get_params()
    a.close()
    b = fresh()
    b.step(5)
    state = b.get_state()
    b.close()
    c = fresh()
    c.step(3)                      # state of the fresh handle must be overwritten, not merged
    c.set_state(state)
    assert c.step_count == 5
    c.step(6)
    got = c.get_params()
    c.close()
    if persist == "0":
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    else:
        assert rel_err(got[0], ref[0]) < 1e-10 and rel_err(got[1], ref[1]) < 1e-10
