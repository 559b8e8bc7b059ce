"""GPU parity of the packed / persistent step kernel (csrc/bb_step_kernel.cuh).

The kernel replaces pass2_kernel<FUSE> for the non-hierarchical models (fitness_normal, multienv): K samples
walked two at a time on packed fp32 arithmetic, bulk-copy (TMA 1-D) staging, and -- persistent mode -- several
ADVI steps per launch with the reduction, the peer exchange and the shared-latent phases inside the kernel.
Every variant must follow the oracle's ADVI trajectory on the same noise lattice (fp64 1e-8, fp32 the stated
2e-3 of the max-norm) and agree with the round-1 kernels up to reduction order.
"""
import numpy as np
import pytest

from helpers import load_fixture, oracle_problem, rel_err

pytestmark = pytest.mark.gpu

MODES = {
    # name: environment (BB_PERSIST = steps per persistent launch; 0 = one launch pair per step)
    "stepk": {"BB_PERSIST": "0"},
    "persist": {"BB_PERSIST": "4"},          # 9 steps = launches of 4 + 4 + 1: chunk boundaries are exercised
    "round1_fused": {"BB_NO_STEPK": "1"},
}


def _env(monkeypatch, mode):
    for k in ("BB_PERSIST", "BB_NO_STEPK", "BB_NO_FUSE"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("BB_STEPK_ODD", "1")       # odd K on one GPU defaults to the round-1 kernel: test the W = 1 one
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)


def _fixture_engine(bb, model, K, dtype):
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    return da, bb.Engine(da, model, n_samples=K, dtype=dtype, seed=1234)


@pytest.mark.parametrize("opt", ["decayed", "truncated"])
@pytest.mark.parametrize("K", [2, 8, 3])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("model", ["fitness_normal", "multienv_fitness_normal"])
def test_step_kernel_follows_oracle(bb, model, dtype, K, opt, monkeypatch):
    from oracle import advi_ref
    n_steps = 9
    kw = dict(eta=0.1, pre=1.0, post=0.9) if opt == "decayed" else dict(eta=0.1, tau=1.0, n=3)
    ref_opt = advi_ref.DecayedADAGrad(0.1, 1.0, 0.9) if opt == "decayed" else advi_ref.TruncatedADAGrad(0.1, 1.0, 3)
    res = {}
    for mode in MODES:
        _env(monkeypatch, mode)
        da, eng = _fixture_engine(bb, model, K, dtype)
        eng.init_params(5)
        mu0, om0 = eng.get_params()
        eng.set_optimizer(opt, **kw)
        eng.step(n_steps)
        res[mode] = eng.get_params()
        assert eng.step_count == n_steps
        eng.close()
    tr = advi_ref.advi_run(model, oracle_problem(da, model), n_steps, K, ref_opt, mu0, om0, seed=1234)
    # fp32 (measured over these cases, tests/_debug_trunc.py): DecayedADAGrad <= 2.3e-7; TruncatedADAGrad <= 4.5e-5 -- the
    # fp32 running window sum of the column latents keeps a cancellation residue of the evicted first steps.
    # (Until the population latents' update moved to double in the persistent kernel this needed 1e-2.)
    tol = 1e-8 if dtype == "f64" else (2e-6 if opt == "decayed" else 3e-4)
    for mode, (mu, om) in res.items():
        assert rel_err(mu, tr.mu) < tol and rel_err(om, tr.omega) < tol, (mode, rel_err(mu, tr.mu), rel_err(om, tr.omega))
    cross = 1e-10 if dtype == "f64" else tol
    for mode in ("persist", "round1_fused"):
        assert rel_err(res["stepk"][0], res[mode][0]) < cross and rel_err(res["stepk"][1], res[mode][1]) < cross, mode


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_step_kernel_many_tiles_and_groups(bb, dtype, monkeypatch):
    """6 000 barcodes: 48 tiles, more CTAs than one reduction group (32), partial last tile, neutral block with
    its chunked second sweep.  All kernel variants against the oracle; persistent launches split 3 + 3 + 1."""
    from oracle import advi_ref
    model, K, n_steps = "fitness_normal", 4, 7
    da, _ = bb.synth.simulate(model, n_neutral=70, n_bc=5931, n_time=5, seed=77)
    res = {}
    # fp64 at K = 4 keeps only two CTAs per SM resident: the engine would choose the round-1 kernels, the test insists
    monkeypatch.setenv("BB_STEPK_MIN_OCC", "1")
    for mode in MODES:
        _env(monkeypatch, mode)
        if mode == "persist":
            monkeypatch.setenv("BB_PERSIST", "3")
        eng = bb.Engine(da, model, n_samples=K, dtype=dtype, seed=99)
        eng.init_params(2)
        mu0, om0 = eng.get_params()
        eng.set_optimizer("decayed")
        eng.step(n_steps)
        res[mode] = eng.get_params()
        if mode == "persist":
            st = eng.persist_stats()
            assert st["tails"] == 4, st          # (3 - 1) + (3 - 1) + 0 in-kernel tails
        eng.close()
    tr = advi_ref.advi_run(model, oracle_problem(da, model), n_steps, K, advi_ref.DecayedADAGrad(), mu0, om0, seed=99)
    tol = 1e-8 if dtype == "f64" else 1e-4
    for mode, (mu, om) in res.items():
        assert rel_err(mu, tr.mu) < tol and rel_err(om, tr.omega) < tol, (mode, rel_err(mu, tr.mu), rel_err(om, tr.omega))


def test_step_kernel_matrix_priors_and_mixed_calls(bb, monkeypatch):
    """Matrix priors are staged by bulk copies too; a traced step (round-1 kernels with ELBO terms), a parity
    call and a state read in between must not disturb the pipelined partial sums of the step kernel."""
    from oracle import advi_ref
    model, K = "fitness_normal", 2
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    R = np.asarray(da.bc_count)
    priors = {"logλ_prior": np.column_stack([np.log(R.T.reshape(-1) + 1.0), np.full(R.size, 2.0)]),
              "s_pop_prior": np.column_stack([np.linspace(-0.1, 0.1, R.shape[0] - 1), np.full(R.shape[0] - 1, 0.5)])}
    _env(monkeypatch, "persist")
    eng = bb.Engine(da, model, priors, n_samples=K, dtype="f64", seed=3)
    eng.init_params(4)
    mu0, om0 = eng.get_params()
    eng.set_optimizer("decayed")
    eng.step(5)
    tr1 = eng.step(2, elbo_trace=True)
    _ = eng.elbo_grad(step=3)
    eng.step(3)
    mu, om = eng.get_params()
    eng.close()
    tr = advi_ref.advi_run(model, oracle_problem(da, model, priors), 10, K, advi_ref.DecayedADAGrad(), mu0, om0, seed=3)
    assert rel_err(mu, tr.mu) < 1e-8 and rel_err(om, tr.omega) < 1e-8
    assert rel_err(tr1, np.asarray(tr.elbo)[5:7]) < 1e-9
