"""Golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle):
CPU -- the oracle still reproduces them and the packer reproduces the packed counts;
GPU -- the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from helpers import FIXTURES, ROOT, load_fixture, oracle_problem, rel_err

GOLD = os.path.join(ROOT, "tests", "golden")


def _load(model):
    return np.load(os.path.join(GOLD, f"{model}.npz"))


@pytest.mark.parametrize("model", list(FIXTURES))
def test_oracle_reproduces_golden(bb, model):
    from oracle import advi_ref, model_ref, philox_ref
    g = _load(model)
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    assert np.array_equal(np.asarray(da.bc_count), g["bc_count"])          # packing is bit-exact
    prob = oracle_problem(da, model)
    for k in range(g["z"].shape[0]):
        lp, gr = model_ref.logjoint_and_grad(model, g["z"][k], prob)
        assert abs(lp - g["logp"][k]) <= 1e-12 * abs(lp) and rel_err(gr, g["grad"][k]) < 1e-12
    elbo, gm, go, _ = advi_ref.elbo_value_and_grad(model, prob, g["mu"], g["omega"], g["eps"])
    assert abs(elbo - float(g["elbo"])) <= 1e-12 * abs(elbo)
    assert rel_err(gm, g["g_mu"]) < 1e-12 and rel_err(go, g["g_omega"]) < 1e-12
    assert np.allclose(philox_ref.noise(model, prob, 2, 3, 7), g["lattice_step3_seed7"], rtol=0, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("model", list(FIXTURES))
def test_cuda_reproduces_golden(bb, model):
    g = _load(model)
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    K = g["z"].shape[0]
    eng = bb.Engine(da, model, n_samples=K, dtype="f64", seed=7)
    logp, grad = eng.logjoint_grad(g["z"])
    assert np.all(np.abs(logp - g["logp"]) <= 1e-9 * np.abs(g["logp"])) and rel_err(grad, g["grad"]) < 1e-9
    eng.set_params(g["mu"], g["omega"])
    elbo, gm, go = eng.elbo_grad(g["eps"])
    assert abs(elbo - float(g["elbo"])) <= 1e-9 * abs(elbo)
    assert rel_err(gm, g["g_mu"]) < 1e-9 and rel_err(go, g["g_omega"]) < 1e-9
    assert np.max(np.abs(eng.get_noise(3) - g["lattice_step3_seed7"])) < 1e-12
    for name, kw in (("decayed", dict(eta=0.1, pre=1.0, post=0.9)), ("truncated", dict(eta=0.1, tau=1.0, n=2))):
        eng.set_params(g["mu"], g["omega"])
        eng.set_optimizer(name, **kw)
        for s in range(g["noise"].shape[0]):
            eng.step_with_noise(g["noise"][s])
        mu, om = eng.get_params()
        assert rel_err(mu, g[f"mu_{name}"]) < 1e-8 and rel_err(om, g[f"omega_{name}"]) < 1e-8
    eng.close()
