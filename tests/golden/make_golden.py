"""Generates tests/golden/*.npz from the oracle (NOT from the reference: Julia is unavailable, parity
is unpinned -- see oracle/__init__.py).  The vectors freeze the oracle's answers on the reference's
four fixtures so that (i) the oracle cannot drift silently and (ii) the GPU box, which has no
/root/reference, compares the CUDA path against committed numbers as well as against a live oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import barbay_b200 as bb  # noqa: E402  (host-side packer / layout only; no GPU needed)
from helpers import FIXTURES, load_fixture, oracle_problem, plausible_latents, plausible_theta  # noqa: E402
from oracle import advi_ref, model_ref, philox_ref  # noqa: E402

K, N_STEPS, SEED = 2, 3, 20261018


def main():
    for model in FIXTURES:
        df, cols = load_fixture(model)
        da = bb.utils.data_to_arrays(df, **cols)
        lay = bb.model.var_groups(bb.model.resolve(model), da.n_time, da.n_rep, da.n_neutral, da.n_bc, da.n_env,
                                  da.n_geno)
        prob = oracle_problem(da, model)
        rng = np.random.default_rng(SEED)
        z = plausible_latents(lay, da, rng, K)
        logp = np.empty(K)
        grad = np.empty((K, lay.n_latent))
        for k in range(K):
            logp[k], grad[k] = model_ref.logjoint_and_grad(model, z[k], prob)
        mu, omega = plausible_theta(lay, da, rng)
        eps = rng.standard_normal((K, lay.n_latent))
        elbo, g_mu, g_om, _ = advi_ref.elbo_value_and_grad(model, prob, mu, omega, eps)
        noise = rng.standard_normal((N_STEPS, K, lay.n_latent))
        out = dict(z=z, logp=logp, grad=grad, mu=mu, omega=omega, eps=eps, elbo=elbo, g_mu=g_mu, g_omega=g_om,
                   noise=noise, lattice_step3_seed7=philox_ref.noise(model, prob, K, 3, 7),
                   bc_count=np.asarray(da.bc_count))
        for name, opt in (("decayed", advi_ref.DecayedADAGrad(0.1, 1.0, 0.9)),
                          ("truncated", advi_ref.TruncatedADAGrad(0.1, 1.0, 2))):
            tr = advi_ref.advi_run(model, prob, N_STEPS, K, opt, mu, omega, eps_fn=lambda s: noise[s])
            out[f"mu_{name}"], out[f"omega_{name}"] = tr.mu, tr.omega
        np.savez_compressed(os.path.join(HERE, f"{model}.npz"), **out)
        print(model, lay.n_latent, logp)


if __name__ == "__main__":
    main()
