"""Shared test helpers: fixture loading and the bridge from the product's DataArrays to the
oracle's problem dict.  (tests/ is allowed to import oracle/.)"""
import os

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")

FIXTURES = {
    "fitness_normal": ("data001_single.csv", {}),
    "replicate_fitness_normal": ("data002_hier-rep.csv", {"rep_col": "rep"}),
    "multienv_fitness_normal": ("data003_multienv.csv", {"env_col": "env"}),
    "genotype_fitness_normal": ("data004_multigen.csv", {"genotype_col": "genotype"}),
}


def load_fixture(model):
    fname, cols = FIXTURES[model]
    return pd.read_csv(os.path.join(DATA, fname)), cols


def uneven_replicates(df):
    """test/vi_tests.jl:102-105: drop the last time point of the last replicate."""
    return df[(df.rep != df.rep.max()) | (df.time != df.time.max())].reset_index(drop=True)


def oracle_problem(da, model_name, priors=None, corrected=True):
    prob = {"bc_count": da.bc_count, "bc_total": da.bc_total, "n_neutral": da.n_neutral, "n_bc": da.n_bc,
            "priors": priors, "corrected": corrected}
    if "multienv" in model_name:
        prob["envs"] = da.envs
    if "genotype" in model_name:
        prob["genotypes"] = da.genotypes
    return prob


def plausible_latents(layout, da, rng, K):
    """z draws near the data: log-lambda around log(count + 1), everything else small."""
    from barbay_b200 import model as M
    D = layout.n_latent
    z = 0.3 * rng.standard_normal((K, D))
    lam = next(g for g in layout.groups if g.name == M.V_LOGLAM)
    if isinstance(da.bc_count, list):
        flat = np.concatenate([np.asarray(m).T.reshape(-1) for m in da.bc_count])
    else:
        R = np.asarray(da.bc_count)
        flat = R.T.reshape(-1) if R.ndim == 2 else R.transpose(2, 1, 0).reshape(-1)
    z[:, lam.start:lam.start + lam.length] += np.log(flat + 1.0)[None, :]
    for g in layout.groups:
        if g.name in (M.V_LOGSIG_POP, M.V_LOGSIG_BC):
            z[:, g.start:g.start + g.length] -= 1.0
        if g.name == M.V_LOGTAU:
            z[:, g.start:g.start + g.length] -= 2.0
    return z


def plausible_theta(layout, da, rng):
    mu = plausible_latents(layout, da, rng, 1)[0]
    omega = -2.0 + 0.3 * rng.standard_normal(layout.n_latent)
    return mu, omega


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
