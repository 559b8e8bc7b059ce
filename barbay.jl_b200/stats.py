"""Host-side mirror of ``BarBay.stats``: the naive priors ``naive_prior`` / ``naive_fitness``
(src/stats.jl:1175-1359, 1040-1106) -- the empirical priors every documented workflow feeds into the
models (docs/src/examples.md:121-140) -- and, at the end of the file, the posterior predictive checks
on a DataFrame of posterior draws (src/stats.jl:35-996).  O(rows) numpy on top of ``utils.data_to_arrays``; the values
go to the kernels as matrix priors.  SURVEY.md §8f rank 3 ("next" row): host arithmetic only, it is
not on the per-step path.

Reference quirks reproduced on purpose: ``logσ_pop_prior`` is ``-std(logfreq)``
(not ``log(std)``) (stats.jl:1338) and the caller's count column is incremented by ``pseudocount``
in place (stats.jl:1185); pass ``mutate=False`` to leave the frame alone.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import utils as _utils


def _finite_mean_std(x: np.ndarray, axis: int):
    """mean / std (ddof = 1, StatsBase.std) over the entries that are not +-Inf (stats.jl:1260-1262)."""
    ok = ~np.isinf(x)
    n = ok.sum(axis=axis)
    xs = np.where(ok, x, 0.0)
    mean = xs.sum(axis=axis) / n
    dev = np.where(ok, x - np.expand_dims(mean, axis), 0.0)
    std = np.sqrt((dev ** 2).sum(axis=axis) / (n - 1))
    return mean, std


def naive_prior(data: pd.DataFrame, *, id_col="barcode", time_col="time", count_col="count",
                neutral_col="neutral", rep_col=None, pseudocount: int = 1, mutate: bool = True,
                device: int | None = None) -> dict:
    """Returns ``{"s_pop_prior", "logσ_pop_prior", "logλ_prior"}`` -- vectors of prior means in the
    latent order of the models (population vectors: time fastest, then replicate; logλ: ``vec`` of the
    count array).  ``device`` (a CUDA ordinal, -1 = current): the arithmetic runs on the GPU from the
    packed count array (``bb_naive_prior``, csrc/bb_naive.cuh) instead of in numpy."""
    if mutate:
        data[count_col] = data[count_col] + pseudocount                     # stats.jl:1185 (mutates the caller's frame)
        frame = data
    else:
        frame = data.assign(**{count_col: data[count_col] + pseudocount})
    da = _utils.data_to_arrays(frame, id_col=id_col, time_col=time_col, count_col=count_col,
                               neutral_col=neutral_col, rep_col=rep_col)
    if device is not None:
        return naive_prior_packed(da, device=device)
    N = da.n_neutral
    if isinstance(da.bc_count, list):                                       # unequal T per replicate :1227-1249
        means, stds, loglam = [], [], []
        for R, tot in zip(da.bc_count, da.bc_total):
            f = R / np.asarray(tot)[:, None]
            lr = np.log(f[1:, :N] / f[:-1, :N])
            m, s = _finite_mean_std(lr, axis=1)
            means.append(m); stds.append(s)
            loglam.append(np.log(R.astype(np.float64)).T.reshape(-1))        # log.(R)[:] column-major
        mean, std, logl = np.concatenate(means), np.concatenate(stds), np.concatenate(loglam)
    else:
        R = np.asarray(da.bc_count)
        tot = np.asarray(da.bc_total)
        f = R / (tot[:, None] if R.ndim == 2 else tot[:, None, :])
        lr = np.log(f[1:, :N] / f[:-1, :N])                                 # (T-1) x N [x R]
        m, s = _finite_mean_std(lr, axis=1)                                 # (T-1) [x R]
        mean = m.T.reshape(-1) if m.ndim == 2 else m                        # logfreq_mean[:] column-major
        std = s.T.reshape(-1) if s.ndim == 2 else s
        Rf = np.log(R.astype(np.float64))
        logl = Rf.T.reshape(-1) if R.ndim == 2 else Rf.transpose(2, 1, 0).reshape(-1)
    return {"s_pop_prior": -mean, "logσ_pop_prior": -std, "logλ_prior": logl}


def naive_prior_packed(da, *, device: int = -1) -> dict:
    """``naive_prior`` of already packed data (``utils.data_to_arrays`` of the frame WITH the pseudocount
    added) on the GPU: one H2D copy of the counts, three kernels (exact integer totals, ``log`` of every
    count, mean / sd of the neutral log-frequency ratios), three D2H copies.  No CPU fallback."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    if isinstance(da.bc_count, (list, tuple)):
        mats = [np.asarray(m, dtype=np.int64) for m in da.bc_count]
        n_time = [m.shape[0] for m in mats]
        flat = np.concatenate([m.T.reshape(-1) for m in mats])              # each T_r x B column-major
    else:
        R = np.asarray(da.bc_count, dtype=np.int64)
        n_time = [R.shape[0]] * (1 if R.ndim == 2 else R.shape[2])
        flat = R.T.reshape(-1) if R.ndim == 2 else R.transpose(2, 1, 0).reshape(-1)
    flat = np.ascontiguousarray(flat)
    nt = np.asarray(n_time, dtype=np.int32)
    n_pop = int(nt.sum()) - nt.size
    s_pop, lsig, logl = np.empty(n_pop), np.empty(n_pop), np.empty(flat.size)
    D = C.POINTER(C.c_double)
    rc = lib.bb_naive_prior(flat.ctypes.data_as(C.POINTER(C.c_int64)), nt.size, nt.ctypes.data_as(C.POINTER(C.c_int32)),
                            int(da.n_neutral), int(da.n_bc), int(device), s_pop.ctypes.data_as(D),
                            lsig.ctypes.data_as(D), logl.ctypes.data_as(D))
    if rc != 0:
        raise _lib.BarBayError(lib.bb_last_error(None).decode())
    return {"s_pop_prior": s_pop, "logσ_pop_prior": lsig, "logλ_prior": logl}


def prior_matrices(prior: dict, s_pop_std: float = 0.05, logsig_pop_std: float = 1.0, loglam_std: float = 3.0) -> dict:
    """The n x 2 ``[mean std]`` matrices the documented workflow builds from ``naive_prior``
    (docs/src/examples.md:121-140: ``hcat(naive[:s_pop_prior], repeat([0.05], ...))`` etc.)."""
    def mat(v, sd):
        v = np.asarray(v, dtype=np.float64)
        return np.column_stack([v, np.full(v.size, sd)])
    return {"s_pop_prior": mat(prior["s_pop_prior"], s_pop_std),
            "logσ_pop_prior": mat(prior["logσ_pop_prior"], logsig_pop_std),
            "logλ_prior": mat(prior["logλ_prior"], loglam_std)}


def naive_fitness(data: pd.DataFrame, *, id_col="barcode", time_col="time", count_col="count",
                  neutral_col="neutral", pseudocount: int = 1) -> pd.DataFrame:
    """Mean over time of the log-frequency ratio of each mutant barcode minus the neutrals' mean
    log-frequency ratio at that time (stats.jl:1040-1106).  Returns columns ``[id_col, "fitness"]``,
    one row per mutant barcode in group (first-appearance) order."""
    frame = data[[id_col, time_col, count_col, neutral_col]].copy()          # :1049 copies: no mutation here
    frame[count_col] = frame[count_col] + pseudocount
    da = _utils.data_to_arrays(frame, id_col=id_col, time_col=time_col, count_col=count_col,
                               neutral_col=neutral_col)
    R = np.asarray(da.bc_count, dtype=np.float64)
    f = R / R.sum(axis=1, keepdims=True)
    logf = np.diff(np.log(f), axis=0)                                        # (T-1) x B
    neutral_mean = logf[:, :da.n_neutral].mean(axis=1)
    fitness = (logf[:, da.n_neutral:] - neutral_mean[:, None]).mean(axis=0)
    # the reference groups by id in order of first appearance over the whole frame
    order = pd.unique(frame.loc[~frame[neutral_col], id_col])
    pos = {b: i for i, b in enumerate(da.bc_ids)}
    idx = [pos[b] for b in order]
    return pd.DataFrame({id_col: list(order), "fitness": fitness[idx]})


# ------------------------------------------------------------------------------------------------------
# Posterior predictive checks: host mirrors of the DataFrame methods of
# src/stats.jl:35-996.  `df` holds one posterior draw per row (e.g. draws of the fitted mean-field
# Gaussian); the draws of the predictive Normal / LogNormal use numpy's Generator, so -- exactly like
# the reference with Julia's global RNG -- results are reproducible only for a caller-supplied `rng`.
def _sorted_vars(df: pd.DataFrame, pattern: str) -> list:
    """``sort(names(df)[occursin.(pattern, names(df))])`` (stats.jl:165-169): code-point order."""
    return sorted(c for c in df.columns if pattern in str(c))


def _flatten(ppc: np.ndarray, flatten: bool) -> np.ndarray:
    """``vcat(collect(eachslice(ppc, dims=3))...)`` (stats.jl:208-212): the n_ppc slices stacked by rows."""
    if not flatten:
        return ppc
    return np.concatenate([ppc[:, :, k] for k in range(ppc.shape[2])], axis=0)


def matrix_quantile_range(quantile, matrix, dims: int = 2) -> np.ndarray:
    """Symmetric quantile ranges of ``matrix`` along ``dims`` (1-based, as in the reference,
    stats.jl:55-88): ``out[:, i, 0/1]`` = the ``(1-q)/2`` and ``1-(1-q)/2`` quantiles (StatsBase
    default = linear interpolation = numpy default)."""
    q = np.asarray(quantile, dtype=np.float64)
    if ((q < 0.0) | (q > 1.0)).any():
        raise _err("All quantiles must be between zero and one")
    if dims not in (1, 2):
        raise _err("Dimensions should match a Matrix dimensiosn, i.e., 1 or 2")
    m = np.asarray(matrix)
    # eachslice(matrix, dims=dims): one slice per index of `dims` (dims = 2: the columns), one quantile pair
    # per slice; the output has size(matrix, dims) rows (stats.jl:69-72: `op_dims` evaluates to `dims`)
    axis = 0 if dims == 2 else 1
    n_out = m.shape[1] if dims == 2 else m.shape[0]
    out = np.empty((n_out, q.size, 2), dtype=np.result_type(m.dtype, np.float64))
    for i, qi in enumerate(q):
        out[:, i, 0] = np.quantile(m, (1.0 - qi) / 2.0, axis=axis)
        out[:, i, 1] = np.quantile(m, 1.0 - (1.0 - qi) / 2.0, axis=axis)
    return out


def _err(msg: str):
    from ._lib import BarBayError
    return BarBayError(msg)


def logfreq_ratio_bc_ppc(df: pd.DataFrame, n_ppc: int, *, param: dict | None = None, flatten: bool = True,
                         rng: np.random.Generator | None = None) -> np.ndarray:
    """``log(f_{t+1}/f_t) ~ N(s⁽ᵐ⁾ - s̄_t, exp(σ⁽ᵐ⁾))`` for every posterior draw (stats.jl:377-418).
    Returns ``(n_draws * n_ppc) x n_times`` (flattened) or ``n_draws x n_times x n_ppc``."""
    p = {"bc_mean_fitness": "s⁽ᵐ⁾", "bc_std_fitness": "σ⁽ᵐ⁾", "population_mean_fitness": "s̲ₜ"}
    p.update(param or {})
    rng = rng or np.random.default_rng()
    mean_vars = _sorted_vars(df, p["population_mean_fitness"])
    s = df[p["bc_mean_fitness"]].to_numpy(np.float64)
    sd = np.exp(df[p["bc_std_fitness"]].to_numpy(np.float64))
    ppc = np.empty((len(df), len(mean_vars), n_ppc))
    for i, var in enumerate(mean_vars):
        mu = s - df[var].to_numpy(np.float64)
        ppc[:, i, :] = mu[:, None] + sd[:, None] * rng.standard_normal((len(df), n_ppc))
    return _flatten(ppc, flatten)


def logfreq_ratio_popmean_ppc(df: pd.DataFrame, n_ppc: int, *, param: dict | None = None, flatten: bool = True,
                              rng: np.random.Generator | None = None) -> np.ndarray:
    """Neutral lineages: ``log(f_{t+1}/f_t) ~ N(-s̄_t, exp(σ_t))`` (stats.jl:571-621)."""
    p = {"population_mean_fitness": "sₜ", "population_std_fitness": "σₜ"}
    p.update(param or {})
    rng = rng or np.random.default_rng()
    mean_vars = _sorted_vars(df, p["population_mean_fitness"])
    std_vars = _sorted_vars(df, p["population_std_fitness"])
    if len(mean_vars) != len(std_vars):
        raise _err("The number of mean and standard deviation variables does not match")
    ppc = np.empty((len(df), len(mean_vars), n_ppc))
    for i, (mv, sv) in enumerate(zip(mean_vars, std_vars)):
        mu = -df[mv].to_numpy(np.float64)
        sd = np.exp(df[sv].to_numpy(np.float64))
        ppc[:, i, :] = mu[:, None] + sd[:, None] * rng.standard_normal((len(df), n_ppc))
    return _flatten(ppc, flatten)


def logfreq_ratio_multienv_ppc(df: pd.DataFrame, n_ppc: int, envs, *, param: dict | None = None,
                               flatten: bool = True, rng: np.random.Generator | None = None) -> np.ndarray:
    """Multi-environment version: the ratio ``t -> t+1`` uses the fitness of the environment of time
    ``t+1`` (stats.jl:789-864); ``envs`` lists the environment of every time point."""
    p = {"bc_mean_fitness": "s̲⁽ᵐ⁾", "bc_std_fitness": "σ̲⁽ᵐ⁾", "population_mean_fitness": "s̲ₜ"}
    p.update(param or {})
    rng = rng or np.random.default_rng()
    env_unique = list(dict.fromkeys(envs))                                  # unique(): first appearance
    env_idx = [env_unique.index(e) for e in envs]                           # indexin (0-based here)
    mean_vars = _sorted_vars(df, p["population_mean_fitness"])
    s_vars = _sorted_vars(df, p["bc_mean_fitness"])
    sd_vars = _sorted_vars(df, p["bc_std_fitness"])
    if len(s_vars) != len(env_unique) or len(sd_vars) != len(env_unique):
        raise _err("# of mutant-related variables does not match # of environments")
    if len(envs) != len(mean_vars) + 1:
        raise _err("Number of given environments does not match time points in chain")
    ppc = np.empty((len(df), len(mean_vars), n_ppc))
    for i, var in enumerate(mean_vars):
        e = env_idx[i + 1]
        mu = df[s_vars[e]].to_numpy(np.float64) - df[var].to_numpy(np.float64)
        sd = np.exp(df[sd_vars[e]].to_numpy(np.float64))
        ppc[:, i, :] = mu[:, None] + sd[:, None] * rng.standard_normal((len(df), n_ppc))
    return _flatten(ppc, flatten)


def freq_bc_ppc(df: pd.DataFrame, n_ppc: int, *, param: dict | None = None, model: str = "lognormal",
                flatten: bool = True, rng: np.random.Generator | None = None) -> np.ndarray:
    """Barcode frequency trajectories ``f_{t+1} = f_t * LogNormal(s⁽ᵐ⁾ - s̄_t, σ)`` from the initial
    frequency column (stats.jl:152-213); ``model="normal"`` exponentiates the std column first."""
    p = {"bc_mean_fitness": "s⁽ᵐ⁾", "bc_std_fitness": "σ⁽ᵐ⁾", "bc_freq": "f̲⁽ᵐ⁾[1]",
         "population_mean_fitness": "s̲ₜ"}
    p.update(param or {})
    if model not in ("lognormal", "normal"):
        raise _err("model must be :normal or :lognormal")
    rng = rng or np.random.default_rng()
    mean_vars = _sorted_vars(df, p["population_mean_fitness"])
    s = df[p["bc_mean_fitness"]].to_numpy(np.float64)
    sd = df[p["bc_std_fitness"]].to_numpy(np.float64)
    if model == "normal":
        sd = np.exp(sd)
    ppc = np.empty((len(df), len(mean_vars) + 1, n_ppc))
    ppc[:, 0, :] = df[p["bc_freq"]].to_numpy(np.float64)[:, None]
    for i, var in enumerate(mean_vars):
        mu = s - df[var].to_numpy(np.float64)
        ppc[:, i + 1, :] = ppc[:, i, :] * np.exp(mu[:, None] + sd[:, None] * rng.standard_normal((len(df), n_ppc)))
    return _flatten(ppc, flatten)
