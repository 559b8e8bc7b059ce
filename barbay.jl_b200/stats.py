"""Naive priors: host-side mirror of ``BarBay.stats.naive_prior`` / ``naive_fitness``
(src/stats.jl:1175-1359, 1040-1106) -- the empirical priors every documented workflow feeds into the
models (docs/src/examples.md:121-140).  O(rows) numpy on top of ``utils.data_to_arrays``; the values
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
                neutral_col="neutral", rep_col=None, pseudocount: int = 1, mutate: bool = True) -> dict:
    """Returns ``{"s_pop_prior", "logσ_pop_prior", "logλ_prior"}`` -- vectors of prior means in the
    latent order of the models (population vectors: time fastest, then replicate; logλ: ``vec`` of the
    count array)."""
    if mutate:
        data[count_col] = data[count_col] + pseudocount                     # stats.jl:1185 (mutates the caller's frame)
        frame = data
    else:
        frame = data.assign(**{count_col: data[count_col] + pseudocount})
    da = _utils.data_to_arrays(frame, id_col=id_col, time_col=time_col, count_col=count_col,
                               neutral_col=neutral_col, rep_col=rep_col)
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
