"""Oracle for the naive priors: loop-level transliteration of BarBay.stats.naive_prior and
naive_fitness (src/stats.jl:1175-1359, 1040-1106).  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED
(test/stats_tests.jl:137-224 checks keys, lengths and NaN-freeness only)."""
from __future__ import annotations

import math

import numpy as np
import pandas as pd

from .packer_ref import data_to_arrays_ref


def _mean_std_noninf(values):
    v = [x for x in values if not math.isinf(x)]
    m = sum(v) / len(v)
    s = math.sqrt(sum((x - m) ** 2 for x in v) / (len(v) - 1))
    return m, s


def naive_prior_ref(data: pd.DataFrame, rep_col=None, pseudocount=1, **cols) -> dict:
    frame = data.copy()
    ccol = cols.get("count_col", "count")
    frame[ccol] = frame[ccol] + pseudocount                                  # stats.jl:1185
    da = data_to_arrays_ref(frame, rep_col=rep_col, **cols)
    N = da.n_neutral
    blocks = []          # (counts T x B, totals T) per replicate, in the order the reference flattens them
    if isinstance(da.bc_count, list):
        blocks = [(np.asarray(R), np.asarray(t)) for R, t in zip(da.bc_count, da.bc_total)]
    elif np.asarray(da.bc_count).ndim == 3:
        R = np.asarray(da.bc_count)
        blocks = [(R[:, :, r], np.asarray(da.bc_total)[:, r]) for r in range(R.shape[2])]
    else:
        blocks = [(np.asarray(da.bc_count), np.asarray(da.bc_total))]
    s_pop, lsig, loglam = [], [], []
    for R, tot in blocks:
        T = R.shape[0]
        for t in range(T - 1):
            ratios = [math.log((R[t + 1, n] / tot[t + 1]) / (R[t, n] / tot[t])) for n in range(N)]   # :1207-1210
            m, s = _mean_std_noninf(ratios)
            s_pop.append(-m)                                                 # :1298
            lsig.append(-s)                                                  # :1338 (quirk: -std, not log(std))
        for b in range(R.shape[1]):
            for t in range(T):
                loglam.append(math.log(R[t, b]))                             # :1347 log.(bc_count)[:]
    return {"s_pop_prior": np.asarray(s_pop), "logσ_pop_prior": np.asarray(lsig), "logλ_prior": np.asarray(loglam)}


def naive_fitness_ref(data: pd.DataFrame, pseudocount=1) -> pd.DataFrame:
    d = data[["barcode", "time", "count", "neutral"]].copy()
    d["count"] = d["count"] + pseudocount
    tot = d.groupby("time")["count"].sum()
    d["freq"] = d["count"] / d["time"].map(tot)
    rows = []
    for bc, g in d.groupby("barcode", sort=False):
        lf = np.diff(np.log(g["freq"].to_numpy()))                           # row order of the frame, as the reference
        for t, v in zip(g["time"].to_numpy()[1:], lf):
            rows.append((bc, t, v, bool(g["neutral"].iloc[0])))
    dl = pd.DataFrame(rows, columns=["barcode", "time", "logf", "neutral"])
    st = dl[dl.neutral].groupby("time")["logf"].mean()
    dl["logf_norm"] = dl["logf"] - dl["time"].map(st)
    out = dl[~dl.neutral].groupby("barcode", sort=False)["logf_norm"].mean().reset_index()
    return out.rename(columns={"logf_norm": "fitness"})
