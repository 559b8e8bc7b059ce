"""Oracle packer: naive transliteration of BarBay.utils.data_to_arrays.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED (the
reference's tests never compare packed values, test/utils_tests.jl).

Follows src/utils.jl:81-382 (_extract_timepoints, _process_*), :409-920
(_extract_R methods) and :996-1033 (data_to_arrays) statement by statement,
including the per-(id, rep) boolean-mask loops of the equal-T replicate path
(:208-220, :252-264), so it is O(ids * reps * rows) like the reference and only
meant for small frames.

DataFrames.jl ``groupby`` (sort=nothing) group order, restated from DataFrames
1.x ``row_group_slots!``: String / generic keys -> order of first appearance;
Integer keys whose value range is narrow (max - min + 1 <= 2 * nrows) -> value
order.  ``unique`` -> first appearance, ``sort(unique(...))`` -> sorted.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import pandas as pd


@dataclass
class DataArraysRef:
    bc_count: Any
    bc_total: Any
    n_neutral: int
    n_bc: int
    bc_ids: list
    neutral_ids: list
    envs: Any
    n_env: int
    n_rep: int
    n_time: Any
    genotypes: Any
    n_geno: int


def _unique(values) -> list:
    seen, out = set(), []
    for v in values:
        if v not in seen:
            seen.add(v)
            out.append(v)
    return out


def _group_order(values: pd.Series) -> list:
    """Group key order of DF.groupby(df, col) for a single key column."""
    uniq = _unique(values.tolist())
    if pd.api.types.is_integer_dtype(values.dtype) and len(values) > 0:
        lo, hi = int(values.min()), int(values.max())
        if hi - lo + 1 <= 2 * len(values):
            return sorted(uniq)
    return uniq


def _process_single(data, mask, id_col, time_col, count_col, n_time, what):
    sub = data[mask]
    ids = _group_order(sub[id_col])
    sizes = [int((sub[id_col] == i).sum()) for i in ids]
    if any(s != n_time for s in sizes):
        raise ValueError(f"Not all {what} barcodes have reported counts in all time points.")
    R = np.empty((n_time, len(ids)), dtype=np.int64)
    for j, i in enumerate(ids):
        d = sub[sub[id_col] == i].sort_values(time_col, kind="stable")
        R[:, j] = d[count_col].to_numpy()
    return R, ids


def _process_multi(data, mask, id_col, time_col, count_col, rep_col, n_time, sort_keys):
    sub = data[mask]
    ids = _unique(sub[id_col].tolist())
    reps = _unique(sub[rep_col].tolist())
    if sort_keys:                      # mutants: sort(unique(...)) utils.jl:242-244
        ids, reps = sorted(ids), sorted(reps)
    R = np.empty((n_time, len(ids), len(reps)), dtype=np.int64)
    for j, i in enumerate(ids):
        for k, rep in enumerate(reps):
            d = sub[(sub[id_col] == i) & (sub[rep_col] == rep)].sort_values(time_col, kind="stable")
            R[:, j, k] = d[count_col].to_numpy()
    return R, ids


def _process_multi_varying(groups, neutral, neutral_col, id_col, time_col, count_col, n_rep_time):
    Rs, ids0 = [], []
    for rep, d_rep in enumerate(groups):
        mask = d_rep[neutral_col].to_numpy() if neutral else ~d_rep[neutral_col].to_numpy()
        what = "neutral" if neutral else "mutant"
        try:
            R, ids = _process_single(d_rep, mask, id_col, time_col, count_col, n_rep_time[rep], what)
        except ValueError:
            raise ValueError(
                f"Not all {what} barcodes have reported counts in all time points for replicate {rep + 1}.")
        if rep == 0:
            ids0 = ids
        Rs.append(R)
    return Rs, ids0


def data_to_arrays_ref(data: pd.DataFrame, id_col="barcode", time_col="time", count_col="count",
                       neutral_col="neutral", rep_col=None, env_col=None, genotype_col=None) -> DataArraysRef:
    for c in (id_col, time_col, count_col, neutral_col, rep_col, env_col, genotype_col):
        if c is not None and c not in data.columns:
            raise ValueError(f"Column {c} does not exist in the dataframe")
    if data[neutral_col].dtype != np.bool_:
        raise ValueError(f"Column {neutral_col} must be of type Bool")
    neutral = data[neutral_col].to_numpy()
    timepoints = sorted(_unique(data[time_col].tolist()))

    if rep_col is None:
        Rn, neutral_ids = _process_single(data, neutral, id_col, time_col, count_col, len(timepoints), "neutral")
        Rm, bc_ids = _process_single(data, ~neutral, id_col, time_col, count_col, len(timepoints), "mutant")
        R = np.concatenate([Rn, Rm], axis=1)
        nt = R.sum(axis=1)
        out = DataArraysRef(R, nt, len(neutral_ids), len(bc_ids), bc_ids, neutral_ids,
                            "env1", 1, 1, len(timepoints), "N/A", 0)
    else:
        rep_keys = _group_order(data[rep_col])
        groups = [data[data[rep_col] == k] for k in rep_keys]
        n_rep = len(groups)
        n_rep_time = [len(_unique(g[time_col].tolist())) for g in groups]
        if len(set(n_rep_time)) == 1:
            Rn, neutral_ids = _process_multi(data, neutral, id_col, time_col, count_col, rep_col,
                                             len(timepoints), sort_keys=False)
            Rm, bc_ids = _process_multi(data, ~neutral, id_col, time_col, count_col, rep_col,
                                        len(timepoints), sort_keys=True)
            R = np.concatenate([Rn, Rm], axis=1)
            nt = R.sum(axis=1)                                 # T x n_rep
        else:
            Rn, neutral_ids = _process_multi_varying(groups, True, neutral_col, id_col, time_col,
                                                     count_col, n_rep_time)
            Rm, bc_ids = _process_multi_varying(groups, False, neutral_col, id_col, time_col,
                                                count_col, n_rep_time)
            R = [np.concatenate([a, b], axis=1) for a, b in zip(Rn, Rm)]
            nt = [r.sum(axis=1) for r in R]
        out = DataArraysRef(R, nt, len(neutral_ids), len(bc_ids), bc_ids, neutral_ids,
                            "env1", 1, n_rep, n_rep_time, "N/A", 0)

    if env_col is not None:
        def env_list(df):
            pairs = df[[time_col, env_col]].drop_duplicates()
            return pairs.sort_values(time_col, kind="stable")[env_col].tolist()
        if rep_col is None:
            envs = env_list(data)
            n_env = len(_unique(envs))
        else:
            envs_r = [env_list(g) for g in groups]
            n_env = len(_unique([e for es in envs_r for e in es]))
            envs = envs_r[0] if all(es == envs_r[0] for es in envs_r) else envs_r
        out.envs, out.n_env = envs, n_env

    if genotype_col is not None:
        geno = {}
        for i, g in zip(data[id_col].tolist(), data[genotype_col].tolist()):
            geno[i] = g
        out.genotypes = [geno[m] for m in out.bc_ids]
        out.n_geno = len(_unique(out.genotypes))
    return out
