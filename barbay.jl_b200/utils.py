"""Data boundary: host-side mirror of ``BarBay.utils``.

``data_to_arrays`` reproduces utils.data_to_arrays (src/utils.jl:996-1033) and
its eight ``_extract_R`` methods (:409-920) -- same validations, same column /
replicate order, bit-exact Int64 counts -- in O(rows log rows) (the reference's
equal-T replicate path is O(ids * reps * rows), utils.jl:208-220, and cannot
reach the BASELINE sizes).  ``advi_to_df`` reproduces utils.advi_to_df
(:1409-1462) and its helpers (:1042-1343): it consumes exactly ``q.dist.m``,
``q.dist.σ``, ``q.transform.ranges_out`` and the variable names.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import pandas as pd

from . import model as _model
from ._lib import BarBayError


@dataclass
class DataArrays:
    """src/utils.jl:48-61."""
    bc_count: Any        # (T, B) | (T, B, R) int64 ndarray | list of (T_r, B)
    bc_total: Any        # (T,) | (T, R) | list of (T_r,)
    n_neutral: int
    n_bc: int
    bc_ids: list
    neutral_ids: list
    envs: Any            # "env1" | list | list of lists
    n_env: int
    n_rep: int
    n_time: Any          # int | list[int]
    genotypes: Any       # "N/A" | list
    n_geno: int


# ---------------------------------------------------------------------------
# group-order helpers (DataFrames.jl semantics)
# ---------------------------------------------------------------------------
def _first_appearance_codes(values: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """codes[i] = index of values[i] among the uniques in order of first appearance."""
    codes, uniques = pd.factorize(values, sort=False)
    return codes.astype(np.int64), np.asarray(uniques, dtype=object)


def _sorted_codes(values: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    codes, uniques = pd.factorize(values, sort=True)
    return codes.astype(np.int64), np.asarray(uniques, dtype=object)


def _groupby_codes(values: pd.Series) -> tuple[np.ndarray, np.ndarray]:
    """Group order of ``DF.groupby(df, col)`` (sort=nothing): first appearance, except
    Integer keys with a narrow value range, which DataFrames groups in value order."""
    arr = values.to_numpy()
    if pd.api.types.is_integer_dtype(values.dtype) and len(arr) > 0:
        lo, hi = int(arr.min()), int(arr.max())
        if hi - lo + 1 <= 2 * len(arr):
            return _sorted_codes(arr)
    return _first_appearance_codes(arr)


def _time_rank(times: np.ndarray) -> np.ndarray:
    """Rank of each time label under ``sort`` (utils.jl:81-90)."""
    codes, _ = pd.factorize(times, sort=True)
    return codes.astype(np.int64)


def _fill_by_group(col: np.ndarray, ncol: int, trank: np.ndarray, counts: np.ndarray, n_time: int, what: str,
                   rep_msg: str = "") -> np.ndarray:
    """R[t, j] = count of the t-th row (by time, stable) of column-group j; checks group sizes."""
    sizes = np.bincount(col, minlength=ncol)
    if np.any(sizes != n_time):
        raise BarBayError(
            f"Not all {what} barcodes have reported counts in all time points{rep_msg}.\n"
            "Please check your data to ensure:\n"
            "    - No missing timepoints for any barcode\n"
            "    - Consistent time series length across barcodes")
    order = np.lexsort((np.arange(col.size), trank, col))       # DF.sort!(d, time_col) within each group
    R = np.empty((n_time, ncol), dtype=np.int64)
    pos_in_group = np.arange(col.size) - np.repeat(np.arange(ncol) * n_time, n_time)
    R[pos_in_group, col[order]] = counts[order]
    return R


def _as_int64_counts(data: pd.DataFrame, count_col: str) -> np.ndarray:
    c = data[count_col].to_numpy()
    if not np.issubdtype(c.dtype, np.integer):
        raise BarBayError(f"Column {count_col} must contain Int64 counts")   # Matrix{Int64} assignment would throw
    return c.astype(np.int64)


def _process_single(data: pd.DataFrame, mask: np.ndarray, id_col, time_col, count_col, n_time, what, rep_msg=""):
    """_process_{neutral,mutant}_barcodes_single (utils.jl:98-169)."""
    sub = data.loc[mask]
    codes, ids = _groupby_codes(sub[id_col])
    R = _fill_by_group(codes, len(ids), _time_rank(sub[time_col].to_numpy()),
                       _as_int64_counts(sub, count_col), n_time, what, rep_msg)
    return R, list(ids)


def _process_multi(data: pd.DataFrame, mask: np.ndarray, id_col, time_col, count_col, rep_col, n_time, sort_keys):
    """_process_{neutral,mutant}_barcodes_multi (utils.jl:174-271): neutrals keep
    ``unique`` order for ids and reps (:198-200), mutants use ``sort(unique(...))`` (:242-244)."""
    sub = data.loc[mask]
    coder = _sorted_codes if sort_keys else _first_appearance_codes
    icodes, ids = coder(sub[id_col].to_numpy())
    rcodes, reps = coder(sub[rep_col].to_numpy())
    ni, nr = len(ids), len(reps)
    col = rcodes * ni + icodes
    sizes = np.bincount(col, minlength=ni * nr)
    if np.any(sizes != n_time):
        # Julia: R[:, j, k] = d[:, count_col] throws DimensionMismatch
        raise BarBayError("DimensionMismatch: a (barcode, replicate) pair does not have one count per time point")
    order = np.lexsort((np.arange(col.size), _time_rank(sub[time_col].to_numpy()), col))
    counts = _as_int64_counts(sub, count_col)
    R = np.empty((n_time, ni, nr), dtype=np.int64)
    pos = np.arange(col.size) - np.repeat(np.arange(ni * nr) * n_time, n_time)
    cs = col[order]
    R[pos, cs % ni, cs // ni] = counts[order]
    return R, list(ids)


def _env_list(df: pd.DataFrame, time_col, env_col) -> list:
    """collect(sort(unique(data[:, [time_col, env_col]]), time_col)[:, env_col]) (utils.jl:576-578)."""
    pairs = df[[time_col, env_col]].drop_duplicates()
    order = np.argsort(_time_rank(pairs[time_col].to_numpy()), kind="stable")
    return pairs[env_col].to_numpy()[order].tolist()


def _unique_list(values) -> list:
    seen, out = set(), []
    for v in values:
        if v not in seen:
            seen.add(v)
            out.append(v)
    return out


def data_to_arrays(data: pd.DataFrame, *, id_col="barcode", time_col="time", count_col="count",
                   neutral_col="neutral", rep_col=None, env_col=None, genotype_col=None) -> DataArrays:
    """Tidy frame -> model inputs (utils.jl:996-1033)."""
    for c in (id_col, time_col, count_col, neutral_col, rep_col, env_col, genotype_col):
        if c is not None and str(c) not in data.columns:
            raise BarBayError(f"Column {c} does not exist in the dataframe")            # :1007-1013
    if data[neutral_col].dtype != np.bool_:
        raise BarBayError(f"Column {neutral_col} must be of type Bool")                 # :1016-1018
    neutral = data[neutral_col].to_numpy()
    n_time_all = len(pd.unique(data[time_col]))                                          # _extract_timepoints

    if rep_col is None:
        Rn, neutral_ids = _process_single(data, neutral, id_col, time_col, count_col, n_time_all, "neutral")
        Rm, bc_ids = _process_single(data, ~neutral, id_col, time_col, count_col, n_time_all, "mutant")
        R = np.concatenate([Rn, Rm], axis=1)                                             # hcat :428
        out = DataArrays(R, R.sum(axis=1), len(neutral_ids), len(bc_ids), bc_ids, neutral_ids,
                         "env1", 1, 1, n_time_all, "N/A", 0)
        groups = None
    else:
        gcodes, gkeys = _groupby_codes(data[rep_col])                                    # DF.groupby(data, rep_col) :489
        groups = [data.loc[gcodes == g] for g in range(len(gkeys))]
        n_rep = len(groups)
        n_rep_time = [len(pd.unique(g[time_col])) for g in groups]
        if len(set(n_rep_time)) == 1:                                                    # :497
            Rn, neutral_ids = _process_multi(data, neutral, id_col, time_col, count_col, rep_col, n_time_all, False)
            Rm, bc_ids = _process_multi(data, ~neutral, id_col, time_col, count_col, rep_col, n_time_all, True)
            if Rn.shape[2] != Rm.shape[2]:
                raise BarBayError("DimensionMismatch: neutral and mutant barcodes span different replicates")
            R = np.concatenate([Rn, Rm], axis=1)                                         # cat(...; dims=2) :505
            nt = R.sum(axis=1)                                                           # T x n_rep :506
        else:
            R, neutral_ids, bc_ids = [], [], []
            for rep, g in enumerate(groups):
                gn = g[neutral_col].to_numpy()
                msg = f" for replicate {rep + 1}"
                Rn, ids_n = _process_single(g, gn, id_col, time_col, count_col, n_rep_time[rep], "neutral", msg)
                Rm, ids_m = _process_single(g, ~gn, id_col, time_col, count_col, n_rep_time[rep], "mutant", msg)
                if rep == 0:
                    neutral_ids, bc_ids = ids_n, ids_m                                   # :295-297, :351-353
                R.append(np.concatenate([Rn, Rm], axis=1))                               # :516
            nt = [r.sum(axis=1) for r in R]                                              # :517
        out = DataArrays(R, nt, len(neutral_ids), len(bc_ids), bc_ids, neutral_ids,
                         "env1", 1, n_rep, n_rep_time, "N/A", 0)

    if env_col is not None:
        if rep_col is None:
            envs = _env_list(data, time_col, env_col)
            n_env = len(_unique_list(envs))
        else:
            envs_r = [_env_list(g, time_col, env_col) for g in groups]                   # :647-650
            n_env = len(_unique_list([e for es in envs_r for e in es]))
            envs = envs_r[0] if all(es == envs_r[0] for es in envs_r) else envs_r        # :656-658
        out.envs, out.n_env = envs, n_env

    if genotype_col is not None:
        geno = dict(zip(data[id_col].tolist(), data[genotype_col].tolist()))            # :705-707
        out.genotypes = [geno[m] for m in out.bc_ids]                                    # :709-713
        out.n_geno = len(_unique_list(out.genotypes))
    return out


# ---------------------------------------------------------------------------
# ADVI output -> tidy frame
# ---------------------------------------------------------------------------
@dataclass
class _Dist:
    m: np.ndarray
    σ: np.ndarray

    @property
    def sigma(self):
        return self.σ


@dataclass
class _Transform:
    ranges_out: list


@dataclass
class MeanFieldPosterior:
    """Stand-in for ``Bijectors.transformed(TuringDiagMvNormal(m, σ), Stacked(...))``: exposes the
    three fields advi_to_df reads (utils.jl:1049, :1060)."""
    dist: _Dist
    transform: _Transform

    @classmethod
    def build(cls, m, sigma, ranges_out):
        return cls(_Dist(np.asarray(m, dtype=np.float64), np.asarray(sigma, dtype=np.float64)),
                   _Transform(list(ranges_out)))


def _decode(codes: np.ndarray, categories) -> "pd.api.extensions.ExtensionArray":
    """categories[codes] as a pandas ``str`` column, decoded by Arrow (an object array of 7 * 10^6 Python strings
    costs pandas 0.8 s to ingest, the dictionary decode 0.1 s); categories may contain None (missing)."""
    import pyarrow as pa
    arr = pa.DictionaryArray.from_arrays(pa.array(np.asarray(codes, dtype=np.int32)),
                                         pa.array(list(categories), type=pa.large_string()))
    return pd.array(arr.cast(pa.large_string()), dtype="str")


def _slice(rng) -> slice:
    return slice(rng.start - 1, rng.stop - 1)      # 1-based UnitRange -> 0-based slice


def _n_time_of(output: DataArrays, r: int) -> int:
    return output.n_time if np.isscalar(output.n_time) else output.n_time[r]


def add_replicate_info(df, var_groups, var_range, output: DataArrays, rep_col) -> None:
    """utils.jl:1100-1161."""
    col = np.empty(len(df), dtype=object)
    reps = range(output.n_rep)
    for name, rng in zip(var_groups, var_range):
        s = _slice(rng)
        if _model.POP_MARK in name:
            if output.n_rep == 1:
                col[s] = "R1"
            else:
                col[s] = np.concatenate([np.repeat(f"R{r + 1}", _n_time_of(output, r) - 1) for r in reps])
        elif name == _model.V_THETA:
            col[s] = "N/A"
        elif name == _model.V_LOGLAM:
            if output.n_rep == 1:
                col[s] = "R1"
            else:
                B = output.n_bc + output.n_neutral
                col[s] = np.concatenate([np.repeat(f"R{r + 1}", B * _n_time_of(output, r)) for r in reps])
        else:
            if output.n_rep == 1:
                col[s] = "R1"
            else:
                col[s] = np.repeat([f"R{r + 1}" for r in reps], output.n_bc * output.n_env)
    df[str(rep_col)] = col


def add_environment_info(df, var_groups, var_range, output: DataArrays, env_col) -> None:
    """utils.jl:1166-1189.  Rows the reference leaves ``#undef`` (barcode-level and
    log-Poisson rows of multi-environment fits) are ``None`` here."""
    col = np.empty(len(df), dtype=object)
    col[:] = None
    for name, rng in zip(var_groups, var_range):
        s = _slice(rng)
        if output.n_env == 1:
            col[s] = "env1"
        elif _model.POP_MARK in name:
            if isinstance(output.envs[0], list):
                # one list per replicate (unequal T): every replicate's envs[2:end], replicate after replicate -- the
                # reference's `output.envs[2:end]` (:1179) does not fit the rows in this case
                tail = [e for es in output.envs for e in es[1:]]
            else:
                tail = list(output.envs[1:])
            n = s.stop - s.start
            col[s] = (tail * (n // max(len(tail), 1)))[:n] if len(tail) != n else tail
        elif name == _model.V_THETA:
            envs = output.envs if not isinstance(output.envs[0], list) else [e for es in output.envs for e in es]
            col[s] = np.tile(np.asarray(_unique_list(envs), dtype=object), output.n_bc)
    df[str(env_col)] = col


def _add_barcode_info_objects(df, var_groups, var_range, output: DataArrays, genotype_col=None) -> None:
    """utils.jl:1194-1279, ids of any type (kept as they are in an object column)."""
    col = np.empty(len(df), dtype=object)
    bc = np.asarray(output.bc_ids, dtype=object)
    for name, rng in zip(var_groups, var_range):
        s = _slice(rng)
        if _model.POP_MARK in name:
            col[s] = "N/A"
        elif name == _model.V_THETA and genotype_col is None:
            col[s] = bc if output.n_env == 1 else np.repeat(bc, output.n_env)
        elif name == _model.V_THETA:
            col[s] = np.asarray(_unique_list(output.genotypes), dtype=object)
        elif name == _model.V_LOGLAM:
            all_ids = np.asarray(list(output.neutral_ids) + list(output.bc_ids), dtype=object)
            if output.n_rep == 1:
                col[s] = np.repeat(all_ids, _n_time_of(output, 0))
            else:
                col[s] = np.concatenate([np.repeat(all_ids, _n_time_of(output, r)) for r in range(output.n_rep)])
        else:
            if output.n_rep == 1 and output.n_env == 1:
                col[s] = bc
            elif output.n_rep == 1:
                col[s] = np.repeat(bc, output.n_env)
            elif output.n_env == 1:
                col[s] = np.tile(bc, output.n_rep)
            else:
                col[s] = np.tile(np.repeat(bc, output.n_env), output.n_rep)
    df["id"] = col


def add_barcode_info(df, var_groups, var_range, output: DataArrays, genotype_col=None) -> None:
    """utils.jl:1194-1279.  The column is assembled as integer codes into one table of ids
    ("N/A" | barcodes | neutrals | genotypes) and decoded once (see _decode)."""
    genos = list(_unique_list(output.genotypes)) if genotype_col is not None else []
    bc_ids, neu_ids = list(output.bc_ids), list(output.neutral_ids)
    if not all(isinstance(x, str) for x in bc_ids + neu_ids + genos):
        return _add_barcode_info_objects(df, var_groups, var_range, output, genotype_col)
    table = ["N/A"] + bc_ids + neu_ids + genos
    o_bc, o_neu, o_gen = 1, 1 + len(bc_ids), 1 + len(bc_ids) + len(neu_ids)
    bc = np.arange(o_bc, o_bc + len(bc_ids), dtype=np.int32)
    all_ids = np.concatenate([np.arange(o_neu, o_neu + len(neu_ids), dtype=np.int32), bc])
    col = np.zeros(len(df), dtype=np.int32)
    for name, rng in zip(var_groups, var_range):
        s = _slice(rng)
        if _model.POP_MARK in name:
            col[s] = 0
        elif name == _model.V_THETA and genotype_col is None:
            col[s] = bc if output.n_env == 1 else np.repeat(bc, output.n_env)
        elif name == _model.V_THETA:
            col[s] = np.arange(o_gen, o_gen + len(genos), dtype=np.int32)
        elif name == _model.V_LOGLAM:
            if output.n_rep == 1:
                col[s] = np.repeat(all_ids, _n_time_of(output, 0))
            else:
                col[s] = np.concatenate([np.repeat(all_ids, _n_time_of(output, r)) for r in range(output.n_rep)])
        else:
            if output.n_rep == 1 and output.n_env == 1:
                col[s] = bc
            elif output.n_rep == 1:
                col[s] = np.repeat(bc, output.n_env)
            elif output.n_env == 1:
                col[s] = np.tile(bc, output.n_rep)
            else:
                col[s] = np.tile(np.repeat(bc, output.n_env), output.n_rep)
    df["id"] = _decode(col, table)


def process_hierarchical_samples(df: pd.DataFrame, output: DataArrays, n_samples: int, genotype_col=None,
                                 seed: int | None = None, chunk: int = 4096, derived=None) -> pd.DataFrame:
    """utils.jl:1284-1343: s = θ + exp(logτ) θ̃ from ``n_samples`` Normal draws per variable;
    the appended ``bc_fitness`` rows carry the *median* in ``mean`` (quirk 5) and the sample std.

    The genotype case indexes θ by genotype (``θ_mat[:, geno_idx]``), the evident intent of
    utils.jl:1310, which in the reference only runs when G == 1 or G == M.
    Columns are processed in chunks so 10^6 barcodes do not need n_samples x M x R doubles at once.
    ``derived`` = (median, sd) per logτ row computed on the device (``Engine.derived_fitness``) replaces the host
    sampling: at 10^6 barcodes the 3 x 10^4 host normals per row cost more than the whole fit.
    """
    rng = np.random.default_rng(seed)
    vt = df["vartype"].to_numpy()
    th = df.loc[vt == "bc_hyperfitness", ["mean", "std"]].to_numpy()
    tau = df.loc[vt == "bc_deviations", ["mean", "std"]].to_numpy()
    tt = df.loc[vt == "bc_noncenter", ["mean", "std"]].to_numpy()
    n_out = tau.shape[0]
    if genotype_col is not None:
        _, gidx = _model.indexin_unique(list(output.genotypes))
        th_index = np.asarray(gidx, dtype=np.int64) - 1                 # column m <- genotype of barcode m
    else:
        th_index = np.tile(np.arange(th.shape[0]), output.n_rep)        # hcat(repeat([θ_mat], n_rep)...)
    # Only per-column marginals (median, std) are reported, so each output column draws its own
    # theta samples; the reference reuses one draw matrix across replicates, which has the same marginals.
    med = np.empty(n_out)
    sd = np.empty(n_out)
    if derived is not None:
        med, sd = np.asarray(derived[0], dtype=np.float64), np.asarray(derived[1], dtype=np.float64)
        if med.shape != (n_out,) or sd.shape != (n_out,):
            raise ValueError("derived rows do not match the logτ rows")
    for a in range(0, n_out if derived is None else 0, chunk):
        b = min(n_out, a + chunk)
        idx = th_index[a:b]
        th_s = rng.normal(th[idx, 0], th[idx, 1], size=(n_samples, b - a))
        tau_s = np.exp(rng.normal(tau[a:b, 0], tau[a:b, 1], size=(n_samples, b - a)))
        tt_s = rng.normal(tt[a:b, 0], tt[a:b, 1], size=(n_samples, b - a))
        s = th_s + tau_s * tt_s
        med[a:b] = np.median(s, axis=0)
        sd[a:b] = np.std(s, axis=0, ddof=1)
    tau_rows = df["varname"].str.contains("τ", regex=False).to_numpy()
    new = pd.DataFrame({"mean": med, "std": sd})
    new["varname"] = df.loc[tau_rows, "varname"].str.replace("logτ", "s", regex=False).to_numpy()
    new["vartype"] = "bc_fitness"
    for c in ("rep", "env"):                                             # :1330-1337 literal column names
        if c in df.columns:
            new[c] = df.loc[tau_rows, c].to_numpy()
    new["id"] = df.loc[tau_rows, "id"].to_numpy()
    return pd.concat([df, new], ignore_index=True)


def advi_to_df(data: pd.DataFrame, dist: MeanFieldPosterior, vars: list, *, id_col="barcode", time_col="time",
               count_col="count", neutral_col="neutral", rep_col=None, env_col=None, genotype_col=None,
               n_samples: int = 10_000, seed: int | None = None, output: DataArrays | None = None,
               derived=None) -> pd.DataFrame:
    """utils.jl:1409-1462.  ``output`` lets a caller that already packed ``data`` skip the second
    data_to_arrays call the reference makes (:1423-1432)."""
    if output is None:
        output = data_to_arrays(data, id_col=id_col, time_col=time_col, count_col=count_col,
                                neutral_col=neutral_col, rep_col=rep_col, env_col=env_col,
                                genotype_col=genotype_col)
    var_range = dist.transform.ranges_out                                     # :1049
    df = pd.DataFrame({"mean": dist.dist.m, "std": dist.dist.σ})              # :1060
    if isinstance(vars, _model.VarNames):          # the names the engine's layout hands out: no Python string per row
        var_groups = vars.groups
        df["varname"] = vars.to_pandas()
    else:
        vars = list(vars)
        var_groups = [v.replace("[1]", "") for v in vars if "[1]" in str(v)]  # :1046
        df["varname"] = vars
    types = ["tmp"] + sorted(set(_model.VARNAME_TO_VARTYPE.values()))
    code = np.zeros(len(df), dtype=np.int32)
    for name, rng in zip(var_groups, var_range):                              # :1083-1093
        code[_slice(rng)] = types.index(_model.VARNAME_TO_VARTYPE[name])
    df["vartype"] = _decode(code, types)
    if rep_col is not None:
        add_replicate_info(df, var_groups, var_range, output, rep_col)
    if env_col is not None:
        add_environment_info(df, var_groups, var_range, output, env_col)
    add_barcode_info(df, var_groups, var_range, output, genotype_col)
    if len(var_groups) == 7 and (output.n_rep > 1 or genotype_col is not None):   # :1457
        df = process_hierarchical_samples(df, output, n_samples, genotype_col, seed, derived=derived)
    return df
