"""Synthetic barcode counts drawn from the models' own generative process
(SURVEY.md §8d "Synthetic inputs").  Used by bench.py and the
large-size property tests; the reference ships no generator (its four CSV
fixtures under test/data come from an external simulator)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd

from .utils import DataArrays

BASE_SEED = 20261018


@dataclass
class SynthTruth:
    s: np.ndarray            # fitness per (env, mutant, replicate) as used by the simulation
    theta: np.ndarray | None
    s_pop: np.ndarray        # population mean fitness per (time-1, replicate)


def _simulate_block(rng, s_cols: np.ndarray, n_time: int, mean_reads: float, noise_sd: float = 0.1):
    """One replicate.  s_cols: (T-1, B) fitness felt by column b on the step t -> t+1."""
    B = s_cols.shape[1]
    logf = rng.standard_normal(B)                      # f_0 proportional to LogNormal(0, 1)
    logf -= np.log(np.exp(logf).sum())
    counts = np.empty((n_time, B), dtype=np.int64)
    s_pop = np.empty(n_time - 1)
    for t in range(n_time):
        f = np.exp(logf)
        counts[t] = rng.poisson(mean_reads * B * f)
        if t == n_time - 1:
            break
        s_pop[t] = float((f * s_cols[t]).sum())
        logf = logf + s_cols[t] - s_pop[t] + noise_sd * rng.standard_normal(B)
        logf -= np.log(np.exp(logf).sum())
    return counts, s_pop


def simulate(model: str, n_neutral: int, n_bc: int, n_time: int, *, n_rep: int = 1, envs=None, n_geno: int = 0,
             seed: int = BASE_SEED, mean_reads: float = 100.0) -> tuple[DataArrays, SynthTruth]:
    """``n_time`` may be a list (one T per replicate: the Vector{Matrix{Int64}} data layout); ``envs`` then is one
    list per replicate for the multi-environment x replicate model."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B = n_neutral + n_bc
    multienv = "multienv" in model
    ragged = isinstance(n_time, (list, tuple))
    n_times = [int(t) for t in n_time] if ragged else [int(n_time)] * n_rep
    if ragged:
        n_rep = len(n_times)
    env_idx_r = [np.zeros(t, dtype=np.int64) for t in n_times]
    n_env = 1
    env_list = "env1"
    if multienv:
        per_rep = [list(es) for es in envs] if isinstance(envs[0], (list, tuple)) else [list(envs)] * n_rep
        uniq = []
        for es in per_rep:
            for e in es:
                if e not in uniq:
                    uniq.append(e)
        env_idx_r = [np.asarray([uniq.index(e) for e in es]) for es in per_rep]
        n_env = len(uniq)
        env_list = per_rep if isinstance(envs[0], (list, tuple)) else list(envs)
    theta = None
    genotypes = "N/A"
    if "genotype" in model:
        g = np.arange(n_bc) % n_geno
        g = g[np.random.Generator(np.random.PCG64(seed + 7919)).permutation(n_bc)]   # fixed shuffle
        theta = 0.25 * rng.standard_normal(n_geno)
        base = theta[g] + 0.05 * rng.standard_normal(n_bc)
        # label genotypes so that first-appearance order is well defined
        genotypes = [f"genotype{int(i):06d}" for i in g]
    elif "replicate" in model:
        theta = 0.25 * rng.standard_normal(n_bc)
        base = theta
    else:
        base = 0.25 * rng.standard_normal(n_bc)
    s = np.empty((n_env, n_bc, n_rep))
    for r in range(n_rep):
        for e in range(n_env):
            dev = 0.0
            if "replicate" in model:
                dev = dev + 0.05 * rng.standard_normal(n_bc)
            if multienv:
                dev = dev + 0.15 * rng.standard_normal(n_bc)
            s[e, :, r] = base + dev
    blocks, s_pops = [], []
    for r in range(n_rep):
        s_cols = np.zeros((n_times[r] - 1, B))
        for t in range(n_times[r] - 1):
            s_cols[t, n_neutral:] = s[env_idx_r[r][t + 1], :, r]
        c, sp = _simulate_block(rng, s_cols, n_times[r], mean_reads)
        blocks.append(c)
        s_pops.append(sp)
    if ragged:
        bc_count = blocks
        bc_total = [c.sum(axis=1) for c in blocks]
        nt_field = n_times
    elif n_rep == 1 and "replicate" not in model:
        bc_count = blocks[0]
        bc_total = bc_count.sum(axis=1)
        nt_field = n_time
    else:
        bc_count = np.stack(blocks, axis=2)
        bc_total = bc_count.sum(axis=1)
        nt_field = [n_time] * n_rep
    width = len(str(max(n_bc, n_neutral)))
    da = DataArrays(
        bc_count=bc_count, bc_total=bc_total, n_neutral=n_neutral, n_bc=n_bc,
        bc_ids=[f"mut{i + 1:0{width}d}" for i in range(n_bc)],
        neutral_ids=[f"neutral{i + 1:0{width}d}" for i in range(n_neutral)],
        envs=env_list, n_env=n_env, n_rep=n_rep, n_time=nt_field, genotypes=genotypes,
        n_geno=n_geno if "genotype" in model else 0)
    return da, SynthTruth(s=s, theta=theta, s_pop=s_pops if ragged else np.stack(s_pops, axis=1))


# the BASELINE.json configurations (index = position in BASELINE.json "configs")
CONFIGS = {
    2: dict(model="fitness_normal", n_neutral=1_000, n_bc=999_000, n_time=5),
    3: dict(model="replicate_fitness_normal", n_neutral=200, n_bc=199_800, n_time=5, n_rep=3),
    4: dict(model="multienv_fitness_normal", n_neutral=500, n_bc=499_500, n_time=8,
            envs=[1, 1, 2, 3, 4, 2, 3, 4]),
    5: dict(model="genotype_fitness_normal", n_neutral=1_000, n_bc=999_000, n_time=5, n_geno=10_000),
}


def config(cfg: int, scale: float = 1.0, seed: int | None = None) -> tuple[str, DataArrays, SynthTruth]:
    """BASELINE config ``cfg`` (2-5); ``scale`` < 1 shrinks the barcode axis for tests."""
    spec = dict(CONFIGS[cfg])
    model = spec.pop("model")
    if scale != 1.0:
        spec["n_neutral"] = max(4, int(round(spec["n_neutral"] * scale)))
        spec["n_bc"] = max(8, int(round(spec["n_bc"] * scale)))
        if "n_geno" in spec:
            spec["n_geno"] = max(2, min(spec["n_bc"], int(round(spec["n_geno"] * scale))))
    da, truth = simulate(model, seed=BASE_SEED + cfg if seed is None else seed, **spec)
    return model, da, truth


def to_tidy_fast(da: DataArrays) -> pd.DataFrame:
    """Vectorised ``to_tidy`` for the BASELINE sizes (5 * 10^6 rows in a second or two): same columns and row order
    (replicate-major, then barcode, then time)."""
    ids = np.asarray(list(da.neutral_ids) + list(da.bc_ids), dtype=object)
    R = np.asarray(da.bc_count)
    n_rep = R.shape[2] if R.ndim == 3 else 1
    T, B = R.shape[0], R.shape[1]
    cols = {
        "time": np.tile(np.arange(1, T + 1), B * n_rep),
        "barcode": np.tile(np.repeat(ids, T), n_rep),
        "count": (R.transpose(2, 1, 0) if R.ndim == 3 else R.T).reshape(-1),
        "neutral": np.tile(np.repeat(np.arange(B) < da.n_neutral, T), n_rep),
    }
    if R.ndim == 3:
        cols["rep"] = np.repeat(np.asarray([f"R{r + 1}" for r in range(n_rep)], dtype=object), T * B)
    if not isinstance(da.envs, str):
        cols["env"] = np.tile(np.asarray(list(da.envs), dtype=object), B * n_rep)
    if not isinstance(da.genotypes, str):
        g = np.asarray(["genotype_neutral"] * da.n_neutral + list(da.genotypes), dtype=object)
        cols["genotype"] = np.tile(np.repeat(g, T), n_rep)
    return pd.DataFrame(cols)


def to_tidy(da: DataArrays) -> pd.DataFrame:
    """DataArrays -> tidy frame with the reference's default column names (small cases only)."""
    rows = []
    ids = list(da.neutral_ids) + list(da.bc_ids)
    R = np.asarray(da.bc_count)
    n_rep = R.shape[2] if R.ndim == 3 else 1
    for r in range(n_rep):
        block = R[:, :, r] if R.ndim == 3 else R
        for b, name in enumerate(ids):
            for t in range(block.shape[0]):
                row = {"time": t + 1, "barcode": name, "count": int(block[t, b]), "neutral": b < da.n_neutral}
                if R.ndim == 3:
                    row["rep"] = f"R{r + 1}"
                if not isinstance(da.envs, str):
                    row["env"] = da.envs[t]
                if not isinstance(da.genotypes, str):
                    row["genotype"] = "genotype_neutral" if b < da.n_neutral else da.genotypes[b - da.n_neutral]
                rows.append(row)
    return pd.DataFrame(rows)
