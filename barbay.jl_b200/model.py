"""Model descriptors: the host-side mirror of ``BarBay.model``.

The reference's models are Turing ``@model`` functions (src/model.jl:1-28 and the
five src/model_*.jl files); ``BarBay.vi.advi`` dispatches on substrings of the
function *name* (src/vi.jl:111-169).  Here a model is a small descriptor object
with the same name whose call collects keyword arguments; the log-joint itself
lives in the CUDA kernels (csrc/bb_kernels.cuh).  The latent layout reproduced by
``var_groups`` is the VarInfo order of the reference model bodies.
"""
from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass, field
from typing import Any

import numpy as np

# variable-group names exactly as the Turing models spell them (utils.jl:1069-1078)
V_S_POP = "s̲ₜ"                               # s̲ₜ
V_LOGSIG_POP = "logσ̲ₜ"                  # logσ̲ₜ
V_S_BC = "s̲⁽ᵐ⁾"                    # s̲⁽ᵐ⁾
V_LOGSIG_BC = "logσ̲⁽ᵐ⁾"       # logσ̲⁽ᵐ⁾
V_THETA = "θ̲⁽ᵐ⁾"              # θ̲⁽ᵐ⁾
V_THETA_TILDE = "θ̲̃⁽ᵐ⁾"  # θ̲̃⁽ᵐ⁾
V_LOGTAU = "logτ̲⁽ᵐ⁾"          # logτ̲⁽ᵐ⁾
V_LOGLAM = "logΛ̲̲"                      # logΛ̲̲
POP_MARK = "̲ₜ"                               # "̲ₜ" (utils.jl:1118)

VARNAME_TO_VARTYPE = {                                  # utils.jl:1069-1078
    V_S_POP: "pop_mean_fitness",
    V_LOGSIG_POP: "pop_std",
    V_S_BC: "bc_fitness",
    V_LOGSIG_BC: "bc_std",
    V_THETA: "bc_hyperfitness",
    V_THETA_TILDE: "bc_noncenter",
    V_LOGTAU: "bc_deviations",
    V_LOGLAM: "log_poisson",
}

DEFAULT_PRIORS = {                                      # model_fitness_normal.jl:125-129, replicates.jl:155
    "s_pop_prior": [0.0, 2.0],
    "logσ_pop_prior": [0.0, 1.0],
    "s_bc_prior": [0.0, 2.0],
    "logσ_bc_prior": [0.0, 1.0],
    "logλ_prior": [3.0, 3.0],
    "logτ_prior": [-2.0, 1.0],
}
# ASCII aliases accepted for the Unicode keyword names
PRIOR_ALIASES = {
    "logsig_pop_prior": "logσ_pop_prior", "logsigma_pop_prior": "logσ_pop_prior",
    "logsig_bc_prior": "logσ_bc_prior", "logsigma_bc_prior": "logσ_bc_prior",
    "loglam_prior": "logλ_prior", "loglambda_prior": "logλ_prior",
    "logtau_prior": "logτ_prior",
}


@dataclass(frozen=True)
class Model:
    """One of the reference's model functions, identified by its name."""
    name: str
    hier: bool
    multienv: bool
    replicate: bool
    genotype: bool
    allowed_kwargs: tuple = ()

    def __str__(self) -> str:            # "$(model)" in src/vi.jl:111
        return self.name

    @property
    def __name__(self) -> str:           # noqa: A003 - mirrors a function object
        return self.name


_COMMON = ("s_pop_prior", "logσ_pop_prior", "s_bc_prior", "logσ_bc_prior", "logλ_prior")
fitness_normal = Model("fitness_normal", False, False, False, False, _COMMON)
replicate_fitness_normal = Model("replicate_fitness_normal", True, False, True, False, _COMMON + ("logτ_prior",))
multienv_fitness_normal = Model("multienv_fitness_normal", False, True, False, False, _COMMON + ("envs",))
genotype_fitness_normal = Model("genotype_fitness_normal", True, False, False, True,
                                _COMMON + ("logτ_prior", "genotypes"))
multienv_replicate_fitness_normal = Model("multienv_replicate_fitness_normal", True, True, True, False,
                                          _COMMON + ("logτ_prior", "envs"))
MODELS = {m.name: m for m in (fitness_normal, replicate_fitness_normal, multienv_fitness_normal,
                              genotype_fitness_normal, multienv_replicate_fitness_normal)}


def resolve(model) -> Model:
    if isinstance(model, Model):
        return model
    if isinstance(model, str) and model in MODELS:
        return MODELS[model]
    name = getattr(model, "__name__", None)
    if name in MODELS:
        return MODELS[name]
    raise TypeError(f"unknown model {model!r}")


def indexin_unique(labels) -> tuple[list, np.ndarray]:
    """``unique(x)`` (first appearance) and 1-based ``indexin(x, unique(x))``
    (multienv.jl:151-155, genotypes.jl:170-174)."""
    uniq: list = []
    pos: dict = {}
    idx = np.empty(len(labels), dtype=np.int32)
    for i, lab in enumerate(labels):
        if lab not in pos:
            pos[lab] = len(uniq)
            uniq.append(lab)
        idx[i] = pos[lab] + 1
    return uniq, idx


def normalise_kwargs(model: Model, model_kwargs: dict | None) -> dict:
    """Keyword arguments of the model call, Unicode names, defaults filled in."""
    out: dict[str, Any] = {}
    for k, v in (model_kwargs or {}).items():
        k = str(k)
        k = PRIOR_ALIASES.get(k, k)
        if k not in model.allowed_kwargs:
            # Julia: MethodError "got unsupported keyword argument"
            raise TypeError(f"{model.name}: got unsupported keyword argument \"{k}\"")
        out[k] = v
    for k in model.allowed_kwargs:
        if k in DEFAULT_PRIORS and k not in out:
            out[k] = DEFAULT_PRIORS[k]
    return out


@dataclass
class VarGroup:
    name: str
    length: int
    start: int = 0        # 0-based offset into the latent vector

    @property
    def range(self) -> range:   # 1-based inclusive UnitRange like q.transform.ranges_out
        return range(self.start + 1, self.start + self.length + 1)


@dataclass
class ModelLayout:
    model: Model
    n_rep: int
    n_time: list
    n_neutral: int
    n_bc: int
    n_env: int
    n_geno: int
    groups: list = field(default_factory=list)

    @property
    def n_latent(self) -> int:
        return sum(g.length for g in self.groups)

    @property
    def var_names(self) -> "VarNames":
        """["<group>[i]" ...] exactly as src/vi.jl:184-198 builds them (a lazy sequence: see VarNames)."""
        return VarNames([(g.name, g.length) for g in self.groups])

    @property
    def ranges_out(self) -> list:
        return [g.range for g in self.groups]


class VarNames(Sequence):
    """The variable names ``"<group>[i]"`` of src/vi.jl:184-198 as a read-only sequence.

    A 10^6-barcode fit has 7 * 10^6 of them; as Python strings they cost more than the fit (1.5 s to build, 1 s to
    scan for the groups, 0.5 s to hand to pandas).  The sequence behaves like the list the reference builds (len,
    indexing, slicing, iteration, ``==`` with a list), knows its groups, and turns into the DataFrame column in one
    vectorised Arrow pass (``to_pandas``)."""

    def __init__(self, groups):
        self._groups = [(str(n), int(k)) for n, k in groups]
        self._start = np.concatenate([[0], np.cumsum([k for _, k in self._groups])]).astype(np.int64)

    @property
    def groups(self) -> list:
        return [n for n, _ in self._groups]

    def __len__(self) -> int:
        return int(self._start[-1])

    def _one(self, i: int) -> str:
        g = int(np.searchsorted(self._start, i, side="right")) - 1
        return f"{self._groups[g][0]}[{i - int(self._start[g]) + 1}]"

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._one(i)

    def __iter__(self):
        for name, k in self._groups:
            for i in range(1, k + 1):
                yield f"{name}[{i}]"

    def __eq__(self, other):
        if isinstance(other, VarNames):
            return self._groups == other._groups
        try:
            return len(other) == len(self) and all(a == b for a, b in zip(self, other))
        except TypeError:
            return NotImplemented

    __hash__ = None

    def to_pandas(self):
        """The ``varname`` column (pandas ``str`` dtype) without a Python string per row."""
        import pandas as pd
        import pyarrow as pa
        import pyarrow.compute as pc
        parts = []
        for name, k in self._groups:
            if k == 0:
                continue
            idx = pc.cast(pa.array(np.arange(1, k + 1, dtype=np.int64)), pa.large_string())
            parts.append(pc.binary_join_element_wise(pa.scalar(f"{name}[", pa.large_string()), idx,
                                                     pa.scalar("]", pa.large_string()), pa.scalar("", pa.large_string())))
        arr = pa.chunked_array(parts, type=pa.large_string()) if parts else pa.chunked_array([], type=pa.large_string())
        return pd.array(arr, dtype="str")


def var_groups(model: Model, n_time, n_rep: int, n_neutral: int, n_bc: int, n_env: int = 1,
               n_geno: int = 0) -> ModelLayout:
    """Latent groups in VarInfo order."""
    nts = [int(n_time)] * n_rep if np.isscalar(n_time) else [int(t) for t in n_time]
    B = n_neutral + n_bc
    n_st = sum(t - 1 for t in nts)
    n_lam = sum(t * B for t in nts)
    E = n_env if model.multienv else 1
    groups = [VarGroup(V_S_POP, n_st), VarGroup(V_LOGSIG_POP, n_st)]
    if model.hier:
        n_hyper = n_geno if model.genotype else E * n_bc
        per = E * n_bc * n_rep
        groups += [VarGroup(V_THETA, n_hyper), VarGroup(V_THETA_TILDE, per), VarGroup(V_LOGTAU, per),
                   VarGroup(V_LOGSIG_BC, per)]
    else:
        groups += [VarGroup(V_S_BC, E * n_bc), VarGroup(V_LOGSIG_BC, E * n_bc)]
    groups.append(VarGroup(V_LOGLAM, n_lam))
    off = 0
    for g in groups:
        g.start = off
        off += g.length
    return ModelLayout(model, n_rep, nts, n_neutral, n_bc, E, n_geno, groups)
