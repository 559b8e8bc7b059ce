"""barbay.jl_b200 -- B200-native ADVI backend for BarBay.jl's hot path.

Host-side mirror of the reference's public surface for this path:
``vi.advi`` (src/vi.jl:86-235), ``utils.data_to_arrays`` / ``utils.advi_to_df``
(src/utils.jl:996-1033, 1409-1462) and the ``model`` descriptors (src/model.jl).
All inference arithmetic runs in ``libbarbay_b200.so`` (hand-written sm_100a
kernels behind the C ABI of include/barbay_b200.h); there is no CPU fallback.

The directory name carries a dot, so import it through the ``barbay_b200`` shim
at the repository root (``import barbay_b200 as bb``).
"""
from . import _lib, model, utils, vi, engine, synth, stats   # noqa: F401
from ._lib import BarBayError, LIB_PATH, load as load_library   # noqa: F401
from .engine import Engine, comm_unique_id   # noqa: F401
from .vi import ADVI, DecayedADAGrad, TruncatedADAGrad, advi   # noqa: F401

__all__ = ["model", "utils", "vi", "engine", "synth", "stats", "Engine", "advi", "ADVI", "TruncatedADAGrad",
           "DecayedADAGrad", "BarBayError", "load_library", "LIB_PATH", "comm_unique_id"]
