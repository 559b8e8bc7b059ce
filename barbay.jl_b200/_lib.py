"""ctypes binding of libbarbay_b200.so (the C ABI of include/barbay_b200.h).

The library is the product path: there is no CPU fallback.  If the shared object
is missing, or no B200 is visible when a handle is created, this fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 2
_HERE = os.path.dirname(os.path.abspath(__file__))
# BB_LIB_PATH: tuning builds of the same library (development only)
LIB_PATH = os.environ.get("BB_LIB_PATH") or os.path.join(_HERE, "libbarbay_b200.so")

MODEL_IDS = {
    "fitness_normal": 0,
    "replicate_fitness_normal": 1,
    "multienv_fitness_normal": 2,
    "genotype_fitness_normal": 3,
    "multienv_replicate_fitness_normal": 4,
}
DTYPE_IDS = {"f32": 0, "fp32": 0, "float32": 0, "f64": 1, "fp64": 1, "float64": 1}
OPT_TRUNCATED, OPT_DECAYED = 0, 1


class BarBayError(RuntimeError):
    """Counterpart of the ErrorException the reference raises with ``error(msg)``."""


class bb_prior(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_double)), ("n", C.c_int64), ("is_matrix", C.c_int32)]


class bb_opt(C.Structure):
    _fields_ = [("kind", C.c_int32), ("eta", C.c_double), ("tau", C.c_double), ("post", C.c_double),
                ("n", C.c_int32)]


class bb_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("dtype", C.c_int32), ("n_rep", C.c_int32),
        ("n_time", C.POINTER(C.c_int32)), ("n_neutral", C.c_int32), ("n_bc", C.c_int32),
        ("bc_count", C.POINTER(C.c_int64)), ("n_env", C.c_int32), ("env_idx", C.POINTER(C.c_int32)),
        ("n_geno", C.c_int32), ("geno_idx", C.POINTER(C.c_int32)),
        ("s_pop_prior", bb_prior), ("logsig_pop_prior", bb_prior), ("s_bc_prior", bb_prior),
        ("logsig_bc_prior", bb_prior), ("loglam_prior", bb_prior), ("logtau_prior", bb_prior),
        ("ragged_as_written", C.c_int32), ("n_samples", C.c_int32), ("seed", C.c_uint64),
        ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("n_devices", C.c_int32),
        ("env_per_rep", C.c_int32),
    ]


# every symbol include/barbay_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
SYMBOLS = [
    ("bb_abi_version", C.c_int32, []),
    ("bb_create", C.c_int, [C.POINTER(bb_desc), C.POINTER(_P)]),
    ("bb_layout_probe", C.c_int, [C.POINTER(bb_desc), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    ("bb_destroy", None, [_P]),
    ("bb_last_error", C.c_char_p, [_P]),
    ("bb_n_latent", C.c_int64, [_P]),
    ("bb_init_params", C.c_int, [_P, C.c_uint64]),
    ("bb_set_params", C.c_int, [_P, _D, _D]),
    ("bb_get_params", C.c_int, [_P, _D, _D]),
    ("bb_get_posterior", C.c_int, [_P, _D, _D]),
    ("bb_logjoint_grad", C.c_int, [_P, _D, C.c_int32, C.c_int32, _D, _D]),
    ("bb_elbo_grad", C.c_int, [_P, _D, C.c_int64, _D, _D]),
    ("bb_get_noise", C.c_int, [_P, C.c_int64, _D]),
    ("bb_set_optimizer", C.c_int, [_P, C.POINTER(bb_opt)]),
    ("bb_step", C.c_int, [_P, C.c_int32, _D]),
    ("bb_step_with_noise", C.c_int, [_P, _D]),
    ("bb_step_count", C.c_int64, [_P]),
    ("bb_step_until", C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32), _D, C.POINTER(C.c_int32)]),
    ("bb_state_size", C.c_int64, [_P]),
    ("bb_get_state", C.c_int, [_P, _D]),
    ("bb_set_state", C.c_int, [_P, _D]),
    ("bb_set_stream", C.c_int, [_P, _P]),
    ("bb_sync", C.c_int, [_P]),
    ("bb_launch_count", C.c_int64, [_P]),
    ("bb_algorithmic_bytes_per_step", C.c_double, [_P]),
    ("bb_time_steps", C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    ("bb_persist_stats", C.c_int, [_P, _D]),
    ("bb_data_plane", C.c_int, [_P, C.POINTER(C.c_int32)]),
    ("bb_derived_fitness", C.c_int, [_P, C.c_int32, C.c_uint64, _D, _D]),
    ("bb_n_derived", C.c_int64, [_P]),
    ("bb_naive_prior", C.c_int, [C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                 C.c_int32, _D, _D, _D]),
    ("bb_peer_handle", C.c_int, [_P, C.c_char * 64]),
    ("bb_peer_attach", C.c_int, [_P, C.c_char_p, C.c_int32]),
    ("bb_comm_unique_id", C.c_int, [C.c_char * 128]),
    ("bb_comm_init", C.c_int, [_P, C.c_char * 128]),
]

_lib = None


def load() -> C.CDLL:
    """Load the shared library and type every entry point; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BarBayError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(barbay_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bb_abi_version() != ABI_VERSION:
        raise BarBayError("libbarbay_b200.so ABI version mismatch")
    _lib = lib
    return lib
