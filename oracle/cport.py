"""ctypes face of oracle/c/libadvi_port.so (CPU port of the fitness_normal ADVI step).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "libadvi_port.so")


class _Problem(C.Structure):
    _fields_ = [("T", C.c_int), ("N", C.c_int), ("M", C.c_int), ("R", C.POINTER(C.c_int64)),
                ("pri", (C.c_double * 2) * 5)]


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "advi_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        D = C.POINTER(C.c_double)
        _lib.port_n_latent.restype = C.c_long
        _lib.port_n_latent.argtypes = [C.POINTER(_Problem)]
        _lib.port_num_threads.restype = C.c_int
        _lib.port_elbo_grad.restype = C.c_double
        _lib.port_elbo_grad.argtypes = [C.POINTER(_Problem), D, D, D, C.c_int, D, D]
        _lib.port_advi_steps.restype = C.c_double
        _lib.port_advi_steps.argtypes = [C.POINTER(_Problem), D, D, D, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_long]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PortProblem:
    """fitness_normal with vector priors (the BASELINE cfg2 workload)."""

    def __init__(self, bc_count, n_neutral, n_bc, priors=None):
        from .model_ref import DEFAULT_PRIORS
        pr = dict(DEFAULT_PRIORS)
        pr.update(priors or {})
        R = np.asarray(bc_count, dtype=np.int64)
        self.T = R.shape[0]
        self._flat = np.ascontiguousarray(R.T.reshape(-1))      # Julia column-major T x B
        self.p = _Problem()
        self.p.T, self.p.N, self.p.M = self.T, int(n_neutral), int(n_bc)
        self.p.R = self._flat.ctypes.data_as(C.POINTER(C.c_int64))
        for i, key in enumerate(["s_pop_prior", "logσ_pop_prior", "s_bc_prior", "logσ_bc_prior", "logλ_prior"]):
            v = np.asarray(pr[key], dtype=np.float64)
            if v.ndim != 1:
                raise ValueError("the C port takes vector priors only")
            self.p.pri[i][0], self.p.pri[i][1] = float(v[0]), float(v[1])
        self.D = int(lib().port_n_latent(C.byref(self.p)))

    @property
    def threads(self) -> int:
        return int(lib().port_num_threads())

    def elbo_grad(self, mu, omega, eps):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        eps = np.ascontiguousarray(eps, dtype=np.float64).reshape(-1, self.D)
        K = eps.shape[0]
        grad, logp = np.empty(2 * self.D), np.empty(K)
        elbo = lib().port_elbo_grad(C.byref(self.p), _ptr(mu), _ptr(omega), _ptr(eps), K, _ptr(grad), _ptr(logp))
        return float(elbo), grad[:self.D], grad[self.D:], logp

    def advi_steps(self, theta, acc, n_steps, K, kind="decayed", eta=0.1, tau_or_pre=1.0, post=0.9, n_ring=100,
                   ring=None, seed=0, first_step=0):
        k = 1 if kind == "decayed" else 0
        if k == 0 and ring is None:
            raise ValueError("TruncatedADAGrad needs a ring buffer")
        rp = _ptr(ring) if ring is not None else None
        return float(lib().port_advi_steps(C.byref(self.p), _ptr(theta), _ptr(acc), rp, k, eta, tau_or_pre, post,
                                           n_ring, K, n_steps, seed, first_step))
