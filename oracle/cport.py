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


_SO_MODELS = os.path.join(_DIR, "libadvi_port_models.so")


def build(force: bool = False) -> str:
    stale = force
    for so, src in ((_SO, "advi_port.c"), (_SO_MODELS, "advi_port_models.c")):
        src = os.path.join(_DIR, src)
        stale = stale or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    if stale:
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
        _lib.port_set_threads.restype = None
        _lib.port_set_threads.argtypes = [C.c_int]
        _lib.port_elbo_grad.restype = C.c_double
        _lib.port_elbo_grad.argtypes = [C.POINTER(_Problem), D, D, D, C.c_int, D, D]
        _lib.port_advi_steps.restype = C.c_double
        _lib.port_advi_steps.argtypes = [C.POINTER(_Problem), D, D, D, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_long]
    return _lib


def set_threads(n: int | None = None) -> int:
    """Use `n` OpenMP threads (default: every host core), whatever OMP_NUM_THREADS says."""
    n = int(n or os.cpu_count() or 1)
    lib().port_set_threads(n)
    mlib().gport_set_threads(n)
    return int(lib().port_num_threads())


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


# ---------------------------------------------------------------------------------------- all model families
class _GProblem(C.Structure):
    _fields_ = [("T", C.c_int), ("N", C.c_int), ("M", C.c_int), ("R", C.c_int), ("E", C.c_int), ("G", C.c_int),
                ("hier", C.c_int), ("env_of_t", C.POINTER(C.c_int32)), ("geno", C.POINTER(C.c_int32)),
                ("counts", C.POINTER(C.c_int64)), ("pri", (C.c_double * 2) * 6)]


_mlib = None


def mlib():
    global _mlib
    if _mlib is None:
        if not os.path.exists(_SO_MODELS):
            build()
        _mlib = C.CDLL(_SO_MODELS)
        D = C.POINTER(C.c_double)
        G = C.POINTER(_GProblem)
        _mlib.gport_n_latent.restype = C.c_long
        _mlib.gport_n_latent.argtypes = [G]
        _mlib.gport_num_threads.restype = C.c_int
        _mlib.gport_set_threads.restype = None
        _mlib.gport_set_threads.argtypes = [C.c_int]
        _mlib.gport_logjoint_grad.restype = None
        _mlib.gport_logjoint_grad.argtypes = [G, D, C.c_int, D, D]
        _mlib.gport_elbo_grad.restype = C.c_double
        _mlib.gport_elbo_grad.argtypes = [G, D, D, D, C.c_int, D, D]
        _mlib.gport_advi_steps.restype = C.c_double
        _mlib.gport_advi_steps.argtypes = [G, D, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_uint64,
                                           C.c_long]
    return _mlib


def _first_appearance_index(values):
    seen, idx = {}, []
    for v in values:
        if v not in seen:
            seen[v] = len(seen)
        idx.append(seen[v])
    return np.asarray(idx, dtype=np.int32), len(seen)


class ModelPort:
    """Any of the five model families (equal time points per replicate, vector priors): log-joint + gradient, ELBO
    gradient and a timed ADVI loop.  Latent order = the reference's VarInfo order (oracle/model_ref.py)."""

    def __init__(self, model: str, bc_count, n_neutral: int, n_bc: int, envs=None, genotypes=None, priors=None):
        from .model_ref import DEFAULT_PRIORS
        pr = dict(DEFAULT_PRIORS)
        pr.update(priors or {})
        R = np.asarray(bc_count, dtype=np.int64)
        self.T = R.shape[0]
        self.nrep = R.shape[2] if R.ndim == 3 else 1
        self._flat = np.ascontiguousarray(R.transpose(2, 1, 0).reshape(-1) if R.ndim == 3 else R.T.reshape(-1))
        p = self.p = _GProblem()
        p.T, p.N, p.M, p.R = self.T, int(n_neutral), int(n_bc), self.nrep
        p.hier = 1 if ("replicate" in model or "genotype" in model) else 0
        p.E, p.G = 1, 0
        self._keep = []
        if "multienv" in model:
            idx, p.E = _first_appearance_index(list(envs))
            self._keep.append(idx)
            p.env_of_t = idx.ctypes.data_as(C.POINTER(C.c_int32))
        if "genotype" in model:
            idx, p.G = _first_appearance_index(list(genotypes))
            self._keep.append(idx)
            p.geno = idx.ctypes.data_as(C.POINTER(C.c_int32))
        p.counts = self._flat.ctypes.data_as(C.POINTER(C.c_int64))
        for i, key in enumerate(["s_pop_prior", "logσ_pop_prior", "s_bc_prior", "logσ_bc_prior", "logλ_prior",
                                 "logτ_prior"]):
            v = np.asarray(pr[key], dtype=np.float64)
            if v.ndim != 1:
                raise ValueError("the C port takes vector priors only")
            p.pri[i][0], p.pri[i][1] = float(v[0]), float(v[1])
        self.D = int(mlib().gport_n_latent(C.byref(p)))

    @property
    def threads(self) -> int:
        return int(mlib().gport_num_threads())

    def logjoint_grad(self, z):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, self.D)
        K = z.shape[0]
        logp, grad = np.empty(K), np.empty((K, self.D))
        mlib().gport_logjoint_grad(C.byref(self.p), _ptr(z), K, _ptr(logp), _ptr(grad))
        return logp, grad

    def elbo_grad(self, mu, omega, eps):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        eps = np.ascontiguousarray(eps, dtype=np.float64).reshape(-1, self.D)
        K = eps.shape[0]
        grad, logp = np.empty(2 * self.D), np.empty(K)
        elbo = mlib().gport_elbo_grad(C.byref(self.p), _ptr(mu), _ptr(omega), _ptr(eps), K, _ptr(grad), _ptr(logp))
        return float(elbo), grad[:self.D], grad[self.D:], logp

    def advi_steps(self, theta, acc, n_steps, K, eta=0.1, pre=1.0, post=0.9, seed=0, first_step=0):
        return float(mlib().gport_advi_steps(C.byref(self.p), _ptr(theta), _ptr(acc), eta, pre, post, K, n_steps,
                                             seed, first_step))
