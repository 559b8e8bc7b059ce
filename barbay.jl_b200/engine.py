"""Python face of one ``bb_handle``: packs a DataArrays + model keyword arguments into a
``bb_desc`` and forwards to the C ABI.  All computation happens in libbarbay_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, model as _model
from ._lib import BarBayError


def _c_doubles(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    """One GPU shard of one ADVI problem.

    Parameters mirror what ``BarBay.vi.advi`` hands to the Turing model (src/vi.jl:172-178):
    the packed arrays, the model, its keyword arguments, plus ``samples_per_step``.
    """

    def __init__(self, data_arrays, model, model_kwargs=None, *, n_samples: int = 1, dtype: str = "f64",
                 seed: int = 0, device: int = -1, rank: int = 0, world: int = 1, corrected_ragged: bool = False,
                 n_devices: int = 1, probe_only: bool = False):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        self.model = _model.resolve(model)
        da = data_arrays
        kw = _model.normalise_kwargs(self.model, model_kwargs)
        self._keep = []      # host buffers referenced by the desc

        if isinstance(da.bc_count, (list, tuple)):
            mats = [np.asarray(m, dtype=np.int64) for m in da.bc_count]
            n_time = [m.shape[0] for m in mats]
            flat = np.concatenate([m.T.reshape(-1) for m in mats])         # each T_r x B column-major
        else:
            R = np.asarray(da.bc_count, dtype=np.int64)
            if R.ndim == 2:
                n_time = [R.shape[0]]
                flat = R.T.reshape(-1)
            else:
                n_time = [R.shape[0]] * R.shape[2]
                flat = R.transpose(2, 1, 0).reshape(-1)                     # Julia memory order of T x B x R
        n_rep = len(n_time)
        if self.model.replicate and n_rep < 1:
            raise BarBayError("replicate model needs replicate data")
        flat = np.ascontiguousarray(flat)
        nt_arr = np.asarray(n_time, dtype=np.int32)
        self._keep += [flat, nt_arr]

        desc = _lib.bb_desc()
        desc.abi_version = _lib.ABI_VERSION
        desc.model = _lib.MODEL_IDS[self.model.name]
        if dtype not in _lib.DTYPE_IDS:
            raise BarBayError(f"dtype must be one of {sorted(_lib.DTYPE_IDS)}")
        desc.dtype = _lib.DTYPE_IDS[dtype]
        desc.n_rep = n_rep
        desc.n_time = nt_arr.ctypes.data_as(C.POINTER(C.c_int32))
        desc.n_neutral = int(da.n_neutral)
        desc.n_bc = int(da.n_bc)
        desc.bc_count = flat.ctypes.data_as(C.POINTER(C.c_int64))
        n_env, n_geno = 1, 0
        if self.model.multienv:
            envs = kw.get("envs", da.envs)
            if isinstance(envs, str):
                raise BarBayError("Models with multiple environments need the list of environments (envs)")
            if len(envs) and isinstance(envs[0], (list, tuple)):
                # one list per replicate: the Vector{Matrix{Int64}} method (…hierarchical_replicates.jl:449-472)
                if self.model.name != "multienv_replicate_fitness_normal":
                    raise BarBayError("one environment list per replicate applies to multienv_replicate_fitness_normal")
                if len(envs) != n_rep or any(len(es) != t for es, t in zip(envs, n_time)):
                    raise BarBayError("Number of time points must match list of of environments for all replicates")   # :463-465
                flat = [e for es in envs for e in es]
                uniq, env_idx = _model.indexin_unique(flat)          # unique(vcat(envs...)), indexin.(envs, Ref(.)) :468-472
                desc.env_per_rep = 1
            else:
                if len(set(int(t) for t in n_time)) != 1:
                    raise BarBayError("replicates with unequal numbers of time points need one environment list per replicate")
                if len(envs) != n_time[0]:
                    raise BarBayError("Number of time points must match list of of environments")   # multienv.jl:146-148
                uniq, env_idx = _model.indexin_unique(list(envs))
            n_env = len(uniq)
            env_idx = np.ascontiguousarray(env_idx, dtype=np.int32)
            self._keep.append(env_idx)
            desc.env_idx = env_idx.ctypes.data_as(C.POINTER(C.c_int32))
        if self.model.genotype:
            genotypes = kw.get("genotypes", da.genotypes)
            if isinstance(genotypes, str) or len(genotypes) != da.n_bc:
                raise BarBayError("List of genotypes must match number of barcodes")             # genotypes.jl:165-167
            uniq, gidx = _model.indexin_unique(list(genotypes))
            n_geno = len(uniq)
            gidx = np.ascontiguousarray(gidx, dtype=np.int32)
            self._keep.append(gidx)
            desc.geno_idx = gidx.ctypes.data_as(C.POINTER(C.c_int32))
        desc.n_env, desc.n_geno = n_env, n_geno

        def prior(key):
            p = _lib.bb_prior()
            val = kw.get(key, _model.DEFAULT_PRIORS[key])
            a = np.asarray(val, dtype=np.float64)
            if a.ndim == 1:
                if a.size != 2:
                    raise BarBayError(f"{key}: vector priors are [mean, std]")
                buf = np.ascontiguousarray(a)
                p.n, p.is_matrix = 2, 0
            elif a.ndim == 2 and a.shape[1] == 2:
                buf = np.ascontiguousarray(a.T.reshape(-1))                  # Julia column-major n x 2
                p.n, p.is_matrix = a.shape[0], 1
            else:
                raise BarBayError(f"{key}: priors are [mean, std] or an n x 2 matrix")
            self._keep.append(buf)
            p.data = _c_doubles(buf)
            return p

        desc.s_pop_prior = prior("s_pop_prior")
        desc.logsig_pop_prior = prior("logσ_pop_prior")
        desc.s_bc_prior = prior("s_bc_prior")
        desc.logsig_bc_prior = prior("logσ_bc_prior")
        desc.loglam_prior = prior("logλ_prior")
        desc.logtau_prior = prior("logτ_prior")
        desc.ragged_as_written = 0 if corrected_ragged else 1
        desc.n_samples = int(n_samples)
        desc.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        desc.device, desc.rank, desc.world = int(device), int(rank), int(world)
        desc.n_devices = int(n_devices)
        self.n_devices = int(n_devices)

        self.layout = _model.var_groups(self.model, n_time, n_rep, da.n_neutral, da.n_bc, n_env, n_geno)
        self.n_samples = int(n_samples)
        self.rank, self.world = int(rank), int(world)
        if probe_only:
            # host-only dry run of the shard layout (bb_layout_probe: no CUDA call, works without a GPU)
            D = self.layout.n_latent
            self.owned = np.zeros(D, dtype=np.int32)
            info = (C.c_int64 * 8)()
            rc = self._lib.bb_layout_probe(C.byref(desc), self.owned.ctypes.data_as(C.POINTER(C.c_int32)), info)
            if rc != 0:
                raise BarBayError(self._lib.bb_last_error(None).decode("utf8", "replace"))
            keys = ("D", "n0", "n1", "m0", "m1", "H", "hy_gid0", "cpad")
            self.probe_info = dict(zip(keys, [int(x) for x in info]))
            self.D = self.probe_info["D"]
            return
        rc = self._lib.bb_create(C.byref(desc), C.byref(self._h))
        if rc != 0:
            msg = self._lib.bb_last_error(None).decode("utf8", "replace")
            self._h = C.c_void_p()
            raise BarBayError(msg)
        self.D = int(self._lib.bb_n_latent(self._h))
        if self.D != self.layout.n_latent:
            raise BarBayError(f"latent count mismatch: library {self.D}, host layout {self.layout.n_latent}")

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int):
        if rc != 0:
            raise BarBayError(self._lib.bb_last_error(self._h).decode("utf8", "replace"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ parameters
    def init_params(self, seed: int = 0):
        self._check(self._lib.bb_init_params(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_params(self, mu, omega):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        if mu.size != self.D or omega.size != self.D:
            raise BarBayError("mu / omega must have length D")
        self._check(self._lib.bb_set_params(self._h, _c_doubles(mu), _c_doubles(omega)))

    def get_params(self):
        mu, om = np.empty(self.D), np.empty(self.D)
        self._check(self._lib.bb_get_params(self._h, _c_doubles(mu), _c_doubles(om)))
        return mu, om

    def get_posterior(self):
        m, s = np.empty(self.D), np.empty(self.D)
        self._check(self._lib.bb_get_posterior(self._h, _c_doubles(m), _c_doubles(s)))
        return m, s

    # ------------------------------------------------------------------ parity entry points
    def logjoint_grad(self, x, eps_is_noise: bool = False):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.D)
        K = x.shape[0]
        logp, grad = np.empty(K), np.empty((K, self.D))
        self._check(self._lib.bb_logjoint_grad(self._h, _c_doubles(x), K, 1 if eps_is_noise else 0,
                                               _c_doubles(logp), _c_doubles(grad)))
        return logp, grad

    def elbo_grad(self, eps=None, step: int = 0, want_grad: bool = True):
        """ELBO estimate and its gradient w.r.t. (mu, omega) at the current parameters (no update).
        ``want_grad=False``: the gradient stays on the device (returns ``(elbo, None, None)``) -- what bench.py times
        as the pure ELBO-gradient evaluation."""
        elbo = C.c_double()
        if not want_grad:
            p = None if eps is None else _c_doubles(np.ascontiguousarray(eps, dtype=np.float64).reshape(self.n_samples, self.D))
            self._check(self._lib.bb_elbo_grad(self._h, p, int(step), C.byref(elbo), None))
            return elbo.value, None, None
        grad = np.empty(2 * self.D)
        if eps is not None:
            eps = np.ascontiguousarray(eps, dtype=np.float64).reshape(self.n_samples, self.D)
            p = _c_doubles(eps)
        else:
            p = None
        self._check(self._lib.bb_elbo_grad(self._h, p, int(step), C.byref(elbo), _c_doubles(grad)))
        return elbo.value, grad[:self.D], grad[self.D:]

    def get_noise(self, step: int = 0):
        eps = np.empty((self.n_samples, self.D))
        self._check(self._lib.bb_get_noise(self._h, int(step), _c_doubles(eps)))
        return eps

    # ------------------------------------------------------------------ optimisation
    def set_optimizer(self, kind: str = "truncated", eta: float = 0.1, tau: float = 1.0, n: int = 100,
                      pre: float = 1.0, post: float = 0.9):
        o = _lib.bb_opt()
        if kind in ("truncated", "TruncatedADAGrad"):
            o.kind, o.eta, o.tau, o.post, o.n = _lib.OPT_TRUNCATED, eta, tau, post, n
        elif kind in ("decayed", "DecayedADAGrad"):
            o.kind, o.eta, o.tau, o.post, o.n = _lib.OPT_DECAYED, eta, pre, post, n
        else:
            raise BarBayError("opt must be TruncatedADAGrad or DecayedADAGrad")
        self._check(self._lib.bb_set_optimizer(self._h, C.byref(o)))

    def step(self, n_steps: int = 1, elbo_trace: bool = False):
        if elbo_trace:
            tr = np.empty(n_steps)
            self._check(self._lib.bb_step(self._h, int(n_steps), _c_doubles(tr)))
            return tr
        self._check(self._lib.bb_step(self._h, int(n_steps), None))
        return None

    def step_until(self, max_iters: int, every: int = 100, window: int = 5, rel_tol: float = 1e-4):
        """Run until the ELBO estimate (one every ``every`` steps) stops moving: returns (steps done, converged,
        ELBO estimates).  An extension -- the reference always runs ``max_iters`` (src/vi.jl:98)."""
        n_done, conv, n_elbo = C.c_int32(), C.c_int32(), C.c_int32()
        est = np.empty(max(1, max_iters // max(every, 1) + 1))
        self._check(self._lib.bb_step_until(self._h, int(max_iters), int(every), int(window), float(rel_tol),
                                            C.byref(n_done), C.byref(conv), _c_doubles(est), C.byref(n_elbo)))
        return int(n_done.value), bool(conv.value), est[:n_elbo.value].copy()

    def step_with_noise(self, eps):
        eps = np.ascontiguousarray(eps, dtype=np.float64).reshape(self.n_samples, self.D)
        self._check(self._lib.bb_step_with_noise(self._h, _c_doubles(eps)))

    @property
    def step_count(self) -> int:
        return int(self._lib.bb_step_count(self._h))

    def get_state(self):
        s = np.empty(int(self._lib.bb_state_size(self._h)))
        self._check(self._lib.bb_get_state(self._h, _c_doubles(s)))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        self._check(self._lib.bb_set_state(self._h, _c_doubles(s)))

    # ------------------------------------------------------------------ plumbing
    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.bb_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def sync(self):
        self._check(self._lib.bb_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.bb_launch_count(self._h))

    @property
    def algorithmic_bytes_per_step(self) -> float:
        return float(self._lib.bb_algorithmic_bytes_per_step(self._h))

    def time_steps(self, n_steps: int):
        """(ms_total, ms_pass1, ms_pass2) for n_steps, CUDA events on the launching stream."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._check(self._lib.bb_time_steps(self._h, int(n_steps), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def derived_fitness(self, n_samples: int = 10_000, seed: int = 0):
        """(median, sd) of ``n_samples`` draws of θ + exp(logτ) θ̃ per logτ row, from the current posterior, on the
        device (utils.jl:1284-1343; hierarchical models only)."""
        n = int(self._lib.bb_n_derived(self._h))
        med, sd = np.zeros(n), np.zeros(n)
        self._check(self._lib.bb_derived_fitness(self._h, int(n_samples), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                 _c_doubles(med), _c_doubles(sd)))
        return med, sd

    def data_plane(self) -> dict:
        """Which kernels / exchange this handle runs (for reporting)."""
        out = (C.c_int32 * 8)()
        self._check(self._lib.bb_data_plane(self._h, out))
        return {"step_kernel": bool(out[0]), "persistent": out[1] > 1, "persist_chunk": int(out[1]),
                "peer_exchange": bool(out[2]), "nccl": bool(out[3]), "ctas_per_sm": int(out[4]),
                "staging_buffers": int(out[5]), "acc_staged": bool(out[6]), "grid": int(out[7])}

    def persist_stats(self) -> dict:
        """Per-step breakdown of the persistent step kernel since the last call (microseconds, CTA 0's clock):
        column phase, arrival -> all ranks' sums (grid reduction + NVLink exchange), sums -> next context."""
        out = np.zeros(16)
        self._check(self._lib.bb_persist_stats(self._h, _c_doubles(out)))
        n, khz = out[3], out[4]
        if n <= 0 or khz <= 0:
            return {"tails": 0}
        us = lambda cyc: cyc / n / khz * 1e3
        return {"tails": int(n), "column_us": us(out[0]), "exchange_us": us(out[1]), "context_us": us(out[2]),
                "context_sums_us": us(out[5]), "context_shared_us": us(out[6]), "reduce_post_us": us(out[7]),
                "column_last_us": us(out[8]),
                "chain_us": [us(out[9]), us(out[10]), us(out[11]), us(out[12])]}

    def peer_handle(self) -> bytes:
        """CUDA IPC handle of this rank's exchange buffer (64 bytes): gather them in rank order, then ``peer_attach``."""
        buf = (C.c_char * 64)()
        self._check(self._lib.bb_peer_handle(self._h, buf))
        return bytes(buf.raw)

    def peer_attach(self, handles):
        """Wire the per-step exchange over NVLink peer memory from the ranks' handles (no NCCL communicator)."""
        blob = b"".join(handles)
        if len(blob) != 64 * len(handles):
            raise BarBayError("peer handles are 64 bytes each")
        self._check(self._lib.bb_peer_attach(self._h, blob, len(handles)))

    def comm_init(self, unique_id: bytes):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._check(self._lib.bb_comm_init(self._h, buf))


def comm_unique_id() -> bytes:
    lib = _lib.load()
    buf = (C.c_char * 128)()
    if lib.bb_comm_unique_id(buf) != 0:
        raise BarBayError(lib.bb_last_error(None).decode("utf8", "replace"))
    return bytes(buf.raw)
