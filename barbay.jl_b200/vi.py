"""``advi``: host-side mirror of ``BarBay.vi.advi`` (src/vi.jl:86-235).

Same keyword arguments, defaults, validation order and error messages, same
DataFrame / CSV output.  The one replaced statement is src/vi.jl:201
(``q = Turing.vi(bayes_model, advi; optimizer=opt)``): here the variational
optimisation runs in libbarbay_b200.so on a B200.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass

from . import model as _model, utils as _utils
from ._lib import BarBayError
from .engine import Engine

logger = logging.getLogger("BarBay.vi")


@dataclass
class ADVI:
    """``Turing.ADVI{AD}(samples_per_step, max_iters)`` (src/vi.jl:98).  The AD backend of the
    reference is irrelevant here: the kernels use the analytic gradient."""
    samples_per_step: int = 1
    max_iters: int = 10_000
    adtype: str = "analytic"


@dataclass
class TruncatedADAGrad:
    """``Turing.Variational.TruncatedADAGrad(η=0.1, τ=1.0, n=100)`` -- the default optimiser (src/vi.jl:99)."""
    eta: float = 0.1
    tau: float = 1.0
    n: int = 100


@dataclass
class DecayedADAGrad:
    """``Turing.Variational.DecayedADAGrad(η=0.1, pre=1.0, post=0.9)``."""
    eta: float = 0.1
    pre: float = 1.0
    post: float = 0.9


def _apply_optimizer(eng: Engine, opt) -> None:
    if isinstance(opt, TruncatedADAGrad):
        eng.set_optimizer("truncated", eta=opt.eta, tau=opt.tau, n=opt.n)
    elif isinstance(opt, DecayedADAGrad):
        eng.set_optimizer("decayed", eta=opt.eta, pre=opt.pre, post=opt.post)
    else:
        raise TypeError("opt must be TruncatedADAGrad or DecayedADAGrad")     # Union type of src/vi.jl:99


def advi(*, data, model, outputname=None, model_kwargs=None, id_col="barcode", time_col="time",
         count_col="count", neutral_col="neutral", rep_col=None, env_col=None, genotype_col=None,
         advi=None, opt=None, verbose=True,
         # backend extras (not part of the reference signature; defaults keep its behaviour)
         seed=0, dtype="f64", device=-1, n_devices=1, n_posterior_samples=10_000, return_engine=False,
         elbo_rel_tol=None, elbo_every=100, elbo_window=5, device_derived_rows=True, corrected_ragged=False):
    """Fit the mean-field Gaussian posterior of a BarBay model with ADVI on a B200.

    Returns the tidy posterior ``DataFrame`` (columns ``mean, std, varname, vartype[, rep][, env], id``)
    or, when ``outputname`` is given, writes ``<outputname>.csv`` and returns ``None``.

    ``n_devices`` > 1 shards the barcodes over that many GPUs of this process inside the library (one handle, one
    blocking call per step batch, per-step exchange over NVLink peer memory): same posterior as one GPU.
    ``elbo_rel_tol`` (extension; the reference always runs ``max_iters``, src/vi.jl:98): stop early once the mean of
    the last ``elbo_window`` ELBO estimates (one every ``elbo_every`` steps) moves by less than that fraction.
    ``corrected_ragged``: replicates with unequal numbers of time points are fitted with the reference's neutral
    pairing as written (model_fitness_normal_hierarchical_replicates.jl:599-605: ratio (t, n) against
    ``s̄[⌈k / N⌉]``) by default; ``True`` pairs ratio (t, n) with ``s̄[t]`` like every other method of the reference
    does (required for ``n_devices`` > 1 on such data).
    """
    mdl = _model.resolve(model)
    advi_cfg = advi if advi is not None else ADVI(1, 10_000)
    opt = opt if opt is not None else TruncatedADAGrad()
    model_kwargs = dict(model_kwargs or {})

    fname = None if outputname is None else f"{outputname}.csv"
    if fname is not None and os.path.isfile(fname):
        raise BarBayError(f"{fname} was already processed")                               # vi.jl:103-108
    if "replicate" in str(mdl) and rep_col is None:
        raise BarBayError("Hierarchical models for experimental replicates require argument `:rep_col`")
    if "multienv" in str(mdl) and env_col is None:
        raise BarBayError("Models with multiple environments require argument `:env_col`")

    if verbose:
        logger.info("Pre-processing data...")
    data_arrays = _utils.data_to_arrays(data, id_col=id_col, time_col=time_col, count_col=count_col,
                                        neutral_col=neutral_col, rep_col=rep_col, env_col=env_col,
                                        genotype_col=genotype_col)
    if verbose:
        logger.info("Initialize Variational Inference Optimization...")
    if "multienv" in str(mdl):                                                            # vi.jl:146-156
        model_kwargs = {"envs": data_arrays.envs, **model_kwargs}
    if "genotype" in str(mdl):                                                            # vi.jl:159-169
        model_kwargs = {"genotypes": data_arrays.genotypes, **model_kwargs}

    eng = Engine(data_arrays, mdl, model_kwargs, n_samples=advi_cfg.samples_per_step, dtype=dtype, seed=seed,
                 device=device, n_devices=n_devices, corrected_ragged=corrected_ragged)
    var_names = eng.layout.var_names                                                      # vi.jl:184-198
    eng.init_params(seed)                                                                 # Turing meanfield()
    _apply_optimizer(eng, opt)
    if elbo_rel_tol is None:
        eng.step(advi_cfg.max_iters)                                                      # vi.jl:201
    else:
        n_done, conv, _ = eng.step_until(advi_cfg.max_iters, elbo_every, elbo_window, elbo_rel_tol)
        if verbose:
            logger.info("ADVI stopped after %d of %d steps (%s)", n_done, advi_cfg.max_iters,
                        "ELBO converged" if conv else "max_iters reached")
    m, sigma = eng.get_posterior()
    q = _utils.MeanFieldPosterior.build(m, sigma, eng.layout.ranges_out)

    # derived bc_fitness rows (utils.jl:1284-1343): sampled on the device from the fitted posterior
    derived = None
    if mdl.hier and device_derived_rows and (data_arrays.n_rep > 1 or genotype_col is not None):
        derived = eng.derived_fitness(min(int(n_posterior_samples), 12000), seed)
    df = _utils.advi_to_df(data, q, var_names, id_col=id_col, time_col=time_col, count_col=count_col,
                           neutral_col=neutral_col, rep_col=rep_col, env_col=env_col,
                           genotype_col=genotype_col, n_samples=n_posterior_samples, seed=seed,
                           output=data_arrays, derived=derived)
    if return_engine:
        return df, eng
    eng.close()
    if fname is None:
        return df
    df.to_csv(fname, index=False)                                                         # CSV.write vi.jl:216-232
    return None
