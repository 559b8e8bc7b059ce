"""Oracle ADVI engine: restatement of Turing.vi / AdvancedVI 0.2 semantics.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED.

The reference calls ``Turing.vi(bayes_model, advi; optimizer=opt)`` at
src/vi.jl:201; the engine behind it (AdvancedVI 0.2.x, Turing 0.36 -- not
vendored, versions pinned only by Project.toml:26-43 compat ranges) is restated
here from its published source:

  meanfield      mu0 = randn(D); sigma0 = softplus.(randn(D));
                 theta = vcat(mu0, invsoftplus.(sigma0))          (Turing variational/advi.jl)
  ELBO           (1/K) sum_k [logpi(z_k) + logjac_k] + entropy(q_base),
                 z_k = mu + softplus(omega) .* eps_k, logjac = 0 (all bijectors identity),
                 entropy = D (1 + log 2pi) / 2 + sum log sigma    (AdvancedVI advi.jl)
  optimize!      max_iters times: grad of -ELBO; D = apply!(opt, theta, grad); theta -= D
  TruncatedADAGrad(eta=0.1, tau=1.0, n=100): ring of the last n squared gradients,
                 s = sum(ring); D = eta * g / (tau + sqrt(s) + 1e-8)
  DecayedADAGrad(eta=0.1, pre=1.0, post=0.9): acc = post*acc + pre*g^2 (acc0 = 1e-8);
                 D = eta * g / (sqrt(acc) + 1e-8)                  (AdvancedVI optimisers.jl)
  update         q.dist.m = theta[1:D]; q.dist.sigma = softplus.(theta[D+1:2D])
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import model_ref, philox_ref

F64 = torch.float64
EPS = 1e-8


def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 0, x + np.log1p(np.exp(-np.abs(x))), np.log1p(np.exp(-np.abs(x))))


def sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def entropy_diag_normal(sigma: np.ndarray) -> float:
    D = sigma.size
    return 0.5 * D * (1.0 + math.log(2.0 * math.pi)) + float(np.log(sigma).sum())


def elbo_value_and_grad(model: str, prob: dict, mu: np.ndarray, omega: np.ndarray, eps: np.ndarray):
    """ELBO(theta) for caller-supplied noise and its gradient w.r.t. (mu, omega).

    Independent route: the whole ELBO (softplus, reparameterisation, entropy) is
    built in torch.float64 and differentiated by autograd -- no analytic formula.
    Returns elbo, grad_mu, grad_omega (gradients of +ELBO), logp per sample.
    """
    mu_t = torch.tensor(mu, dtype=F64, requires_grad=True)
    om_t = torch.tensor(omega, dtype=F64, requires_grad=True)
    eps_t = torch.as_tensor(np.asarray(eps, dtype=np.float64))
    sigma = torch.nn.functional.softplus(om_t)
    K = eps_t.shape[0]
    D = mu_t.numel()
    total = torch.zeros((), dtype=F64)
    logps = []
    for k in range(K):
        z = mu_t + sigma * eps_t[k]
        lp = model_ref.logjoint(model, z, prob)
        logps.append(float(lp.detach()))
        total = total + lp / K
    elbo = total + 0.5 * D * (1.0 + math.log(2.0 * math.pi)) + torch.log(sigma).sum()
    g_mu, g_om = torch.autograd.grad(elbo, (mu_t, om_t))
    return float(elbo.detach()), g_mu.numpy().copy(), g_om.numpy().copy(), np.asarray(logps)


@dataclass
class TruncatedADAGrad:
    eta: float = 0.1
    tau: float = 1.0
    n: int = 100
    ring: np.ndarray | None = None
    iters: int = 1

    def apply(self, g: np.ndarray) -> np.ndarray:
        if self.ring is None:
            self.ring = np.zeros((self.n, g.size), dtype=np.float64)
        idx = (self.iters - 1) % self.n
        self.ring[idx] = g ** 2
        s = self.ring.sum(axis=0)
        self.iters += 1
        return g * (self.eta / (self.tau + np.sqrt(s) + EPS))


@dataclass
class DecayedADAGrad:
    eta: float = 0.1
    pre: float = 1.0
    post: float = 0.9
    acc: np.ndarray | None = None

    def apply(self, g: np.ndarray) -> np.ndarray:
        if self.acc is None:
            self.acc = np.full(g.size, EPS, dtype=np.float64)
        self.acc = self.post * self.acc + self.pre * g ** 2
        return g * (self.eta / (np.sqrt(self.acc) + EPS))


@dataclass
class AdviTrace:
    mu: np.ndarray
    omega: np.ndarray
    elbo: list = field(default_factory=list)

    @property
    def sigma(self):
        return softplus(self.omega)


def advi_run(model: str, prob: dict, n_steps: int, n_samples: int, opt, mu0, omega0,
             eps_fn=None, seed: int = 0, first_step: int = 0) -> AdviTrace:
    """AdvancedVI.optimize! with the ELBO above.

    ``eps_fn(step) -> eps[K, D]`` supplies the noise; default is the Philox noise
    lattice of oracle/philox_ref.py keyed by ``seed``.
    """
    mu = np.array(mu0, dtype=np.float64)
    om = np.array(omega0, dtype=np.float64)
    D = mu.size
    tr = AdviTrace(mu, om)
    for i in range(n_steps):
        step = first_step + i
        eps = eps_fn(step) if eps_fn is not None else philox_ref.noise(model, prob, n_samples, step, seed)
        elbo, g_mu, g_om, _ = elbo_value_and_grad(model, prob, mu, om, eps)
        g = -np.concatenate([g_mu, g_om])         # gradient of the objective -ELBO
        delta = opt.apply(g)
        mu -= delta[:D]
        om -= delta[D:]
        tr.elbo.append(elbo)
    tr.mu, tr.omega = mu, om
    return tr
