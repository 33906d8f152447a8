"""Likelihoods (mirror of gpfy/likelihoods.py): same classes, `init`, `variational_expectations(param,
Fmu, Fvar)(Y)` and `log_prob`; the per-point arithmetic runs in `gpfy_variational_expectations`
(closed form for the Gaussian, likelihoods.py:188-220; 20-node Gauss-Hermite of the probit Bernoulli
log-density, likelihoods.py:254-292 + quadrature.py:30-53)."""
import dataclasses
from typing import Optional

import numpy as np

from . import _runtime
from .param import Param, positive


@dataclasses.dataclass(frozen=True)
class Likelihood:
    name: str = "Likelihood"
    _kind = None

    def init(self, param: Optional[Param] = None) -> Param:
        return param or Param()

    def _sigma2(self, param: Param):
        return None

    def variational_expectations(self, param: Param, Fmu, Fvar):
        P = int(np.shape(Fmu)[-1])
        hp = _runtime.hot_path(1, [1], num_outputs=P, likelihood=self._kind)
        s2 = self._sigma2(param)
        return lambda Y: hp.variational_expectations(Fmu, Fvar, Y, s2)

    def log_prob(self, param: Param, F, Y):
        """log p(Y | F) summed over outputs (likelihoods.py:65-79) = the expectation at zero variance."""
        import torch
        zeros = torch.zeros_like(F) if torch.is_tensor(F) else np.zeros_like(np.asarray(F, dtype=np.float64))
        return self.variational_expectations(param, F, zeros)(Y)

    def conditional_distribution(self, param: Param, F):
        raise NotImplementedError


@dataclasses.dataclass(frozen=True)
class NormalDistribution:
    """Host-side stand-in for `tfp.distributions.Normal(loc, scale)` (TFP is not installable here): the members GPfY's
    callers use on the object `conditional_distribution` returns (likelihoods.py:172-186)."""
    loc: np.ndarray
    scale: np.ndarray

    def mean(self):
        return np.asarray(self.loc, dtype=np.float64)

    def stddev(self):
        return np.broadcast_to(np.asarray(self.scale, dtype=np.float64), np.shape(self.loc)).copy()

    def variance(self):
        return self.stddev() ** 2

    def log_prob(self, y):
        z = (np.asarray(y, dtype=np.float64) - self.mean()) / self.scale
        return -0.5 * z * z - np.log(self.scale) - 0.5 * np.log(2.0 * np.pi)

    def sample(self, sample_shape=(), seed=0):
        rng = np.random.default_rng(seed)
        return self.mean() + self.scale * rng.standard_normal(tuple(np.atleast_1d(sample_shape).astype(int)) + np.shape(self.loc))


@dataclasses.dataclass(frozen=True)
class BernoulliDistribution:
    """Host-side stand-in for `tfp.distributions.Bernoulli(probs=...)` (likelihoods.py:236-252)."""
    probs: np.ndarray

    def mean(self):
        return np.asarray(self.probs, dtype=np.float64)

    def variance(self):
        p = self.mean()
        return p * (1.0 - p)

    def log_prob(self, y):
        y, p = np.asarray(y, dtype=np.float64), self.mean()
        return y * np.log(p) + (1.0 - y) * np.log1p(-p)

    def sample(self, sample_shape=(), seed=0):
        rng = np.random.default_rng(seed)
        return (rng.random(tuple(np.atleast_1d(sample_shape).astype(int)) + np.shape(self.probs)) < self.mean()).astype(np.float64)


@dataclasses.dataclass(frozen=True)
class Gaussian(Likelihood):
    name: str = "Gaussian"
    _kind = "gaussian"

    def init(self, param: Optional[Param] = None) -> Param:
        """likelihoods.py:141-170."""
        collection = "likelihood"
        all_params, all_trainables, all_bijectors, all_constants, constrained = {}, {collection: {}}, {}, {}, True
        if param:
            all_params, all_bijectors = dict(param.params), dict(param._bijectors)
            all_trainables, all_constants, constrained = dict(param._trainables), dict(param.constants), param._constrained
        all_params[collection] = {self.name: {"variance": np.array(1.0)}}
        all_bijectors[collection] = {self.name: {"variance": positive()}}
        return Param(params=all_params, _trainables=all_trainables, _bijectors=all_bijectors, constants=all_constants,
                     _constrained=constrained)

    def _sigma2(self, param: Param):
        return param.params["likelihood"][self.name]["variance"]

    def conditional_distribution(self, param: Param, F):
        """p(Y | F) = N(F, sigma^2) (likelihoods.py:172-186), as a host-side `NormalDistribution`."""
        F = F.detach().cpu().numpy() if hasattr(F, "detach") else np.asarray(F, dtype=np.float64)
        return NormalDistribution(loc=F.astype(np.float64), scale=np.sqrt(np.asarray(self._sigma2(param), dtype=np.float64)))


@dataclasses.dataclass(frozen=True)
class Bernoulli(Likelihood):
    name: str = "Bernoulli"
    _kind = "bernoulli"

    def conditional_distribution(self, param: Param, F):
        """p(Y | F) = Bernoulli(inv_probit(F)) (likelihoods.py:236-252), as a host-side `BernoulliDistribution`."""
        F = F.detach().cpu().numpy() if hasattr(F, "detach") else np.asarray(F, dtype=np.float64)
        return BernoulliDistribution(probs=inv_probit(F))


def inv_probit(x):
    """likelihoods.py:281-292 (host numpy; the device kernel applies the same map inside the quadrature)."""
    from scipy.special import erf
    jitter = 1e-3
    return 0.5 * (1.0 + erf(np.asarray(x) / np.sqrt(2.0))) * (1 - 2 * jitter) + jitter
