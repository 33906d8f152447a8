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
        raise NotImplementedError("TFP distribution objects are not part of the accelerated path; use log_prob")


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


@dataclasses.dataclass(frozen=True)
class Bernoulli(Likelihood):
    name: str = "Bernoulli"
    _kind = "bernoulli"


def inv_probit(x):
    """likelihoods.py:281-292 (host numpy; the device kernel applies the same map inside the quadrature)."""
    from scipy.special import erf
    jitter = 1e-3
    return 0.5 * (1.0 + erf(np.asarray(x) / np.sqrt(2.0))) * (1 - 2 * jitter) + jitter
