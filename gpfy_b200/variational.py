"""Variational distributions q(u) (mirror of gpfy/variational.py): mu [M, P], "Sigma" [P, M, M].

Projections of an explicit projection matrix A [M, N] run on the device (`gpfy_project_q`); the
whitened-prior KL of the TriL parameterisation too (`gpfy_prior_kl_valgrad`).  O(M^2) accessors
(`logdet`, `trace`) are host numpy.  `project_variance` ([P, N, N]) is `gpfy_project_variance`."""
import dataclasses
from typing import Optional

import numpy as np

from . import _runtime
from .param import FillTriangular, Param, identity, tree_leaves


@dataclasses.dataclass(frozen=True)
class VariationalDistribution:
    name: str = "VariationalDistribution"
    _q_kind = None

    def _define_covariance_bijector(self):
        raise NotImplementedError()

    def init(self, num_inducing_features: int, num_independent_processes: int = 1, param: Optional[Param] = None) -> Param:
        """variational.py:55-121: mu = 0 [M, P], Sigma = I [P, M, M]."""
        feats = tree_leaves(param.params.get("variational")) if param and param.params.get("variational") else None
        if feats:
            if num_inducing_features != sum(np.shape(x)[0] for x in feats):
                raise ValueError("`param` object contains different number of inducing features.")
        collection = "variational"
        M, P = int(num_inducing_features), int(num_independent_processes)
        var_params = {"mu": np.zeros((M, P)), "Sigma": np.stack([np.eye(M) for _ in range(P)])}
        var_bijectors = {"mu": identity(), "Sigma": self._define_covariance_bijector()}
        all_params, all_trainables, all_bijectors, all_constants, constrained = {collection: {}}, {collection: {}}, {collection: {}}, {}, True
        if param:
            all_params = {k: dict(v) if isinstance(v, dict) else v for k, v in param.params.items()}
            all_bijectors = {k: dict(v) if isinstance(v, dict) else v for k, v in param._bijectors.items()}
            all_trainables = {k: dict(v) if isinstance(v, dict) else v for k, v in param._trainables.items()}
            all_constants, constrained = dict(param.constants), param._constrained
            all_params.setdefault(collection, {}); all_bijectors.setdefault(collection, {}); all_trainables.setdefault(collection, {})
        all_params[collection][self.name] = var_params
        all_bijectors[collection][self.name] = var_bijectors
        all_trainables[collection].pop(self.name, None)
        return Param(params=all_params, _trainables=all_trainables, _bijectors=all_bijectors, constants=all_constants,
                     _constrained=constrained)

    def mu(self, param: Param):
        return param.params["variational"][self.name]["mu"]

    def cov_part(self, param: Param):
        return param.params["variational"][self.name]["Sigma"]

    def _hot(self, param: Param):
        M, P = self.mu(param).shape
        return _runtime.hot_path(1, [M], num_outputs=P, q_kind=self._q_kind)

    def project_mean(self, param: Param, A):
        """A^T mu -> [N, P] (variational.py:288-305 / :401-418)."""
        return self._hot(param).project_q(A, self.mu(param), self.cov_part(param))[0]

    def project_diag_variance(self, param: Param, A):
        """diag(A^T S A) -> [N, P] (variational.py:307-328 / :420-441)."""
        return self._hot(param).project_q(A, self.mu(param), self.cov_part(param))[1]

    def project_variance(self, param: Param, A):
        """A^T S A -> [P, N, N] (variational.py:330-351 / :443-464)."""
        return self._hot(param).project_variance(A, self.cov_part(param))

    def logdet(self, param: Param):
        raise NotImplementedError()

    def trace(self, param: Param):
        raise NotImplementedError()

    def prior_KL(self, param: Param) -> float:
        """variational.py:225-240 (whitened prior)."""
        return self.prior_KL_and_grad(param)[0]

    def prior_KL_and_grad(self, param: Param):
        """(KL, dKL/dmu [M, P], dKL/dSigma [P, M, M]) w.r.t. the constrained mu / Sigma."""
        mu, S = self.mu(param), self.cov_part(param)
        KL = -0.5 * float(np.size(mu)) - 0.5 * float(np.sum(self.logdet(param)))
        KL += 0.5 * float(np.sum(np.square(mu))) + 0.5 * float(np.sum(self.trace(param)))
        return KL, mu.copy(), self._dkl_dcov(S)


@dataclasses.dataclass(frozen=True)
class VariationalDistributionTriL(VariationalDistribution):
    """q(u) = N(mu, L L^T) (variational.py:243-351)."""
    name: str = "VariationalDistributionTriL"
    _q_kind = "tril"

    def _define_covariance_bijector(self):
        return FillTriangular()

    def logdet(self, param: Param):
        L = self.cov_part(param)
        return np.sum(np.log(np.square(np.diagonal(L, axis1=-2, axis2=-1))), axis=-1)

    def trace(self, param: Param):
        return np.sum(np.square(self.cov_part(param)), axis=(-1, -2))

    def prior_KL_and_grad(self, param: Param):
        kl, dmu, dsig = self._hot(param).prior_kl_valgrad(self.mu(param), self.cov_part(param))
        return float(kl), dmu.cpu().numpy(), dsig.cpu().numpy()


@dataclasses.dataclass(frozen=True)
class VariationalDistributionFullCovariance(VariationalDistribution):
    """q(u) = N(mu, Sigma) (variational.py:354-464).  The KL (Cholesky, logdet, Sigma^-1: O(M^3), N-independent) runs on the device
    through gpfy_prior_kl_valgrad_ws; `logdet` / `trace` below are the host mirrors of the reference's methods."""
    name: str = "VariationalDistributionFullCovariance"
    _q_kind = "fullcov"

    def _define_covariance_bijector(self):
        return identity()

    def logdet(self, param: Param):
        L = np.linalg.cholesky(self.cov_part(param))
        return np.sum(np.log(np.square(np.diagonal(L, axis1=-2, axis2=-1))), axis=-1)

    def trace(self, param: Param):
        return np.trace(self.cov_part(param), axis1=-2, axis2=-1)

    def prior_KL_and_grad(self, param: Param):
        kl, dmu, dsig = self._hot(param).prior_kl_valgrad(self.mu(param), self.cov_part(param))
        return float(kl), dmu.cpu().numpy(), dsig.cpu().numpy()

    def _dkl_dcov(self, S):
        # d/dS [-1/2 logdet S + 1/2 tr S]; the Cholesky VJP symmetrises its cotangent like jnp.linalg.cholesky
        inv = np.linalg.inv(S)
        return 0.5 * np.eye(S.shape[-1])[None] - 0.25 * (inv + np.swapaxes(inv, -1, -2))
