"""`GP` (mirror of gpfy/model.py:33-302): init / conditional / mu / var / predict_diag / fit.

`conditional(sh, q)` returns a GP that carries the same four closures as the reference AND remembers
(sh, q), so `predict_diag` runs as ONE fused device call (`gpfy_predict_diag`) instead of materialising
the [M, N] projection; the closures remain available (and are tested to agree with the fused path, like
tests/test_model.py:160-183 does for the reference).  `cov` / `predict` build the [P, N, N] covariance on the device (csrc/kmat.cu)."""
import dataclasses
from typing import Callable, Optional, Tuple

import numpy as np

from . import _runtime
from .param import Param


@dataclasses.dataclass(frozen=True)
class GP:
    kernel: object
    conditional_fn: Tuple[Callable, ...] = ()
    name: str = "GP"
    _sh: object = None
    _q: object = None

    def init(self, key, input_dim: int, num_independent_processes: int = 1, projection_dim: Optional[int] = None,
             likelihood=None, sh_features=None, variational_dist=None) -> Param:
        """model.py:60-101."""
        rng = _runtime.as_rng(key)
        k1, k2 = rng.spawn(2)
        param = self.kernel.init(k2, input_dim, projection_dim)
        param = likelihood.init(param) if likelihood else param
        if variational_dist and sh_features:
            param = sh_features.init(k1, projection_dim or input_dim, param)
            param = variational_dist.init(sh_features.num_inducing(param), num_independent_processes, param)
        return param

    def conditional(self, sh, q) -> "GP":
        """model.py:103-145."""
        proj_fun, cond_cov, cond_var = sh.conditional_fun(self.kernel)
        cond_mean_fun = lambda param, proj: q.project_mean(param, proj)
        cond_cov_fun = lambda param, x, proj: cond_cov(param, x, proj) + q.project_variance(param, proj)
        cond_var_fun = lambda param, x, proj: cond_var(param, x, proj)[..., None] + q.project_diag_variance(param, proj)
        return GP(self.kernel, (proj_fun, cond_mean_fun, cond_cov_fun, cond_var_fun), _sh=sh, _q=q)

    def mu(self, param: Param):
        """model.py:147-163."""
        if self.conditional_fn:
            proj_fun, cond_mean_fun, _, __ = self.conditional_fn
            return lambda x: cond_mean_fun(param, proj_fun(param, x))
        return lambda x: _zeros(x)

    def var(self, param: Param):
        """model.py:165-181."""
        if self.conditional_fn:
            proj_fun, _, __, cond_var_fun = self.conditional_fn
            return lambda x: cond_var_fun(param, x, proj_fun(param, x))
        return lambda x: self.kernel.K_diag(param, x)

    def cov(self, param: Param):
        """model.py:183-199: the [P, N, N] conditional covariance, or the prior K(X) [N, N]."""
        if self.conditional_fn:
            return lambda x: self.predict(param, x)[1]
        return lambda x: self.kernel.K(param, x)

    def predict(self, param: Param, X):
        """(mean [N, P], covariance [P, N, N]) — model.py:231-259.  K(X) - A^T A + A^T S A is formed as
        K + A^T (S - I) A in one pass over the output (gpfy_project_variance with base = K)."""
        if self.conditional_fn:
            if self._sh is not None:
                sh, q, kernel = self._sh, self._q, self.kernel
                A = self.conditional_fn[0](param, X)
                K = kernel.K(param, X)
                mean, _ = q._hot(param).project_q(A, q.mu(param), q.cov_part(param))
                return mean, q._hot(param).project_variance(A, q.cov_part(param), subtract_identity=True, base=K)
            proj_fun, cond_mean_fun, cond_cov_fun, _ = self.conditional_fn
            proj = proj_fun(param, X)
            return cond_mean_fun(param, proj), cond_cov_fun(param, X, proj)
        return _zeros(X), self.kernel.K(param, X)

    # ---- the fused path ----
    def hot_args(self, param: Param, lik=None):
        """(HotPath, constrained device-call arguments) of this conditional model: w, b, variance, eig, vhat, lcat, mu, sigma."""
        sh, q, kernel = self._sh, self._q, self.kernel
        if sh is None or q is None:
            raise _runtime.GpfyError("the fused path needs a GP built by `GP(kernel).conditional(sh, q)`")
        kp = param.params[kernel.name]
        vhat, lcat, m = sh._cat(param)
        P = q.mu(param).shape[1]
        wv = kp["weight_variances"]
        input_dim = int(np.shape(wv)[0]) if np.ndim(wv) >= 1 else int(vhat.shape[1]) - 1
        hp = _runtime.hot_path(input_dim, m, num_outputs=P, kernel_order=kernel.order,
                               likelihood=lik._kind if lik is not None else "gaussian", q_kind=q._q_kind,
                               weight_mode=kernel._weight_mode(param), projection_dim=kernel._projection_dim(param))
        args = (kp["weight_variances"], kp["bias_variance"], kp["variance"], kernel.eigenvalues(param, sh.levels), vhat, lcat,
                q.mu(param), q.cov_part(param))
        return hp, args

    def predict_diag(self, param: Param, X):
        """(mean [N, P], diagonal variance [N, P]) — model.py:201-229."""
        if self.conditional_fn:
            if self._sh is not None:
                hp, args = self.hot_args(param)
                return hp.predict_diag(X, *args)
            proj_fun, cond_mean_fun, _, cond_var_fun = self.conditional_fn
            proj = proj_fun(param, X)
            return cond_mean_fun(param, proj), cond_var_fun(param, X, proj)
        return _zeros(X), self.kernel.K_diag(param, X)[..., None]

    def fit(self, param: Param, train_step, optimizer, num_iters: int, progress_bar: bool = True):
        """model.py:261-302: optimise the unconstrained parameters; frozen leaves get zero updates."""
        from .training import TrainState, masked
        param_free = param.unconstrained()
        state = TrainState.create(apply_fn=lambda p: param_free.replace(params=p), params=param_free.params,
                                  tx=masked(optimizer, param_free._trainables))
        losses = []
        it = range(num_iters)
        if progress_bar:
            try:
                from tqdm import tqdm
                it = tqdm(it)
            except ImportError:
                pass
        for _ in it:
            state, loss = train_step(state, None)
            losses.append(loss)
        return state.apply_fn(state.params).constrained(), state, np.asarray(losses)


def _zeros(x):
    import torch
    n = int(np.shape(x)[0])
    dev = x.device if torch.is_tensor(x) and x.is_cuda else None
    return torch.zeros((n, 1), dtype=torch.float64, device=dev)
