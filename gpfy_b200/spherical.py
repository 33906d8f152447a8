"""Spherical kernels (mirror of gpfy/spherical.py): `init`, `_scale_X`, `to_sphere`, `K_diag`, `eigenvalues`,
`shape_function`, `K`, `K2` for NTK / ArcCosine / PolynomialDecay.

Everything that touches the N data rows runs on the device (`to_sphere`, `K_diag`, `K`, `K2`, and `shape_function` of a
CUDA tensor).  `shape_function` of a host array is the N-independent parameter algebra of `init` (the Funk-Hecke
quadrature of the eigenvalues) and stays in numpy like the bijectors.
"""
import dataclasses
import math
from typing import Optional

import numpy as np

import torch

from . import _lib, _runtime, ops
from .gegenbauer import GegenbauerLookupTable, gegenbauer_host
from .harmonics.utils import funk_hecke_lambda
from .param import Param, identity


@dataclasses.dataclass(frozen=True)
class Spherical:
    order: int = 1
    ard: bool = True
    active_dims: Optional[object] = None
    name: str = "Spherical"

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)

    # ---- spherical.py:82-151 ----
    def init(self, key, input_dim: int, projection_dim: Optional[int] = None, max_num_eigvals: Optional[int] = None) -> Param:
        rng = _runtime.as_rng(key)
        bijectors = {}
        if not projection_dim:
            weight_variances = np.ones(input_dim) if self.ard else np.array(1.0)
        else:
            weight_variances = rng.standard_normal((input_dim, projection_dim))
            bijectors = {"weight_variances": identity()}
        params = {"weight_variances": weight_variances, "bias_variance": np.array(1.0), "variance": np.array(1.0)}
        sphere_dim = projection_dim or input_dim
        dim = sphere_dim + 1
        alpha = (dim - 2.0) / 2.0
        constants = {"sphere": {"sphere_dim": sphere_dim, "alpha": alpha}, self.name: {}}
        if isinstance(self, PolynomialDecay):
            params["beta"] = np.array(1.0)
            constants["sphere"]["gegenbauer_lookup_table"] = GegenbauerLookupTable(self.truncation_level, alpha)
        else:
            constants[self.name]["eigenvalues"] = self._compute_eigvals(sphere_dim, max_num_eigvals=max_num_eigvals)
        return Param(params={self.name: params}, _bijectors={self.name: bijectors}, constants=constants)

    def _compute_eigvals(self, sphere_dim: int, *, max_num_eigvals=None, gegenbauer_lookup_table=None):
        """spherical.py:153-180: Funk-Hecke eigenvalues for degrees 0..max_num_eigvals-1 (default 20)."""
        max_num_eigvals = max_num_eigvals or 20
        dim = sphere_dim + 1
        fn = lambda x: self.shape_function(Param(), x)
        return np.array([funk_hecke_lambda(fn, n, dim) for n in range(max_num_eigvals)])

    def eigenvalues(self, param: Param, levels) -> np.ndarray:
        """spherical.py:182-195: precomputed eigenvalues, zeros when absent."""
        levels = np.asarray(levels, dtype=np.int64)
        eigval = param.constants.get(self.name, {}).get("eigenvalues", np.zeros(len(levels)))
        return np.asarray(eigval)[levels]

    # ---- spherical.py:214-249, 384-400: evaluated on the device (ARD, scalar and [d, p] projection weights) ----
    def _weight_mode(self, param: Param) -> str:
        w = param.params[self.name]["weight_variances"]
        return "projection" if np.ndim(w) == 2 else ("scalar" if np.ndim(w) == 0 else "ard")

    def _projection_dim(self, param: Param) -> int:
        w = param.params[self.name]["weight_variances"]
        return int(np.shape(w)[1]) if np.ndim(w) == 2 else 0

    def to_sphere(self, param: Param, X):
        kp = param.params[self.name]
        d = int(np.shape(X)[-1])
        hp = _runtime.hot_path(d, [1], weight_mode=self._weight_mode(param), kernel_order=self.order,
                               projection_dim=self._projection_dim(param))
        return hp.to_sphere(X, kp["weight_variances"], kp["bias_variance"])

    def K_diag(self, param: Param, X):
        _, r = self.to_sphere(param, X)
        return float(param.params[self.name]["variance"]) * r ** (2 * self.order)

    def shape_function(self, param: Param, x):
        """spherical.py:251-263: derived classes implement it; CUDA tensors are evaluated on the device."""
        if torch.is_tensor(x) and x.is_cuda:
            alpha = param.constants["sphere"]["alpha"] if "sphere" in param.constants else 0.0
            return ops.shape_function(self._shape_fn(param), alpha, x, device=x.device)
        return self._shape_function_host(param, x)

    def _shape_function_host(self, param: Param, x):
        raise NotImplementedError

    def _shape_fn(self, param: Param):
        raise NotImplementedError

    def _kernel_matrix(self, param: Param, X, X2, unit_diagonal: bool):
        kp = param.params[self.name]
        d = int(np.shape(X)[-1])
        hp = _runtime.hot_path(d, [1], weight_mode=self._weight_mode(param), kernel_order=self.order,
                               projection_dim=self._projection_dim(param))
        return hp.kernel_matrix(self._shape_fn(param), X, X2, kp["weight_variances"], kp["bias_variance"], kp["variance"],
                                unit_diagonal=unit_diagonal)

    def K(self, param: Param, X, X2=None):
        """spherical.py:318-352: [N, N2] covariance; the diagonal cosine is exactly 1 when X2 is None."""
        return self._kernel_matrix(param, X, X2, True)

    def K2(self, param: Param, X, X2=None):
        """spherical.py:354-382: the same matrix point pair by point pair (no diagonal fix-up)."""
        return self._kernel_matrix(param, X, X2, False)


def _kappa(u, order: int):
    """spherical.py:279-316 (host numpy; only used by the init-time Funk-Hecke quadrature)."""
    if order == 0:
        return (np.pi - np.arccos(u)) / np.pi
    if order == 1:
        return (u * (np.pi - np.arccos(u)) + np.sqrt(1.0 - np.square(u))) / np.pi
    return ((1.0 + 2.0 * np.square(u)) / 3 * (np.pi - np.arccos(u)) + np.sqrt(1.0 - np.square(u))) / np.pi


@dataclasses.dataclass(frozen=True)
class ArcCosine(Spherical):
    depth: int = 1
    name: str = "ArcCosine"

    def __post_init__(self):
        if self.order not in {0, 1, 2}:
            raise ValueError("Requested order is not implemented.")

    def _shape_fn(self, param):
        return _lib.make_shape_fn(_lib.SHAPE_ARCCOSINE, depth=self.depth, order=self.order)

    def _shape_function_host(self, param, x):
        """spherical.py:438-455."""
        y = np.clip(x, -1.0, 1.0)
        for _ in range(self.depth):
            y = _kappa(y, self.order)
        return y


@dataclasses.dataclass(frozen=True)
class NTK(Spherical):
    depth: int = 1
    name: str = "NTK"

    def __post_init__(self):
        object.__setattr__(self, "order", 1)

    def _shape_fn(self, param):
        return _lib.make_shape_fn(_lib.SHAPE_NTK, depth=self.depth)

    def _shape_function_host(self, param, x):
        """spherical.py:479-500."""
        x = np.clip(x, -1.0, 1.0)
        a, y = x, x
        for _ in range(self.depth):
            a, y = _kappa(a, 1), y * _kappa(a, 0) + _kappa(a, 1)
        return y / (self.depth + 1)


@dataclasses.dataclass(frozen=True)
class PolynomialDecay(Spherical):
    truncation_level: int = 10
    name: str = "PolynomialDecay"

    def __post_init__(self):
        object.__setattr__(self, "order", 1)

    def _compute_eigvals(self, sphere_dim: int, *, max_num_eigvals=None, gegenbauer_lookup_table=None):
        """spherical.py:526-561: alpha / (n + alpha) / C_n^alpha(1)."""
        if not gegenbauer_lookup_table:
            raise ValueError("Lookup table should be provided with `PolyDecay` kernel.")
        alpha = (sphere_dim + 1 - 2.0) / 2.0
        C_1 = np.array([gegenbauer_lookup_table.at_one(n) for n in range(self.truncation_level)])
        n = np.arange(self.truncation_level, dtype=np.float64)
        return alpha / (n + alpha) / C_1

    def eigenvalues(self, param: Param, levels) -> np.ndarray:
        """spherical.py:563-587, including the clip of `levels` to truncation_level - 1 (:586)."""
        return self._eigenvalues_and_dbeta(param, levels)[0]

    def _eigenvalues_and_dbeta(self, param: Param, levels):
        n = np.arange(self.truncation_level, dtype=np.float64)
        beta = float(param.params[self.name]["beta"])
        sphere_dim = param.constants["sphere"]["sphere_dim"]
        geg = param.constants["sphere"]["gegenbauer_lookup_table"]
        decay = (1 + n) ** (-beta)
        cf = self._compute_eigvals(sphere_dim, gegenbauer_lookup_table=geg)
        S = np.sum(decay)
        eig = decay * cf / S
        ddecay = -np.log1p(n) * decay
        deig = ddecay * cf / S - decay * cf * np.sum(ddecay) / S**2
        lv = np.clip(np.asarray(levels, dtype=np.int64), 0, self.truncation_level - 1)
        return eig[lv], deig[lv]

    def _shape_coef(self, param):
        """lambda_n (n + alpha) / alpha, n < truncation_level (spherical.py:605-608)."""
        alpha = param.constants["sphere"]["alpha"]
        levels = np.arange(self.truncation_level)
        return self.eigenvalues(param, levels) * (levels.astype(np.float64) + alpha) / alpha

    def _shape_fn(self, param):
        geg = param.constants["sphere"]["gegenbauer_lookup_table"]
        return _lib.make_shape_fn(_lib.SHAPE_POLYDECAY, coef=self._shape_coef(param), grid_resolution=geg.grid_resolution)

    def _shape_function_host(self, param, x):
        """spherical.py:589-611: sum_n lambda_n (n + alpha) / alpha C_n^alpha(x) through the lookup table."""
        x = np.clip(np.asarray(x, dtype=np.float64), -1.0, 1.0)
        alpha = param.constants["sphere"]["alpha"]
        geg = param.constants["sphere"]["gegenbauer_lookup_table"]
        return sum(c * geg(n, alpha, x) for n, c in enumerate(self._shape_coef(param)))
