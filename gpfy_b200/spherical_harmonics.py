"""Sparse spherical-harmonic inducing features (mirror of gpfy/spherical_harmonics.py).

Same dataclass, fields, parameter-tree keys and NaN-sentinel semantics as the reference.  What depends
on the N data points — `polynomial_expansion`, `Kuf`, the projection — runs in `gpfy_sh_features_fwd`;
what is N-independent (normalising V, the Gram matrices and their Cholesky factors, Kuu) is O(sum m_l^3)
numpy on the host, exactly where the reference keeps it in plain JAX."""
import dataclasses
from typing import List, Optional

import numpy as np
import torch
from scipy.special import comb

from . import _runtime
from .gegenbauer import GegenbauerLookupTable, gegenbauer_host
from .harmonics.fund_set import fundamental_set_loader
from .param import Param, identity


def _num_harmonics(dim: int, frequency: int) -> int:
    """spherical_harmonics.py:31-47."""
    if frequency == 0:
        return 1
    return int(comb(frequency + dim - 3, frequency - 1) + comb(frequency + dim - 2, frequency))


@dataclasses.dataclass(frozen=True)
class SphericalHarmonics:
    num_frequencies: int = None  # the reference's default_factory=10 makes the no-argument call raise too (:75)
    phase_truncation: int = 2**31 - 1

    def __post_init__(self):
        if self.num_frequencies is None:
            raise TypeError("SphericalHarmonics needs `num_frequencies`")

    @property
    def levels(self):
        return np.arange(self.num_frequencies, dtype=np.int32)

    # ---- spherical_harmonics.py:82-192 ----
    def init(self, key, input_dim: int, param: Optional[Param] = None) -> Param:
        sphere_dim = input_dim
        dim = sphere_dim + 1
        sphere = dict(param.constants.get("sphere") or {}) if param else {}
        if sphere and sphere_dim != sphere["sphere_dim"]:
            raise ValueError("`param` contains a sphere that is not compatible to `input_dim`.")
        try:
            fund_set = fundamental_set_loader(dim)
        except ValueError:
            fund_set = None
        rng = _runtime.as_rng(key)
        Vs, trainables = {}, {}
        for n in self.levels:
            n = int(n)
            num_phase = _num_harmonics(dim, n)
            if (num_phase <= self.phase_truncation) and fund_set:
                Vs[f"V_{n:02d}"] = np.array(fund_set(n), dtype=np.float64)  # ValueError for a missing degree, like :120
                trainables[f"V_{n:02d}"] = False
            else:
                Vs[f"V_{n:02d}"] = rng.standard_normal((min(self.phase_truncation, num_phase), dim))
                trainables[f"V_{n:02d}"] = True
        orth_basis = [np.full((v.shape[0], v.shape[0]), np.nan) for _, v in sorted(Vs.items())]
        bijectors = {k: identity() for k in Vs}
        collection = "variational"
        if (not sphere) or ("gegenbauer_lookup_table" not in sphere) or (sphere["gegenbauer_lookup_table"].max_level < self.num_frequencies):
            alpha = (dim - 2.0) / 2.0
            sphere["sphere_dim"], sphere["alpha"] = sphere_dim, alpha
            sphere["gegenbauer_lookup_table"] = GegenbauerLookupTable(self.num_frequencies, alpha)
        all_params, all_trainables, all_bijectors, all_constants, constrained = {}, {}, {}, {}, True
        if param:
            all_params, all_bijectors = dict(param.params), dict(param._bijectors)
            all_trainables, all_constants, constrained = dict(param._trainables), dict(param.constants), param._constrained
        all_params[collection] = {"inducing_features": Vs}
        all_bijectors[collection] = {"inducing_features": bijectors}
        all_trainables[collection] = {"inducing_features": trainables}
        all_constants[collection] = {"inducing_features": {"orthogonal_basis": orth_basis}}
        all_constants["sphere"] = sphere
        param = Param(params=all_params, _trainables=all_trainables, _bijectors=all_bijectors, constants=all_constants,
                      _constrained=constrained)
        if not any(trainables.values()):
            orth_basis = self.orthogonalise_basis(param)
        all_constants = dict(all_constants)
        all_constants[collection] = {"inducing_features": {"orthogonal_basis": orth_basis}}
        return param.replace(constants=all_constants)

    # ---- spherical_harmonics.py:194-288 ----
    @staticmethod
    def _learned(param: Param) -> bool:
        ob = param.constants["variational"]["inducing_features"]["orthogonal_basis"]
        return bool(np.isnan(np.sum(ob[0])))

    def Vs(self, param: Param) -> List[np.ndarray]:
        V = [v for _, v in sorted(param.params["variational"]["inducing_features"].items())]
        if self._learned(param):
            return [v / np.sqrt(np.sum(np.square(v), axis=1, keepdims=True)) for v in V]
        return V

    def Ls(self, param: Param) -> List[np.ndarray]:
        if self._learned(param):
            return self.orthogonalise_basis(param)
        return list(param.constants["variational"]["inducing_features"]["orthogonal_basis"])

    def num_phase_in_frequency(self, param: Param) -> List[int]:
        return [int(v.shape[0]) for v in self.Vs(param)]

    def num_inducing(self, param: Param) -> int:
        return sum(self.num_phase_in_frequency(param))

    def orthogonalise_basis(self, param: Param) -> List[np.ndarray]:
        alpha = param.constants["sphere"]["alpha"]
        out = []
        for n, v in zip(self.levels, self.Vs(param)):
            c = alpha / (alpha + float(n))
            B = c * gegenbauer_host(int(n), alpha, v @ v.T)
            out.append(np.linalg.cholesky(B + 1e-16 * np.eye(B.shape[0])))
        return out

    # ---- device side ----
    def _cat(self, param: Param):
        """(vhat [M, dim], lcat [sum m_l^2], m) as the kernels consume them.  In learned mode (phase truncation) the
        normalisation, the Gram matrices and their Cholesky factors are rebuilt on the device every call
        (`gpfy_inducing_basis_fwd_ws`) when every m_l <= 256; otherwise it is host numpy, as in fixed mode where the
        factors are constants."""
        raw = [v for _, v in sorted(param.params["variational"]["inducing_features"].items())]
        m = [int(v.shape[0]) for v in raw]
        if self._learned(param) and max(m) <= 256 and _runtime.cuda_available():
            hp = _runtime.hot_path(raw[0].shape[1] - 1, m)
            vhat, lcat = hp.inducing_basis(np.concatenate(raw, 0), normalise=True)
            return vhat, lcat, m
        Vs, Ls = self.Vs(param), self.Ls(param)
        return np.concatenate(Vs, 0), np.concatenate([l.ravel() for l in Ls]), m

    def polynomial_expansion(self, param: Param, X):
        """[M, N] harmonics of points X [N, dim] that already lie on the sphere (spherical_harmonics.py:290-320)."""
        vhat, lcat, m = self._cat(param)
        dim = int(vhat.shape[1])
        if int(np.shape(X)[-1]) != dim:
            raise ValueError(f"X has {np.shape(X)[-1]} columns, the harmonics live in dimension {dim}")
        hp = _runtime.hot_path(dim - 1, m)
        return hp.features(X, vhat, lcat, apply_to_sphere=False)[0]

    def Kuu(self, param: Param, kernel) -> np.ndarray:
        """diag: 1 / lambda_l repeated m_l times (spherical_harmonics.py:322-342)."""
        eigs = kernel.eigenvalues(param, self.levels)
        reps = self.num_phase_in_frequency(param)
        return np.concatenate([np.ones(r) / e for e, r in zip(eigs, reps)])

    def _features(self, param: Param, kernel, x, degree_scale):
        vhat, lcat, m = self._cat(param)
        kp = param.params[kernel.name]
        hp = _runtime.hot_path(int(np.shape(x)[-1]), m, weight_mode=kernel._weight_mode(param), kernel_order=kernel.order,
                               projection_dim=kernel._projection_dim(param))
        return hp.features(x, vhat, lcat, w=kp["weight_variances"], b=kp["bias_variance"], degree_scale=degree_scale, mul_r=True)[0]

    def Kuf(self, param: Param, kernel, x):
        """Phi * r * sqrt(variance) -> [M, N] (spherical_harmonics.py:344-362)."""
        v = float(param.params[kernel.name]["variance"])
        return self._features(param, kernel, x, np.full(self.num_frequencies, np.sqrt(v)))

    def conditional_fun(self, kernel):
        """(project_fun, conditional_cov_fun, conditional_var_fun) in the reference's order (:364-395)."""
        def project_fun(param, x):
            v = float(param.params[kernel.name]["variance"])
            return self._features(param, kernel, x, np.sqrt(kernel.eigenvalues(param, self.levels) * v))

        def conditional_var_fun(param, x, proj):
            return kernel.K_diag(param, x) - (proj * proj).sum(-2)

        def conditional_cov_fun(param, x, proj):
            # K(x) - proj^T proj (:391-393) as one accumulate pass: base = K, S = 0, minus the identity
            M = proj.shape[0]
            hp = _runtime.hot_path(1, [M], num_outputs=1, q_kind="fullcov")
            zero = torch.zeros((1, M, M), dtype=torch.float64, device=hp.device)
            return hp.project_variance(proj, zero, subtract_identity=True, base=kernel.K(param, x))[0]

        return (project_fun, conditional_cov_fun, conditional_var_fun)
