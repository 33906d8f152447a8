"""Rebuilds the gpfy_b200 API objects (kernel, sh, q, lik, model, Param) of a golden case."""
import numpy as np

from gpfy_b200.likelihoods import Bernoulli, Gaussian
from gpfy_b200.model import GP
from gpfy_b200.spherical import NTK, ArcCosine, PolynomialDecay
from gpfy_b200.spherical_harmonics import SphericalHarmonics
from gpfy_b200.variational import VariationalDistributionFullCovariance, VariationalDistributionTriL


def objects_from_golden(g):
    F, T, d, P = int(g["meta_F"]), int(g["meta_T"]), int(g["meta_input_dim"]), int(g["meta_P"])
    kname = str(g["meta_kernel"])
    ard = bool(g["meta_ard"]) if "meta_ard" in g else True
    if kname == "PolynomialDecay":
        kernel = PolynomialDecay(truncation_level=int(g["meta_truncation_level"]), ard=ard)
    elif kname == "NTK":
        kernel = NTK(depth=int(g["meta_depth"]), ard=ard)
    else:
        kernel = ArcCosine(depth=int(g["meta_depth"]), order=int(g["meta_order"]), ard=ard)
    lik = Gaussian() if str(g["meta_lik"]) == "Gaussian" else Bernoulli()
    q = VariationalDistributionTriL() if str(g["meta_q"]).endswith("TriL") else VariationalDistributionFullCovariance()
    sh = SphericalHarmonics(F, phase_truncation=T)
    m = GP(kernel).conditional(sh, q)
    proj = int(g["meta_projection_dim"]) or None
    param = m.init(0, input_dim=d, num_independent_processes=P, projection_dim=proj, likelihood=lik, sh_features=sh, variational_dist=q)
    p = {k: (dict(v) if isinstance(v, dict) else v) for k, v in param.params.items()}
    kp = dict(p[kernel.name])
    for k in list(kp):
        kp[k] = np.array(g[f"p_{k}"])
    p[kernel.name] = kp
    var = {k: (dict(v) if isinstance(v, dict) else v) for k, v in p["variational"].items()}
    var["inducing_features"] = {f"V_{n:02d}": np.array(g[f"p_V_{n:02d}"]) for n in range(F)}
    var[q.name] = {"mu": np.array(g["p_mu"]), "Sigma": np.array(g["p_Sigma"])}
    p["variational"] = var
    if "p_sigma2" in g:
        p["likelihood"] = {lik.name: {"variance": np.array(g["p_sigma2"])}}
    consts = dict(param.constants)
    consts["variational"] = {"inducing_features": {"orthogonal_basis": [np.array(g[f"c_orth_{n:02d}"]) for n in range(F)]}}
    if "const_eigenvalues" in g:
        consts[kernel.name] = {"eigenvalues": np.array(g["const_eigenvalues"])}
    param = param.replace(params=p, constants=consts)
    return dict(kernel=kernel, lik=lik, q=q, sh=sh, m=m, param=param, init_param=m.init(0, input_dim=d, num_independent_processes=P, projection_dim=proj,
                                                                                       likelihood=lik, sh_features=sh, variational_dist=q))
