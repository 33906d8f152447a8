"""Funk-Hecke eigenvalues of a zonal shape function (mirror of gpfy/harmonics/utils.py:23-71).
Init-time, N-independent: 1000-node Gauss-Legendre on the host."""
import math

import numpy as np
from scipy.special import loggamma, roots_legendre

from ..gegenbauer import gegenbauer_host

legendre_x, legendre_w = roots_legendre(1000)  # utils.py:23-25


def weight_func(t, d):
    """utils.py:28-39."""
    return (1.0 - np.square(t)) ** ((d - 3) / 2)


def funk_hecke_lambda(fn, n: int, d: int) -> float:
    """utils.py:42-71: lambda_n = omega / C_n^alpha(1) * sum_i w_i f(t_i) C_n^alpha(t_i) (1 - t_i^2)^((d-3)/2)."""
    alpha = (d - 2.0) / 2.0
    C1 = gegenbauer_host(n, alpha, np.array(1.0))
    ratio = math.exp(loggamma(d / 2) - loggamma((d - 1) / 2)) / math.sqrt(math.pi)
    integrand = gegenbauer_host(n, alpha, legendre_x) * fn(legendre_x) * weight_func(legendre_x, d)
    return float(ratio / C1 * np.dot(legendre_w, integrand))
