import numpy as _np
from scipy.linalg import solve_triangular as _st


def triangular_solve(a, b, *, left_side=False, lower=False, transpose_a=False, conjugate_a=False,
                     unit_diagonal=False):
    a = _np.asarray(a)
    b = _np.asarray(b)
    if not left_side:
        # x a = b  <=>  a^T x^T = b^T
        return _st(a.T, b.T, lower=not lower, trans=1 if transpose_a else 0, unit_diagonal=unit_diagonal).T
    return _st(a, b, lower=lower, trans=1 if transpose_a else 0, unit_diagonal=unit_diagonal)
