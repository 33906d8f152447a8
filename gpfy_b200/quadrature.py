"""20-node Gauss-Hermite rule (mirror of gpfy/quadrature.py:26-27).  The device kernels carry the same
nodes baked in (csrc/gh20.h); this host copy exists for API parity and for the tests."""
import numpy as np

gh_points, gh_weights = np.polynomial.hermite.hermgauss(20)  # quadrature.py:26-27
