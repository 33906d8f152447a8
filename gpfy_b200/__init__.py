"""gpfy_b200 — B200-native (sm_100a) implementation of GPfY's data-parallel hot path.

Spherical-harmonic features Phi(X) and the sparse variational ELBO (+ gradient, + predict_diag),
behind the gpfy Python API.  All N-dependent arithmetic runs in hand-written CUDA kernels reached
through the C-ABI in include/gpfy_b200.h; there is no CPU / PyTorch fallback.

    from gpfy_b200.spherical import NTK, ArcCosine, PolynomialDecay
    from gpfy_b200.spherical_harmonics import SphericalHarmonics
    from gpfy_b200.variational import VariationalDistributionTriL, VariationalDistributionFullCovariance
    from gpfy_b200.likelihoods import Gaussian, Bernoulli
    from gpfy_b200.model import GP
    from gpfy_b200.optimization import elbo, elbo_value_and_grad, create_training_step
"""
__version__ = "0.1.0"
