"""gpfy_b200 — B200-native (sm_100a) implementation of GPfY's data-parallel hot path.

Spherical-harmonic features Phi(X) and the sparse variational ELBO (+ gradient, + predict_diag),
behind the gpfy Python API.  All arithmetic runs in hand-written CUDA kernels reached through the
C-ABI in include/gpfy_b200.h; there is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
