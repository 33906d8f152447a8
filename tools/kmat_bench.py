"""Times the kernel-matrix side on one GPU: K(X) of the three kernels and GP.predict's covariance.

Reports elements/s and the output-write bandwidth; the shape function (acos / sqrt chains, or 2 x truncation_level
recurrence steps for PolynomialDecay) makes K compute-bound, the covariance is two DMMA GEMMs with K = M.
Usage: python tools/kmat_bench.py [N]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gpfy_b200.likelihoods import Gaussian  # noqa: E402
from gpfy_b200.model import GP  # noqa: E402
from gpfy_b200.spherical import NTK, ArcCosine, PolynomialDecay  # noqa: E402
from gpfy_b200.spherical_harmonics import SphericalHarmonics  # noqa: E402
from gpfy_b200.variational import VariationalDistributionTriL  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    d = 8
    rng = np.random.default_rng(0)
    X = torch.as_tensor(rng.standard_normal((N, d)), device="cuda")
    out = {"N": N, "d": d}
    for kernel in (NTK(depth=1), NTK(depth=3), ArcCosine(depth=1, order=1), PolynomialDecay(truncation_level=11)):
        param = kernel.init(0, d)
        ms = timed(lambda: kernel.K(param, X))
        tag = f"{kernel.name}" + (f"_depth{kernel.depth}" if hasattr(kernel, "depth") else f"_T{kernel.truncation_level}")
        out[tag] = {"ms": ms, "Gelem_per_s": N * N / ms / 1e6, "write_GBps": N * N * 8 / ms / 1e6}
    kernel = NTK(depth=3)
    sh = SphericalHarmonics(11, phase_truncation=30)
    q = VariationalDistributionTriL()
    m = GP(kernel).conditional(sh, q)
    param = m.init(0, input_dim=d, likelihood=Gaussian(), sh_features=sh, variational_dist=q)
    M = int(sh.num_inducing(param))
    Np = min(N, 4096)
    Xp = X[:Np]
    ms = timed(lambda: m.predict(param, Xp))
    flops = 2.0 * M * M * Np + 2 * 2.0 * M * Np * Np    # L^T A, tmp^T tmp, A^T A
    out["predict_cov"] = {"N": Np, "M": M, "ms": ms, "gemm_TFLOPs": flops / ms / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
