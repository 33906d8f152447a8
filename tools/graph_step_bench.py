"""Launch-bound regime: per-step time of minibatch training, eager C-ABI calls vs one captured CUDA graph replayed.
Usage: python tools/graph_step_bench.py [batch] [steps]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gpfy_b200.device_step import DeviceTrainer  # noqa: E402
from gpfy_b200.likelihoods import Gaussian  # noqa: E402
from gpfy_b200.model import GP  # noqa: E402
from gpfy_b200.spherical import NTK  # noqa: E402
from gpfy_b200.spherical_harmonics import SphericalHarmonics  # noqa: E402
from gpfy_b200.variational import VariationalDistributionTriL  # noqa: E402


def main():
    batch = int(float(sys.argv[1])) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    N, d = 2_000_000, 8
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    X = torch.randn((N, d), dtype=torch.float64, device="cuda", generator=g)
    Y = torch.sin(X[:, :1]) + 0.1 * torch.randn((N, 1), dtype=torch.float64, device="cuda", generator=g)
    kernel, sh, q, lik = NTK(depth=3), SphericalHarmonics(11, phase_truncation=30), VariationalDistributionTriL(), Gaussian()
    m = GP(kernel).conditional(sh, q)
    param = m.init(0, input_dim=d, likelihood=lik, sh_features=sh, variational_dist=q)
    out = {"batch": batch, "steps": steps, "N": N, "M": int(sh.num_inducing(param))}
    for tag, use_graph in (("eager", False), ("graph", True)):
        tr = DeviceTrainer(param, m, q, lik, learning_rate=1e-2)
        tr.fit_minibatch(X, Y, batch, 5, seed=0, use_graph=use_graph)       # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.fit_minibatch(X, Y, batch, steps, seed=1, use_graph=use_graph)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ls = tr.losses()
        out[tag] = {"ms_per_step": dt / steps * 1e3, "rows_per_s": batch * steps / dt, "loss_first": float(ls[5]), "loss_last": float(ls[-1])}
    out["speedup"] = out["eager"]["ms_per_step"] / out["graph"]["ms_per_step"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
