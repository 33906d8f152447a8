"""Train a sparse spherical-harmonic GP with the gpfy API on synthetic data (the flow of the reference's
docs/notebooks/example.py: kernel + SphericalHarmonics + q + likelihood -> GP.conditional -> fit -> predict_diag),
every N-dependent operation running in the CUDA kernels of libgpfy_b200.so.

    python examples/train_synthetic.py [N] [iterations]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gpfy_b200.likelihoods import Gaussian
from gpfy_b200.model import GP
from gpfy_b200.optimization import create_training_step
from gpfy_b200.spherical import NTK
from gpfy_b200.spherical_harmonics import SphericalHarmonics
from gpfy_b200.training import adam
from gpfy_b200.variational import VariationalDistributionTriL

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d = 6
g = torch.Generator(device="cuda")
g.manual_seed(0)
X = torch.randn((N, d), dtype=torch.float64, device="cuda", generator=g)
f = torch.sin(2 * X[:, :1]) * torch.cos(X[:, 1:2]) + 0.3 * X[:, 2:3]
Y = f + 0.1 * torch.randn((N, 1), dtype=torch.float64, device="cuda", generator=g)
Xt = torch.randn((5000, d), dtype=torch.float64, device="cuda", generator=g)
ft = torch.sin(2 * Xt[:, :1]) * torch.cos(Xt[:, 1:2]) + 0.3 * Xt[:, 2:3]

kernel = NTK(depth=5)                                   # docs/notebooks/example.py:22
sh = SphericalHarmonics(num_frequencies=7, phase_truncation=30)
lik, q = Gaussian(), VariationalDistributionTriL()
m = GP(kernel).conditional(sh, q)
param = m.init(0, input_dim=d, likelihood=lik, sh_features=sh, variational_dist=q)
print(f"N = {N}, M = {sh.num_inducing(param)} inducing harmonics, trainable V: "
      f"{[k for k, t in param._trainables['variational']['inducing_features'].items() if t]}")

step = create_training_step(m, {"x": X, "y": Y}, ("x", "y"), q, lik)
t0 = time.perf_counter()
param, state, losses = m.fit(param, step, adam(5e-2), iters, progress_bar=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mu, var = m.predict_diag(param, Xt)
rmse = float(torch.sqrt(torch.mean((mu - ft) ** 2)))
print(f"-ELBO {losses[0]:.1f} -> {losses[-1]:.1f} in {iters} Adam steps ({dt / iters * 1e3:.1f} ms/step, {N * iters / dt * 1e-6:.1f} M points/s)")
print(f"test RMSE {rmse:.4f}, mean predictive sd {float(var.sqrt().mean()):.4f}, noise sd {float(np.sqrt(param.params['likelihood']['Gaussian']['variance'])):.4f}")

# ---- the same model trained without the host in the loop: shuffled minibatches gathered on the device, one captured
# CUDA graph replayed per step (gpfy_b200.device_step.DeviceTrainer), then the full predictive covariance of a few rows ----
from gpfy_b200.device_step import DeviceTrainer

param0 = m.init(0, input_dim=d, likelihood=lik, sh_features=sh, variational_dist=q)
trainer = DeviceTrainer(param0, m, q, lik, learning_rate=5e-2)
batch = min(N, 65_536)
DeviceTrainer(param0, m, q, lik).fit_minibatch(X, Y, batch, 2, use_graph=True)   # warm-up: scratch sizes, first capture
torch.cuda.synchronize()
t0 = time.perf_counter()
trainer.fit_minibatch(X, Y, batch, iters, seed=0, use_graph=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mb_param = trainer.param()
mu_mb, _ = m.predict_diag(mb_param, Xt)
mean, cov = m.predict(mb_param, Xt[:500])                       # [500, 1], [1, 500, 500]
ls = trainer.losses()
print(f"minibatch (batch {batch}, graph replay): -ELBO estimate {ls[0]:.1f} -> {ls[-1]:.1f}, {dt / iters * 1e3:.2f} ms/step, "
      f"test RMSE {float(torch.sqrt(torch.mean((mu_mb - ft) ** 2))):.4f}, "
      f"covariance min eigenvalue {float(torch.linalg.eigvalsh(cov[0]).min()):.2e}")
