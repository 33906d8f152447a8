"""Development aid: time of one API-level step (gpfy_b200.optimization.loss_and_grad_unconstrained on a Param) vs the
packed C-ABI call, headline shape, device-resident rows."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200.likelihoods import Gaussian
from gpfy_b200.model import GP
from gpfy_b200.spherical import PolynomialDecay
from gpfy_b200.spherical_harmonics import SphericalHarmonics
from gpfy_b200.variational import VariationalDistributionTriL
from gpfy_b200 import optimization as opt

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
kernel, lik, q, sh = PolynomialDecay(), Gaussian(), VariationalDistributionTriL(), SphericalHarmonics(11, phase_truncation=30)
m = GP(kernel).conditional(sh, q)
param = m.init(0, input_dim=8, likelihood=lik, sh_features=sh, variational_dist=q)
g = torch.Generator(device="cuda"); g.manual_seed(0)
X = torch.randn((N, 8), dtype=torch.float64, device="cuda", generator=g)
Y = torch.sin(X).sum(1, keepdim=True) / 8 ** 0.5 + 0.1 * torch.randn((N, 1), dtype=torch.float64, device="cuda", generator=g)
free = param.unconstrained()
for _ in range(2):
    loss, grads = opt.loss_and_grad_unconstrained(free, m, q, lik, (X, Y))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    loss, grads = opt.loss_and_grad_unconstrained(free, m, q, lik, (X, Y))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
hp, args = m.hot_args(param.constrained() if not param._constrained else param, lik)
s2 = lik._sigma2(param)
dargs = [hp._dev(a) for a in args] + [hp._dev(s2)]
hp.elbo_valgrad_packed(X, Y, *dargs); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    hp.elbo_valgrad_packed(X, Y, *dargs)
torch.cuda.synchronize()
dk = (time.perf_counter() - t0) / 3
print(f"N={N}: API step {dt*1e3:.2f} ms ({N/dt*1e-6:.1f} M pts/s), packed C-ABI call {dk*1e3:.2f} ms; host-side share {100*(dt-dk)/dt:.1f}%  loss {loss:.6f}")

from gpfy_b200.device_step import DeviceTrainer
tr = DeviceTrainer(param, m, q, lik, learning_rate=1e-2)
for _ in range(2):
    tr.step(X, Y)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    tr.step(X, Y)
torch.cuda.synchronize()
dd = (time.perf_counter() - t0) / 5
print(f"N={N}: DeviceTrainer step {dd*1e3:.2f} ms ({N/dd*1e-6:.1f} M pts/s) vs packed call {dk*1e3:.2f} ms: driver overhead {100*(dd-dk)/dd:.1f}%; loss {float(tr.loss_trace[-1]):.4f}")
