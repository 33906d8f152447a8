"""A few eager minibatch steps at a small batch, for an ncu launch list of the N-independent (prologue / epilogue) kernels.
Usage: python tools/small_batch_profile.py [batch] [steps]"""
import sys

import torch

sys.path.insert(0, ".")
from gpfy_b200.device_step import DeviceTrainer  # noqa: E402
from gpfy_b200.likelihoods import Gaussian  # noqa: E402
from gpfy_b200.model import GP  # noqa: E402
from gpfy_b200.spherical import NTK  # noqa: E402
from gpfy_b200.spherical_harmonics import SphericalHarmonics  # noqa: E402
from gpfy_b200.variational import VariationalDistributionTriL  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, d = 200_000, 8
X = torch.randn((N, d), dtype=torch.float64, device="cuda")
Y = torch.sin(X[:, :1])
kernel, sh, q, lik = NTK(depth=3), SphericalHarmonics(11, phase_truncation=30), VariationalDistributionTriL(), Gaussian()
m = GP(kernel).conditional(sh, q)
param = m.init(0, input_dim=d, likelihood=lik, sh_features=sh, variational_dist=q)
tr = DeviceTrainer(param, m, q, lik, learning_rate=1e-2)
tr.fit_minibatch(X, Y, batch, steps, seed=0)
torch.cuda.synchronize()
print("ok", tr.losses()[-1])
