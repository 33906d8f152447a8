"""Development aid: per-kernel-class device time of one ELBO step for a given shape."""
import sys
import torch
sys.path.insert(0, ".")
from gpfy_b200 import ops
from gpfy_b200.synthetic import headline_problem
d, F, T, P, lik, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(float(sys.argv[6]))
for opt in sys.argv[7:]:          # development switches, e.g. no_fused_bwd=1 chunk_mult=8
    k, v = opt.split("=")
    ops.debug_option(k, int(v))
prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda")
hp, args = prob["hot"], prob["args"]
for _ in range(2):
    hp.elbo_valgrad_packed(*args)
torch.cuda.synchronize()
ops.kernel_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); hp.elbo_valgrad_packed(*args); e1.record(); torch.cuda.synchronize()
kt = ops.kernel_times()
ops.kernel_timing(False)
print(f"d={d} F={F} T={T} P={P} {lik} N={N} M={hp.M} {' '.join(sys.argv[7:])}: {e0.elapsed_time(e1):.2f} ms", {k: round(v[0], 2) for k, v in kt.items()})
