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
import os
prec = os.environ.get("GPFY_BENCH_PRECISION", "f64")   # tools only: the library itself reads no environment variable
prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda", precision=prec)
hp, args = prob["hot"], prob["args"]
for _ in range(2):
    hp.elbo_valgrad_packed(*args)
torch.cuda.synchronize()
ops.kernel_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); hp.elbo_valgrad_packed(*args); e1.record(); torch.cuda.synchronize()
kt = ops.kernel_times()
ops.kernel_timing(False)
extra = ""
if prec != "f64":   # accuracy against the f64 path on the same rows
    p64 = headline_problem(min(N, 200000), d=d, F=F, T=T, P=P, likelihood=lik, device="cuda")
    S = min(N, 200000)
    a64 = p64["hot"].unpack(p64["hot"].elbo_valgrad_packed(*p64["args"]))
    a32 = hp.unpack(hp.elbo_valgrad_packed(args[0][:S], args[1][:S], *args[2:]))
    extra = " max rel err vs f64: " + ", ".join(f"{k} {float((a32[k]-a64[k]).abs().max()/a64[k].abs().max()):.1e}" for k in ("sigma", "eig", "lcat", "mu", "vhat") if float(a64[k].abs().max()) > 0)
print(f"d={d} F={F} T={T} P={P} {lik} N={N} M={hp.M} {prec} {' '.join(sys.argv[7:])}: {e0.elapsed_time(e1):.2f} ms", {k: round(v[0], 2) for k, v in kt.items()}, extra)
