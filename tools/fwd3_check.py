"""Development aid: zonal_fwd3_kernel (barrier-free, warp = 32-point tile) against zonal_fwd2_kernel: relative difference of every block
of the packed ELBO result and the forward class's device time, several shapes."""
import sys
import torch
sys.path.insert(0, ".")
from gpfy_b200 import ops
from gpfy_b200.synthetic import headline_problem

for d, F, T, P, lik, N in ((8, 11, 30, 1, "gaussian", 4_000_000), (8, 13, 30, 1, "gaussian", 2_000_001), (32, 5, 64, 1, "bernoulli", 1_000_000),
                           (3, 9, 30, 2, "gaussian", 500_003), (8, 11, 30, 1, "gaussian", 37)):
    prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda")
    hp, args = prob["hot"], prob["args"]
    out = {}
    for off in (1, 0):
        ops.debug_option("no_fwd3", off)
        for _ in range(2):
            r = hp.elbo_valgrad_packed(*args)
        torch.cuda.synchronize()
        ops.kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = hp.elbo_valgrad_packed(*args); e1.record(); torch.cuda.synchronize()
        kt = ops.kernel_times(); ops.kernel_timing(False)
        out[off] = (hp.unpack(r.clone()), e0.elapsed_time(e1), kt["zonal_fwd"][0])
    a, b = out[1][0], out[0][0]
    rel = max(float((a[k] - b[k]).abs().max() / max(float(a[k].abs().max()), 1e-300)) for k in a if torch.is_tensor(a[k]) and a[k].numel())
    print(f"d={d} F={F} T={T} P={P} {lik} N={N} M={hp.M}: fwd2 {out[1][2]:.3f} ms (step {out[1][1]:.2f}), fwd3 {out[0][2]:.3f} ms (step {out[0][1]:.2f}), max block rel diff {rel:.2e}")
