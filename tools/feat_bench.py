"""Development aid: device-resident timing of Phi(X) (features), predict_diag on the headline shape."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200.synthetic import headline_problem
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
prob = headline_problem(N, device="cuda")
hp, args = prob["hot"], prob["args"]
X, Y, w, b, v, eig, vhat, lcat, mu, sg, s2 = args
def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: hp.features(X, vhat, lcat, w=w, b=b, degree_scale=torch.sqrt(eig * v), mul_r=True))
print(f"features  N={N}: {ms:.2f} ms  {N/ms*1e-3:.1f} M pts/s  write {8*hp.M*N/ms*1e-6:.0f} GB/s")
ms = timeit(lambda: hp.predict_diag(X, w, b, v, eig, vhat, lcat, mu, sg))
print(f"predict   N={N}: {ms:.2f} ms  {N/ms*1e-3:.1f} M pts/s")
