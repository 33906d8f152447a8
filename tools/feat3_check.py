"""Development aid: features3_kernel (barrier-free, warp = (tile, degree group)) against features2_kernel on the same inputs
(max |diff|, timing), headline shape and a ragged tail."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200 import ops
from gpfy_b200.synthetic import headline_problem


def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


N0 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
for N, F, d in ((N0, 11, 8), (100_003, 11, 8), (37, 11, 8), (4_000_000, 13, 8), (2_000_001, 9, 3), (1_000_000, 6, 16)):
    prob = headline_problem(N, d=d, F=F, device="cuda")
    hp, args = prob["hot"], prob["args"]
    X, Y, w, b, v, eig, vhat, lcat, mu, sg, s2 = args
    f = lambda: hp.features(X, vhat, lcat, w=w, b=b, degree_scale=torch.sqrt(eig * v), mul_r=True)
    ops.debug_option("no_features3", 1)
    ref = f()[0].clone(); ms2 = timeit(f)
    ops.debug_option("no_features3", 0)
    new = f()[0].clone(); ms3 = timeit(f)
    dmax = (new - ref).abs().max().item()
    print(f"N={N} F={F} d={d} M={hp.M}: features2 {ms2:.3f} ms, features3 {ms3:.3f} ms ({N / ms3 * 1e-3:.1f} M pts/s), max|diff| {dmax:.3e}, max|ref| {ref.abs().max().item():.3e}, finite {bool(torch.isfinite(new).all())}")
    del ref, new
