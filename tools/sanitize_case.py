"""Small end-to-end pass over every C-ABI entry point (ragged sizes) for compute-sanitizer."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200.synthetic import headline_problem
from gpfy_b200.ops import gegenbauer
for (N, T, P, lik, d, F) in [(1000, 30, 1, "gaussian", 8, 11), (333, 12, 2, "bernoulli", 5, 5), (65, 30, 1, "gaussian", 8, 11), (257, 8, 1, "gaussian", 2, 6),
                             (700, 30, 1, "gaussian", 8, 13), (515, 64, 1, "bernoulli", 32, 5), (300, 64, 1, "gaussian", 8, 11)]:   # int8 contraction on the zonal_bwd4 paths
    prob = headline_problem(N, d=d, F=F, T=T, P=P, likelihood=lik, device="cuda")
    hp, args = prob["hot"], prob["args"]
    X, Y, w, b, v, eig, vhat, lcat, mu, sg, s2 = args
    out = hp.elbo_valgrad_packed(*args)
    hp.predict_diag(X, w, b, v, eig, vhat, lcat, mu, sg)
    A, r = hp.features(X, vhat, lcat, w=w, b=b, degree_scale=torch.sqrt(eig * v), mul_r=True)
    hp.project_q(A, mu, sg)
    hp.to_sphere(X, w, b)
    hp.prior_kl_valgrad(mu, sg)
    hp.elbo_valgrad_host(X.cpu().pin_memory(), Y.cpu().pin_memory(), *args[2:])
    gegenbauer(7, 3.5, X[:, 0] / 10, with_grad=True)
    torch.cuda.synchronize()
    print("ok", N, T, P, lik, float(out[0]))
