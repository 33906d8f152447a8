"""Other BASELINE.json shapes at scale (device-resident, 1 GPU): T sweep, degree 12, d = 32 Bernoulli.
Prints one JSON object per case; each case is also checked against the oracle on a 2001-row sample, which is why this
script lives under tests/ (only tests/, smoke() and bench.py's CPU baseline may use oracle/).
Run on a GPU box:  python tests/sweep_configs.py   ->  gpurun_out/sweep.json"""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200.synthetic import headline_problem
from oracle import gpfy_oracle as O

CASES = [
    dict(name="cfg3 d=8 F=11 T=30 gaussian", d=8, F=11, T=30, lik="gaussian", N=2_000_000),
    dict(name="cfg3 d=8 F=11 T=64 gaussian", d=8, F=11, T=64, lik="gaussian", N=1_000_000),
    dict(name="cfg3 d=8 F=11 T=100 gaussian", d=8, F=11, T=100, lik="gaussian", N=500_000),
    dict(name="cfg4 d=8 F=13 T=30 gaussian", d=8, F=13, T=30, lik="gaussian", N=2_000_000),
    dict(name="cfg5 d=32 F=5 T=64 bernoulli", d=32, F=5, T=64, lik="bernoulli", N=1_000_000),
    dict(name="cfg1 d=2 F=11 full (S^2) gaussian", d=2, F=11, T=2**31 - 1, lik="gaussian", N=2_000_000),
]
out = []
for c in CASES:
    prob = headline_problem(c["N"], d=c["d"], F=c["F"], T=c["T"], likelihood=c["lik"], device="cuda")
    hp, args, host = prob["hot"], prob["args"], prob["host"]
    f = lambda: hp.elbo_valgrad_packed(*args)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    S = 2001
    sub = hp.unpack(hp.elbo_valgrad_packed(args[0][:S].contiguous(), args[1][:S].contiguous(), *args[2:]))
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": 1, "mu": host["mu"], "Lq": host["Lq"], "sigma2": host["sigma2"], "lik": c["lik"]}
    oval, og = O.elbo_and_grads(pack, args[0][:S].cpu().numpy(), args[1][:S].cpu().numpy(), data_only=True)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    F = len(host["V"])
    errs = {"value": abs(float(sub["data_term"]) - oval) / abs(oval), "mu": rel(sub["mu"].cpu().numpy(), og["mu"]),
            "Lq": rel(sub["sigma"].cpu().numpy(), og["Lq"]),
            "vhat": rel(sub["vhat"].cpu().numpy(), np.concatenate([og[f"V{i}"] for i in range(F)], 0)), "w": rel(sub["w"].cpu().numpy(), og["w"])}
    flops = prob["alg_flops_per_point"]
    r = dict(case=c["name"], N=c["N"], M=prob["M"], ms_per_step=ms, points_per_s=c["N"] / ms * 1e3, alg_flops_per_point=flops,
             alg_tflops=flops * c["N"] / ms * 1e-9, frac_of_fp64_peak=flops * c["N"] / ms * 1e-9 / 37.141, max_rel_err_vs_oracle=max(errs.values()), errs=errs)
    out.append(r)
    print(json.dumps(r), flush=True)
    del prob, hp, args
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
