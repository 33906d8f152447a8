"""Quick device-resident timing of the ELBO+grad step (development aid; bench.py is the contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200.ops import HotPath, launch_count
from gpfy_b200.synthetic import headline_problem

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prob = headline_problem(N, T=T, device="cuda")
hp = prob["hot"]
args = prob["args"]
torch.cuda.synchronize()
for _ in range(2):
    out = hp.elbo_valgrad_packed(*args)
torch.cuda.synchronize()
import os
if os.environ.get("GPFY_PROFILE"):
    torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = launch_count()
e0.record()
for _ in range(steps):
    out = hp.elbo_valgrad_packed(*args)
e1.record(); torch.cuda.synchronize()
if os.environ.get("GPFY_PROFILE"):
    torch.cuda.cudart().cudaProfilerStop()
ms = e0.elapsed_time(e1) / steps
flops = prob["alg_flops_per_point"]
print(f"N={N} T={T} M={hp.M}: {ms:.2f} ms/step  {N/ms*1e-3:.2f} M pts/s  alg {flops*N/ms*1e-9:.2f} TF/s ({flops*N/ms*1e-9/37.1*100:.1f}% of 37.1)  launches/step={(launch_count()-l0)//steps}")
print("data_term", float(out[0]))
