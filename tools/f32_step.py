"""Development aid: a few ELBO steps of the f32 (tcgen05) path at the headline shape, for ncu captures."""
import sys
import torch
sys.path.insert(0, ".")
from gpfy_b200.synthetic import headline_problem
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_818_624
prob = headline_problem(N, device="cuda", precision=sys.argv[2] if len(sys.argv) > 2 else "f32")
hp, args = prob["hot"], prob["args"]
for _ in range(3):
    out = hp.elbo_valgrad_packed(*args)
torch.cuda.synchronize()
print("ok", float(out[0]))
