"""Times the shuffled-minibatch gather on one GPU: GB/s of rows moved (read + write) for a
device-resident data set, against the HBM copy peak.  Usage: python tools/minibatch_bench.py [N] [batch]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from gpfy_b200.minibatch import DeviceMinibatches  # noqa: E402


def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    batch = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
    d, P = 8, 1
    X = torch.randn((N, d), dtype=torch.float64, device="cuda")
    Y = torch.randn((N, P), dtype=torch.float64, device="cuda")
    mb = DeviceMinibatches(X, Y, batch, seed=0)
    for _ in range(3):
        mb.next()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        mb.next()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    moved = 2 * batch * (d + P) * 8
    print(json.dumps({"N": N, "batch": batch, "d": d, "ms": ms, "rows_per_s": batch / ms * 1e3, "GBps_read_plus_write": moved / ms / 1e6}))


if __name__ == "__main__":
    main()
