#!/usr/bin/env python
"""bench.py — ELBO+grad points/sec of the SVGP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): full SVGP ELBO + gradient step, input_dim d = 8 (ambient dim 9,
alpha = 3.5), degrees 0..10 (num_frequencies = 11), phase truncation T = 30 (M = 280 inducing
harmonics, degrees >= 2 trainable), TriL q, Gaussian likelihood, P = 1, float64, N = 10 M synthetic
rows in total, sharded by rows over the ranks (strong scaling), one all-reduce of the packed buffer.

A step = one ELBO value + all gradients over every row of the data set (+ the KL term and the
all-reduce).  `value` has the rows resident in HBM; `e2e` streams them from pinned host memory through
`HotPath.elbo_valgrad_host` and reads the packed result back, every step.

`--impl reference` times the CPU restatement of the reference (oracle/gpfy_oracle.py, torch float64 on
the host cores with autograd — JAX is not installable in this image) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ELBO+grad points/sec at N=10M d=8 deg10"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=float, default=10_000_000, help="total rows over all ranks")
    ap.add_argument("--T", type=int, default=30, help="phase truncation")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="rows per CPU-baseline step")
    return ap.parse_args()


def workload_config(N, T, M, world):
    return {
        "workload": "svgp_elbo_valgrad", "N_total": int(N), "input_dim": 8, "ambient_dim": 9, "num_frequencies": 11,
        "max_degree": 10, "phase_truncation": int(T), "num_inducing": int(M), "num_outputs": 1, "likelihood": "gaussian",
        "q": "tril", "kernel": "PolynomialDecay(truncation_level=10)", "parallelism": f"rows/{world}",
        "l2": "per-step inputs (>= 80 MB/rank) plus the per-chunk panels (Z, C, dT: 6.3 GB per 0.9 M-row chunk) exceed the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------
# clocks sampled during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference (oracle) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(host, X_sample, Y_sample, steps, warmup):
    """points/s of oracle ELBO+grad (torch float64 CPU, autograd) on `steps` passes over the sample."""
    import torch
    from oracle import gpfy_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    pack = {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": host["order"], "mu": host["mu"], "Lq": host["Lq"],
            "sigma2": host["sigma2"], "lik": host["lik"]}
    for _ in range(warmup):
        O.elbo_and_grads(pack, X_sample, Y_sample)
    t0 = time.perf_counter()
    for _ in range(steps):
        val, _ = O.elbo_and_grads(pack, X_sample, Y_sample)
    dt = time.perf_counter() - t0
    return X_sample.shape[0] * steps / dt, dt / steps, val


def run_reference(a):
    """`--impl reference`: the oracle port on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from gpfy_b200 import synthetic
    S = int(a.cpu_sample)
    host, X, Y, m = synthetic.headline_host_problem(S, T=a.T)
    rate, sec, _ = cpu_reference_rate(host, X, Y, max(1, a.steps), max(1, a.warmup))
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a.n, a.T, sum(m), a.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{S} rows of the workload per step, torch {torch.__version__} float64 CPU + autograd "
                                   f"(oracle/gpfy_oracle.py; JAX is not installable here)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gpfy_b200 import distributed as gd
    from gpfy_b200 import ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (gpfy_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group(backend="nccl", device_id=dev)
    N = int(a.n)
    prob = synthetic.headline_problem(N, T=a.T, device=dev, rank=rank, world=world)
    hp, args = prob["hot"], prob["args"]
    X, Y = args[0], args[1]
    n_local = X.shape[0]
    mu_d, sig_d = args[8], args[9]
    allreduce = gd.PackedAllReduce()
    packed = torch.empty((hp.layout.total,), dtype=torch.float64, device=dev)

    def step():
        hp.elbo_valgrad_packed(*args, out=packed)
        allreduce(packed)
        return hp.prior_kl_valgrad(mu_d, sig_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(3, a.warmup)):
        step()
    barrier()

    # ---- timed region: K steps, device time, max over ranks ----
    # rank 0 samples its own GPU (eight concurrent nvidia-smi loops stall the driver and with it every rank's launches)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ops.kernel_timing(True)
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        kl = step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / a.steps
    launches = (ops.launch_count() - l0) // a.steps
    ktimes = ops.kernel_times()
    ops.kernel_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    value = N / (ms * 1e-3)
    data_term = float(packed[0].item())

    # ---- end to end: rows in pinned host memory, packed result read back, every step ----
    e2e = None
    if not a.no_e2e:
        Xh = torch.empty(X.shape, dtype=torch.float64).pin_memory()
        Yh = torch.empty(Y.shape, dtype=torch.float64).pin_memory()
        Xh.copy_(X); Yh.copy_(Y)
        out_h = torch.empty((hp.layout.total,), dtype=torch.float64).pin_memory()
        pargs = args[2:]
        reduce = allreduce if world > 1 else None

        def step_e2e():
            hp.elbo_valgrad_host(Xh, Yh, *pargs, out_host=out_h, reduce=reduce)
            return float(out_h[0])

        for _ in range(2):
            v_e2e = step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            v_e2e = step_e2e()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)) / a.steps
        if abs(v_e2e - data_term) > 1e-9 * abs(data_term):
            raise SystemExit(f"bench.py: host-streamed result {v_e2e} differs from the device-resident {data_term}")
        e2e = {"value": N / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int((Xh.numel() + Yh.numel()) * 8 * world), "d2h_bytes_per_step": int(out_h.numel() * 8 * world),
               "api": "gpfy_b200.ops.HotPath.elbo_valgrad_host (gpfy_svgp_elbo_begin/rows/finish), pinned host rows"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (FP64 tensor-core GEMM  C = W2 Z), timed live with CUDA events ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))
    except OSError:
        pass
    fp64_peak = float(peaks.get("dmma884_bps8_tflops", 37.1))
    g_ms, g_n = ktimes["gemm_dmma"]
    M, P = prob["M"], hp.P
    pts_per_launch = n_local * a.steps / max(g_n, 1)
    # algorithmic flops of the contraction the kernel carries: L_q^T a (M^2) and L_q (L_q^T a) (M^2) per row and
    # output, counting the triangular structure; the kernel executes 2 Mp^2 on the folded dense matrix
    gemm_alg_flops = 2.0 * M * M * pts_per_launch
    g_avg_s = g_ms * 1e-3 / max(g_n, 1)
    achieved = gemm_alg_flops / g_avg_s * 1e-12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get("gemm_dmma_kernel", {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    step_tflops = prob["alg_flops_per_point"] * value / world * 1e-12
    roofline = {
        "bound": "tensor", "kernel": "gemm_dmma_kernel (FP64 DMMA m8n8k4)", "achieved": achieved, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic,
        "peak_source": "measured FP64 DMMA peak on this pool's B200 (profiles/fp64_peak_r01.json, tools/fp64_peak.cu); "
                       "MEASURED_PEAKS.json has no FP64 entry and tcgen05 has no f64 kind",
        "avg_launch_ms": g_avg_s * 1e3, "launches": g_n, "points_per_launch": pts_per_launch,
        "alg_flops_per_point_kernel": 2.0 * M * M, "kernel_share_of_step": g_ms / (ms * a.steps),
        "step": {"alg_flops_per_point": prob["alg_flops_per_point"], "achieved_tflops_per_gpu": step_tflops,
                 "frac_of_fp64_peak": step_tflops / fp64_peak,
                 "hbm_alg_bytes_per_point": prob["alg_bytes_per_point"],
                 "hbm_alg_gbs_per_gpu": prob["alg_bytes_per_point"] * value / world * 1e-9},
        "kernels_ms_per_step": {k: v[0] / a.steps for k, v in ktimes.items()},
    }

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        S = int(a.cpu_sample)
        host = prob["host"]
        Xs, Ys = X[:S].cpu().numpy(), Y[:S].cpu().numpy()
        rate, sec, val = cpu_reference_rate(host, Xs, Ys, steps=3, warmup=1)
        # cross-check while we are here: the GPU path on the same sample agrees with the oracle
        chk = hp.elbo_valgrad_packed(X[:S], Y[:S], *args[2:])
        kl_o = float(kl[0])
        gpu_val = float(chk[0]) - kl_o
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{S} rows x 3 steps ({sec:.2f} s/step), torch {torch.__version__} float64 CPU + autograd "
                         f"(oracle/gpfy_oracle.py; JAX not installable)",
               "elbo_rel_diff_gpu_vs_oracle_on_sample": abs(gpu_val - val) / abs(val)}

    # ---- BASELINE.json configs[1]: Phi(X) only (projection A = sqrt(lambda v) r Phi), device resident ----
    phi = None
    if world == 1:
        nphi = n_local  # the full N = 10 M: Phi is 8 M N = 22.4 GB of the 180 GB
        ds = torch.sqrt(args[5] * args[4])
        # the output buffers are allocated once, outside the timed region (a fresh 22.4 GB torch allocation per call
        # made this figure swing between 16 and 29 ms with the caching allocator's state)
        out_phi = (torch.empty((M, nphi), dtype=torch.float64, device=X.device), torch.empty((nphi,), dtype=torch.float64, device=X.device))
        f = lambda: hp.features(X[:nphi], args[6], args[7], w=args[2], b=args[3], degree_scale=ds, mul_r=True, out=out_phi)
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            f()
        e1.record(); torch.cuda.synchronize()
        del out_phi
        torch.cuda.empty_cache()
        pms = e0.elapsed_time(e1) / 3
        hbm_peak = 6549.8
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except (OSError, ValueError, KeyError):
            pass
        gbs = 8.0 * (8 + M) * nphi / (pms * 1e-3) * 1e-9
        phi = {"workload": "Phi(X) [M, N] feature matrix, d=8, degrees 0..10, T=%d" % a.T, "N": nphi, "value": nphi / (pms * 1e-3),
               "unit": UNIT, "ms": pms, "alg_bytes_per_point": 8 * (8 + M), "achieved_gbs": gbs, "hbm_peak_gbs": hbm_peak,
               "frac_of_hbm_peak": gbs / hbm_peak}

    # ---- a full training step through the gpfy API objects, driven on the device (DeviceTrainer: bijectors, inducing
    # basis, ELBO + gradient, KL, chain rules, Adam - no host round trip) ----
    train = None
    if world == 1:
        from gpfy_b200.device_step import DeviceTrainer
        from gpfy_b200.likelihoods import Gaussian
        from gpfy_b200.model import GP
        from gpfy_b200.spherical import PolynomialDecay
        from gpfy_b200.spherical_harmonics import SphericalHarmonics
        from gpfy_b200.variational import VariationalDistributionTriL
        sh_, q_, lik_ = SphericalHarmonics(11, phase_truncation=a.T), VariationalDistributionTriL(), Gaussian()
        m_ = GP(PolynomialDecay()).conditional(sh_, q_)
        par_ = m_.init(0, input_dim=8, likelihood=lik_, sh_features=sh_, variational_dist=q_)
        tr = DeviceTrainer(par_, m_, q_, lik_, learning_rate=1e-2)
        for _ in range(2):
            tr.step(X, Y)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            tr.step(X, Y)
        e1.record(); torch.cuda.synchronize()
        tms = e0.elapsed_time(e1) / a.steps
        ls = tr.losses()
        train = {"api": "gpfy_b200.device_step.DeviceTrainer (SphericalHarmonics(11, T) + PolynomialDecay + TriL + Gaussian, Adam)",
                 "value": N / (tms * 1e-3), "unit": UNIT, "ms_per_step": tms, "loss_first": float(ls[0]), "loss_last": float(ls[-1])}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(N, a.T, M, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "phi_only": phi, "train_step": train, "elbo_data_term": data_term, "kl": float(kl[0]),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
