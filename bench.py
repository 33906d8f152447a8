#!/usr/bin/env python
"""bench.py — ELBO+grad points/sec of the SVGP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`; the default is the one the metric is quoted on):
  cfg3  configs[2]: full SVGP ELBO + gradient step, d = 8 (ambient dim 9, alpha = 3.5), degrees 0..10, phase truncation
        T = 30 (M = 280 inducing harmonics), TriL q, Gaussian likelihood, P = 1, float64, N = 10 M rows.
  cfg4  configs[3]: airline-scale regression, d = 8, degrees 0..12 (F = 13, M = 340), T = 30, N = 100 M rows.
  cfg5  configs[4]: d = 32 (ambient dim 33), degrees 0..4, T = 64 (M = 226), Bernoulli likelihood, N = 20 M rows.
Rows are sharded over the ranks (strong scaling: the global data set is the same for every world size, see
gpfy_b200.synthetic.global_rows), one all-reduce of the packed buffer per step.

A step = one ELBO value + all gradients over every row of the data set (+ the KL term and the all-reduce).  `value`
has the rows resident in HBM; `e2e` streams them from pinned host memory through `HotPath.elbo_valgrad_host` and reads
the packed result back, every step.

`--impl reference` times the CPU restatement of the reference (oracle/gpfy_oracle.py, torch float64 on the host cores
with autograd — JAX is not installable in this image) on a bounded sample of the same workload.
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

UNIT = "points/s"
CONFIGS = {
    "cfg3": dict(d=8, F=11, T=30, lik="gaussian", N=10_000_000, seed=2, metric="ELBO+grad points/sec at N=10M d=8 deg10",
                 kernel="PolynomialDecay(truncation_level=10)"),
    "cfg4": dict(d=8, F=13, T=30, lik="gaussian", N=100_000_000, seed=3, metric="ELBO+grad points/sec at N=100M d=8 deg12",
                 kernel="PolynomialDecay(truncation_level=10)"),
    "cfg5": dict(d=32, F=5, T=64, lik="bernoulli", N=20_000_000, seed=4, metric="ELBO+grad points/sec at N=20M d=32 deg4 Bernoulli",
                 kernel="PolynomialDecay(truncation_level=10)"),
}
PREFIX_ROWS = 200_000   # rows of the sharded-vs-single / two-device self-checks
PRECISION = {"f64": "f64", "f64dmma": "f64dmma", "f32": "f32"}   # --dtype -> HotPath precision
CONTRACTION = {
    "f64": "C = W2 Z on the FP64 tensor cores (DMMA m8n8k4)",
    "f64i8": "C = W2 Z as an error-free split into 6 x 6 signed radix-256 digits, 21 tcgen05 kind::i8 products into exact int32 TMEM "
             "accumulators, recombined in float64 (normwise ~1e-14; GPFY_PREC_F64_I8); everything else float64",
    "f32": "C = W2 Z and the weighted Gram matrix as 3 x TF32 on tcgen05 (GPFY_PREC_TF32X3); everything else float64",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=float, default=None, help="total rows over all ranks (default: the config's N)")
    ap.add_argument("--T", type=int, default=None, help="phase truncation (default: the config's T)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f64dmma", "f32"],
                    help="f64: the reference's precision (headline): float64 results, the dense contraction as exact int8 digit products "
                         "on tcgen05 where built (GPFY_PREC_F64_I8), DMMA elsewhere.  f64dmma: DMMA only (GPFY_PREC_F64).  "
                         "f32: the contraction as 3 x TF32 on tcgen05 (GPFY_PREC_TF32X3)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Phi-only / predict_diag / train-step side measurements")
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="rows per CPU-baseline step")
    a = ap.parse_args()
    c = dict(CONFIGS[a.config])
    if a.n is not None:
        c["N"] = int(a.n)
    if a.T is not None:
        c["T"] = int(a.T)
    c["precision"] = PRECISION[a.dtype]
    a.cfg = c
    return a


def workload_config(name, c, M, world):
    return {
        "workload": f"svgp_elbo_valgrad/{name}", "N_total": int(c["N"]), "input_dim": c["d"], "ambient_dim": c["d"] + 1,
        "num_frequencies": c["F"], "max_degree": c["F"] - 1, "phase_truncation": int(c["T"]), "num_inducing": int(M), "num_outputs": 1,
        "likelihood": c["lik"], "q": "tril", "kernel": c["kernel"], "parallelism": f"rows/{world}",
        "data": "gpfy_b200.synthetic.global_rows(seed=%d): one global data set, identical for every world size and both arms" % c["seed"],
        "l2": "per-step inputs (>= 80 MB/rank) plus the per-chunk panels (Z, C: GBs per chunk) exceed the 126 MB L2",
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
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
        note = None
        if not self.lines and getattr(self, "warm", None):
            self.lines, note = list(self.warm), "timed region shorter than the 100 ms sampling period: the last warm-up samples (same workload, back to back)"
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
        out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
               "reasons": sorted(reasons)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference (oracle) on the host cores
# ------------------------------------------------------------------------------------------------
def oracle_pack(host):
    return {"dim": host["dim"], "alpha": host["alpha"], "V": host["V"], "L": host["L"], "w": host["w"], "b": host["b"],
            "v": host["v"], "eig": host["eig"], "order": host["order"], "mu": host["mu"], "Lq": host["Lq"],
            "sigma2": host["sigma2"], "lik": host["lik"]}


def cpu_reference_rate(host, X_sample, Y_sample, steps, warmup):
    """points/s of oracle ELBO+grad (torch float64 CPU, autograd) on `steps` passes over the sample."""
    import torch
    from oracle import gpfy_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    pack = oracle_pack(host)
    for _ in range(warmup):
        O.elbo_and_grads(pack, X_sample, Y_sample)
    t0 = time.perf_counter()
    for _ in range(steps):
        val, grads = O.elbo_and_grads(pack, X_sample, Y_sample)
    dt = time.perf_counter() - t0
    return X_sample.shape[0] * steps / dt, dt / steps, val, grads


def run_reference(a):
    """`--impl reference`: the oracle port on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from gpfy_b200 import synthetic
    c = a.cfg
    S = int(a.cpu_sample)
    host, X, Y, m = synthetic.headline_host_problem(S, d=c["d"], F=c["F"], T=c["T"], likelihood=c["lik"], seed=c["seed"])
    rate, sec, _, _ = cpu_reference_rate(host, X, Y, max(1, a.steps), max(1, a.warmup))
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": c["metric"], "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(a.config, c, sum(m), a.gpus),
        "rows_per_step": S,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step processes rows [0, {S}) of the config's N_total = {c['N']} global rows (a bounded sample; "
                                   f"the rate is per row), torch {torch.__version__} float64 CPU + autograd "
                                   f"(oracle/gpfy_oracle.py; JAX is not installable here)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# self-checks that need more than one GPU (pytest only ever gets one): run under torchrun after the timed region
# ------------------------------------------------------------------------------------------------
def packed_checksum(hp, packed):
    """data term and the 2-norm of every gradient block of an (all-reduced) packed buffer."""
    import torch
    g = hp.unpack(packed)
    out = {"data_term": float(g["data_term"].item())}
    for k, v in g.items():
        if k != "data_term":
            out["norm_" + k] = float(torch.linalg.vector_norm(v.reshape(-1)).item())
    return out


def sharded_prefix_check(c, dev, rank, world, allreduce):
    """Every rank takes its shard of the first PREFIX_ROWS global rows, the packed buffers are all-reduced; rank 0 also
    computes the whole prefix alone.  Returns max |sharded - single| / max |single| (rank 0), to be <= 1e-12."""
    import torch
    from gpfy_b200 import synthetic
    n = min(PREFIX_ROWS, c["N"])
    kw = dict(d=c["d"], F=c["F"], T=c["T"], likelihood=c["lik"], seed=c["seed"], device=dev, precision=c.get("precision", "f64"))
    mine = synthetic.headline_problem(n, rank=rank, world=world, **kw)
    part = mine["hot"].elbo_valgrad_packed(*mine["args"]).clone()
    allreduce(part)
    torch.cuda.synchronize()
    if rank != 0:
        return None
    full = synthetic.headline_problem(n, rank=0, world=1, **kw)
    ref = full["hot"].elbo_valgrad_packed(*full["args"])
    return float(((part - ref).abs().max() / ref.abs().max()).item())


def two_device_threads_check(c, devices):
    """ONE process, two host threads, one device each (XLA's execution model): each thread runs the hot path
    on its half of the prefix on its own device, concurrently; the host sum must equal the single-device result."""
    import torch
    from gpfy_b200 import synthetic
    n = min(PREFIX_ROWS, c["N"])
    kw = dict(d=c["d"], F=c["F"], T=c["T"], likelihood=c["lik"], seed=c["seed"], precision=c.get("precision", "f64"))
    res, err = [None, None], [None, None]

    def work(i):
        try:
            torch.cuda.set_device(devices[i])
            p = synthetic.headline_problem(n, rank=i, world=2, device=devices[i], **kw)
            for _ in range(3):   # repeated so that the two threads really overlap inside the library
                out = p["hot"].elbo_valgrad_packed(*p["args"])
            torch.cuda.synchronize(devices[i])
            res[i] = out.cpu()
        except Exception as e:  # noqa: BLE001 - reported in the JSON line
            err[i] = repr(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if any(err):
        return {"error": [e for e in err if e]}
    torch.cuda.set_device(devices[0])
    full = synthetic.headline_problem(n, rank=0, world=1, device=devices[0], **kw)
    ref = full["hot"].elbo_valgrad_packed(*full["args"]).cpu()
    return {"rel": float(((res[0] + res[1] - ref).abs().max() / ref.abs().max()).item()), "devices": [str(d) for d in devices]}


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gpfy_b200 import distributed as gd
    from gpfy_b200 import ops, synthetic

    c = a.cfg
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
    N = int(c["N"])
    prob = synthetic.headline_problem(N, d=c["d"], F=c["F"], T=c["T"], likelihood=c["lik"], seed=c["seed"], device=dev,
                                      rank=rank, world=world, precision=PRECISION[a.dtype])
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

    # rank 0 samples its own GPU (eight concurrent nvidia-smi loops stall the driver and with it every rank's launches).  The sampler is
    # started BEFORE the warm-up steps and its lines are discarded at the start of the timed region: the first nvidia-smi attach on a
    # fresh box takes seconds and stalls launches on every GPU of the node while it lasts - started right before the timed region it
    # once put 27 ms per step onto a 4-GPU run (profiles: 46.8 ms against 19.5 ms in the two runs after it).
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_attach = time.time()
        while not sampler.lines and sampler.proc is not None and time.time() - t_attach < 20.0:
            time.sleep(0.05)     # the first sample is out: the attach is over
    barrier()
    for _ in range(max(3, a.warmup)):
        step()
    barrier()

    # ---- timed region: K steps, device time, max over ranks ----
    if rank == 0:
        sampler.warm = list(sampler.lines[-2:])   # the same workload, back to back, right before the timed region
        sampler.lines.clear()
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
    checksum = packed_checksum(hp, packed)

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
               "api": "gpfy_b200.ops.HotPath.elbo_valgrad_host (gpfy_svgp_elbo_begin/rows/finish), pinned host rows, ramped slabs"}
        del Xh, Yh

    # ---- multi-GPU self-checks (all ranks take part in the first; rank 0 alone runs the second) ----
    checks = {}
    if world > 1:
        rel = sharded_prefix_check(c, dev, rank, world, allreduce)
        barrier()
        if rank == 0:
            checks["sharded_vs_single_prefix"] = {"rows": min(PREFIX_ROWS, N), "rel": rel, "tolerance": 1e-12}
            if not (rel <= 1e-12):
                raise SystemExit(f"bench.py: sharded + all-reduced result differs from the single-GPU one on the same rows: rel {rel}")
            if torch.cuda.device_count() >= 2:
                other = torch.device("cuda", (local + 1) % torch.cuda.device_count())
                td = two_device_threads_check(c, [dev, other])
                checks["one_process_two_threads_two_devices"] = td
                if "error" in td or not (td["rel"] <= 1e-12):
                    raise SystemExit(f"bench.py: one process driving two devices from two threads failed: {td}")
                torch.cuda.set_device(local)
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream (gpfy_debug_kernel_timing) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))
    except OSError:
        pass
    fp64_peak = float(peaks.get("dmma884_bps8_tflops", 37.1))
    fp64_src = ("measured FP64 DMMA peak on this pool's B200 (profiles/fp64_peak_r01.json, tools/fp64_peak.cu); "
                "MEASURED_PEAKS.json has no FP64 entry and tcgen05 has no f64 kind")
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    M, P = prob["M"], hp.P
    Mp = (M + 31) // 32 * 32
    path = hp.contraction_path(n_local)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except (OSError, ValueError):
        pass

    def kernel_line(key, name, alg_per_point, peak, unit, bound, src, ncu_key=None):
        k_ms, k_n = ktimes.get(key, (0.0, 0))
        if k_n == 0 or k_ms <= 0.0:
            return None
        pts = n_local * a.steps / k_n
        avg_s = k_ms * 1e-3 / k_n
        ach = alg_per_point * pts / avg_s * (1e-12 if unit.startswith("TFLOP") else 1e-9)
        traffic = ncu.get(ncu_key, {}).get("dram_bytes_per_launch") if (ncu_key and a.config == "cfg3" and ncu.get("dtype", "f64dmma") == a.dtype) else None
        return {"bound": bound, "kernel": name, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": traffic,
                "peak_source": src, "avg_launch_ms": avg_s * 1e3, "launches": k_n, "points_per_launch": pts,
                "alg_per_point_kernel": alg_per_point, "kernel_share_of_step": k_ms / (ms * a.steps)}

    classes = {}
    if path == "f64i8":
        # the float64 contraction as 21 int8 digit products: algorithmic flops 2 M^2 per point, against the int8 tensor peak expressed
        # in float64-equivalent flops (8192 int8 MAC/clk/SM measured by tools/i8_probe.cu at N = 256 = 4.77 POP/s at 1.965 GHz, / 21)
        i8_peak = 2.0 * 8192 * 148 * 1.965e9 * 1e-12
        classes["gemm"] = kernel_line("gemm_dmma", "gemm_i8_kernel (tcgen05 kind::i8, 21 digit products per float64 product, TMEM accumulators, TMA bulk loads)",
                                      2.0 * M * M, i8_peak / 21.0, "TFLOP/s", "tensor",
                                      "int8 dense peak measured on this pool's B200 (tools/i8_probe.cu: 8191 MAC/clk/SM back to back = 4 765 TOP/s at 1 965 MHz) "
                                      "/ 21 digit products per float64-accurate product; achieved counts the algorithmic 2 M^2 flops per point once",
                                      "gemm_i8_kernel")
        classes["i8_slice"] = kernel_line("i8_slice", "z_slice_i8_kernel (float64 panel -> six int8 digit planes)", 14.0 * Mp,
                                          float(measured.get("hbm_gbs", 6549.8)), "GB/s", "hbm",
                                          "MEASURED_PEAKS.json hbm_gbs; bytes = 8 Mp read + 6 Mp written per point", "z_slice_i8_kernel")
    elif path == "f32":
        bf16 = float(measured.get("bf16_tflops", 1638.0))
        classes["gemm"] = kernel_line("gemm_dmma", "gemm_tf32x3_kernel (tcgen05 kind::tf32, 3 MMAs per product, TMEM accumulators, TMA tensor loads)",
                                      2.0 * M * M, bf16 / 2.0 / 3.0, "TFLOP/s", "tensor",
                                      "MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 runs at half the bf16 rate) / 3 (three TF32 MMAs per "
                                      "fp32-accurate product); achieved counts the algorithmic 2 M^2 flops per point once")
    else:
        # algorithmic flops of the contraction: L_q^T a (M^2) and L_q (L_q^T a) (M^2) per row and output; the kernel
        # executes 2 Mp^2 on the folded dense matrix
        classes["gemm"] = kernel_line("gemm_dmma", "gemm_dmma_kernel (FP64 DMMA m8n8k4)", 2.0 * M * M, fp64_peak, "TFLOP/s", "tensor", fp64_src,
                                      "gemm_dmma_kernel")
    # weighted Gram matrix G = Z diag(wq) Z^T: M^2 algorithmic flops per point (the symmetric half)
    if path == "f32":   # the weighted Gram runs as 3 x TF32 on tcgen05 too (where its two stages fit; the FP64 kernel otherwise)
        classes["syrk"] = kernel_line("syrk_dmma", "syrk_tf32x3_kernel (tcgen05 kind::tf32, 3 MMAs per product; syrk_dmma_kernel where it does not fit)",
                                      1.0 * M * M, float(measured.get("bf16_tflops", 1638.0)) / 2.0 / 3.0, "TFLOP/s", "tensor",
                                      "MEASURED_PEAKS.json bf16_tflops / 2 / 3 (see gemm)")
    else:
        classes["syrk"] = kernel_line("syrk_dmma", "syrk_dmma_kernel (FP64 DMMA m8n8k4)", 1.0 * M * M, fp64_peak, "TFLOP/s", "tensor", fp64_src,
                                      "syrk_dmma_kernel")
    z_ms = ktimes["zonal_fwd"][0] + ktimes["zonal_bwd"][0] + ktimes["lik_point"][0] + ktimes["dvhat_dmma"][0]
    classes = {k: v for k, v in classes.items() if v}
    dominant = max(classes.values(), key=lambda v: v["kernel_share_of_step"])
    step_tflops = prob["alg_flops_per_point"] * value / world * 1e-12
    roofline = dict(dominant)
    roofline["contraction_path"] = path
    roofline["kernels"] = {k: {kk: v[kk] for kk in ("kernel", "bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "kernel_share_of_step")}
                           for k, v in classes.items()}
    roofline["kernels"]["zonal"] = {"kernel": "zonal_fwd3/2 + zonal_bwd3/4 (+ lik_point, dvhat)", "bound": "fp64",
                                    "achieved": (prob["alg_flops_per_point"] - 3.0 * M * M) * n_local / (z_ms * 1e-3 / a.steps) * 1e-12 if z_ms > 0 else None,
                                    "peak": fp64_peak, "unit": "TFLOP/s", "kernel_share_of_step": z_ms / (ms * a.steps)}
    roofline["step"] = {"alg_flops_per_point": prob["alg_flops_per_point"], "achieved_tflops_per_gpu": step_tflops,
                        "frac_of_fp64_peak": step_tflops / fp64_peak,
                        "note": ("algorithmic float64 flops per second over the FP64 pipe's peak; with the contraction on the int8 tensor cores "
                                 "(contraction_path f64i8) 2 M^2 of them no longer run on that pipe, so this ratio can exceed what the FP64 pipe alone allows"),
                        "hbm_alg_bytes_per_point": prob["alg_bytes_per_point"],
                        "hbm_alg_gbs_per_gpu": prob["alg_bytes_per_point"] * value / world * 1e-9}
    roofline["kernels_ms_per_step"] = {k: v[0] / a.steps for k, v in ktimes.items()}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        S = int(a.cpu_sample)
        host = prob["host"]
        Xs, Ys = X[:S].cpu().numpy(), Y[:S].cpu().numpy()
        rate, sec, val, og = cpu_reference_rate(host, Xs, Ys, steps=3, warmup=1)
        # cross-check while we are here: the GPU path on the same sample agrees with the oracle, value AND gradients
        chk = hp.unpack(hp.elbo_valgrad_packed(X[:S], Y[:S], *args[2:]))
        kl_g = hp.prior_kl_valgrad(mu_d, sig_d)
        rel = lambda x, y: float(np.max(np.abs(x - y)) / max(np.max(np.abs(y)), 1e-300))
        gpu_val = float(chk["data_term"]) - float(kl_g[0])
        F = len(host["V"])
        gerr = {
            "mu": rel(chk["mu"].cpu().numpy() - kl_g[1].cpu().numpy(), og["mu"]),
            "Lq": rel(chk["sigma"].cpu().numpy() - kl_g[2].cpu().numpy(), og["Lq"]),
            "vhat": rel(chk["vhat"].cpu().numpy(), np.concatenate([og[f"V{i}"] for i in range(F)], 0)),
            "w": rel(chk["w"].cpu().numpy(), og["w"]), "b": rel(chk["b"].cpu().numpy(), og["b"]),
            "variance": rel(chk["v"].cpu().numpy(), og["v"]), "eig": rel(chk["eig"].cpu().numpy(), og["eig"]),
        }
        if c["lik"] == "gaussian":
            gerr["sigma2"] = rel(chk["sigma2"].cpu().numpy(), og["sigma2"])
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"rows [0, {S}) x 3 steps ({sec:.2f} s/step), torch {torch.__version__} float64 CPU + autograd "
                         f"(oracle/gpfy_oracle.py; JAX not installable)",
               "elbo_rel_diff_gpu_vs_oracle_on_sample": abs(gpu_val - val) / abs(val),
               "grad_rel_diff_gpu_vs_oracle_on_sample": gerr}

    f32_err = None
    if (a.dtype == "f32" or path == "f64i8") and world == 1:
        # the f32 instantiation against the f64 (DMMA) path on the same rows (tolerance of BASELINE.json north_star: 1e-5); the int8
        # digit-product contraction is a float64 path: held to 1e-11 against DMMA
        tol_alt = 1e-5 if a.dtype == "f32" else 1e-11
        S = min(PREFIX_ROWS, N)
        p64 = synthetic.headline_problem(S, d=c["d"], F=c["F"], T=c["T"], likelihood=c["lik"], seed=c["seed"], device=dev, precision="f64dmma")
        ref = p64["hot"].unpack(p64["hot"].elbo_valgrad_packed(*p64["args"]))
        got = hp.unpack(hp.elbo_valgrad_packed(X[:S], Y[:S], *args[2:]))
        f32_err = {"rows": S, "tolerance": tol_alt}
        for k in ref:
            den = float(ref[k].abs().max())
            if den > 0:
                diff = (got[k] - ref[k]).abs()
                f32_err[k] = {"max_rel": float(diff.max()) / den, "mean_rel": float(diff.mean()) / den}
        if max(v["max_rel"] for v in f32_err.values() if isinstance(v, dict)) > tol_alt:
            raise SystemExit(f"bench.py: the {a.dtype} path is outside {tol_alt} of the f64 (DMMA) path: {f32_err}")
        del p64

    extras = {}
    if world == 1 and not a.no_extras:
        extras = side_measurements(a, c, prob, e0, e1)
        if path == "f64i8":
            # the same step with the contraction on the FP64 tensor cores (DMMA) only, for comparison inside the same run
            from gpfy_b200.ops import HotPath
            hd = HotPath(c["d"], prob["m"], num_outputs=1, kernel_order=1, likelihood=c["lik"], q_kind="tril", weight_mode="ard", device=dev,
                         precision="f64dmma")
            pk = torch.empty_like(packed)
            for _ in range(2):
                hd.elbo_valgrad_packed(*args, out=pk)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                hd.elbo_valgrad_packed(*args, out=pk)
            e1.record()
            torch.cuda.synchronize()
            ms_d = e0.elapsed_time(e1) / 3
            extras["f64_dmma"] = {"workload": "the same step with precision f64dmma (contraction on the FP64 tensor cores)", "value": N / (ms_d * 1e-3),
                                  "unit": UNIT, "ms_per_step": ms_d, "steps": 3,
                                  "rel_diff_of_packed_result": float(((pk - packed).abs().max() / packed.abs().max()).item())}
            del hd, pk

    line = {
        "metric": c["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if a.dtype == "f64dmma" else a.dtype, "data": "synthetic",
        "config": dict(workload_config(a.config, c, M, world), contraction=CONTRACTION[path]), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "elbo_data_term": data_term, "kl": float(kl[0]), "checksum": checksum,
        "checks": checks,
    }
    if f32_err is not None:
        line["f32_vs_f64" if a.dtype == "f32" else "i8_vs_dmma"] = f32_err
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def side_measurements(a, c, prob, e0, e1):
    """BASELINE.json configs[1] (Phi only), predict_diag and a whole training step at the same shape, 1 GPU."""
    import torch
    hp, args, M = prob["hot"], prob["args"], prob["M"]
    X, Y = args[0], args[1]
    d, N = c["d"], X.shape[0]
    hbm_peak = 6549.8
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        pass
    fp64_peak = 37.141
    out = {}

    # ---- Phi(X) only (projection A = sqrt(lambda v) r Phi), device resident ----
    nphi = min(N, int(40e9 / (8 * M)))   # Phi is 8 M N bytes: 22.4 GB at the headline shape, capped at 40 GB
    ds = torch.sqrt(args[5] * args[4])
    # the output buffers are allocated once, outside the timed region
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
    from gpfy_b200.synthetic import algorithmic_flops_per_point
    phi_flops = algorithmic_flops_per_point(d + 1, prob["m"], with_grad=False)
    gbs = 8.0 * (d + M) * nphi / (pms * 1e-3) * 1e-9
    tf = phi_flops * nphi / (pms * 1e-3) * 1e-12
    out["phi_only"] = {"workload": "Phi(X) [M, N] feature matrix (BASELINE configs[1] at this shape)", "N": nphi,
                       "value": nphi / (pms * 1e-3), "unit": UNIT, "ms": pms, "alg_bytes_per_point": 8 * (d + M),
                       "achieved_gbs": gbs, "hbm_peak_gbs": hbm_peak, "frac_of_hbm_peak": gbs / hbm_peak,
                       "alg_flops_per_point": phi_flops, "achieved_tflops": tf, "frac_of_fp64_peak": tf / fp64_peak}

    # ---- predict_diag over all rows (model.py:201-229) ----
    pargs = [args[i] for i in (0, 2, 3, 4, 5, 6, 7, 8, 9)]
    for _ in range(2):
        hp.predict_diag(*pargs)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        hp.predict_diag(*pargs)
    e1.record(); torch.cuda.synchronize()
    qms = e0.elapsed_time(e1) / 3
    pred_flops = phi_flops - sum(v * v for v in prob["m"]) + M * M + 2 * M   # no whitening solve: folded into W2; psi on the lower triangle
    ptf = pred_flops * N / (qms * 1e-3) * 1e-12
    out["predict_diag"] = {"workload": "GP.predict_diag (fmu, fvar) [N, 1]", "N": N, "value": N / (qms * 1e-3), "unit": UNIT, "ms": qms,
                           "alg_flops_per_point": pred_flops, "achieved_tflops": ptf, "frac_of_fp64_peak": ptf / fp64_peak,
                           "alg_bytes_per_point": 8 * (d + 2)}

    # ---- a full training step through the gpfy API objects, driven on the device (DeviceTrainer) ----
    from gpfy_b200.device_step import DeviceTrainer
    from gpfy_b200.likelihoods import Bernoulli, Gaussian
    from gpfy_b200.model import GP
    from gpfy_b200.spherical import PolynomialDecay
    from gpfy_b200.spherical_harmonics import SphericalHarmonics
    from gpfy_b200.variational import VariationalDistributionTriL
    sh_, q_ = SphericalHarmonics(c["F"], phase_truncation=c["T"]), VariationalDistributionTriL()
    lik_ = Gaussian() if c["lik"] == "gaussian" else Bernoulli()
    m_ = GP(PolynomialDecay()).conditional(sh_, q_)
    par_ = m_.init(0, input_dim=d, likelihood=lik_, sh_features=sh_, variational_dist=q_)
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
    out["train_step"] = {"api": "gpfy_b200.device_step.DeviceTrainer (SphericalHarmonics(F, T) + PolynomialDecay + TriL, Adam)",
                         "value": N / (tms * 1e-3), "unit": UNIT, "ms_per_step": tms, "loss_first": float(ls[0]), "loss_last": float(ls[-1])}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
