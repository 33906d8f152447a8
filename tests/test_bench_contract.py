"""bench.py's JSON contract on the CPU: the reference arm (oracle port on the host cores) runs without a GPU and
prints one line with the agreed keys; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "2000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("ELBO+grad points/sec") and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "svgp_elbo_valgrad/cfg3" and line["config"]["num_inducing"] == 280
    assert line["config"]["N_total"] == 10_000_000 and line["rows_per_step"] == 2000   # the config's N, and what a step really processed
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None


def test_reference_arm_other_configs():
    """--config cfg5 (d = 32, Bernoulli) and cfg4 (13 degrees) on the CPU arm: same line, their own shapes."""
    for cfg, M, lik in (("cfg5", 226, "bernoulli"), ("cfg4", 340, "gaussian")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg, "--steps", "1",
                              "--warmup", "1", "--cpu-sample", "1000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["config"]["workload"] == "svgp_elbo_valgrad/" + cfg and line["config"]["num_inducing"] == M
        assert line["config"]["likelihood"] == lik and line["value"] > 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CPU-only check")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_graft_entry_build_exports_the_abi():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
