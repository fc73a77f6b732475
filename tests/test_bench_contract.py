"""bench.py prints ONE JSON line with the keys the driver reads.  CPU: the reference arm (the unmodified reference
binary + shim on the host cores, a tiny bounded sample).  GPU: the product arm on a small grid."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    # ONE line on stdout, and it is the JSON line (libraries that print to fd 1 -- NCCL's version banner -- go to stderr)
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and lines[0].startswith("{"), r.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "main_n512.out")):
        pytest.skip("oracle/_ref not built (needs /root/reference or the prebuilt binaries)")
    d = _run(["--impl", "reference", "--steps", "3", "--warmup", "1", "--ref-grid", "512"], 300)
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "rk4_grid_point_steps_per_s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_line_for_the_slab_configuration():
    """N > 1: the product arm is the slab-decomposed constant vortex (strong scaling); the reference arm must name the
    same workload and scaling (rank 0 runs it; here with a tiny grid instead of 16384^2)"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "main_n256.out")):
        pytest.skip("oracle/_ref not built (needs /root/reference or the prebuilt binaries)")
    d = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--ref-grid", "256"], 300)
    assert d["impl"] == "reference" and d["scaling"] == "strong" and d["n_gpus"] == 2
    assert "const vortex 16384^2" in d["config"]["workload"] and d["config"]["same_grid_as_product_arm"] is False
    assert d["cpu_baseline"]["single_thread"]["cores"] == 1


@pytest.mark.gpu
def test_product_arm_line_small_grid():
    d = _run(["--grid", "512", "--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "3"], 600)
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["gpu_launches"] == 8 * d["steps"]
    ro = d["roofline"]
    assert ro["bound"] == "hbm" and ro["unit"] == "GB/s" and abs(ro["frac"] - ro["achieved"] / ro["peak"]) < 1e-9
    assert set(ro["kernels"]) == {"col_step", "row_jac"}
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4 * 512 * 512 and e["d2h_bytes_per_step"] == 4 * 512 * 512
    assert e["value"] > 0 and e["results_identical"] is True and e["serial"]["value"] > 0
    assert d["config"]["state_finite"] is True
