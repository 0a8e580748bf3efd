"""bench.py's output contract: exactly one JSON line on stdout with the keys the driver reads.  The reference arm runs on
the CPU (it times the oracle/_ref build of the reference, or the C port); our arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
          "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("workload", ["tiny", "image_tiny"])
def test_reference_arm_line(workload):
    d = run_bench("--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1")
    assert COMMON <= set(d) and d["impl"] == "reference" and d["metric"] == "voxel_steps_per_sec" and d["unit"] == "voxel-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["config"]["workload"] == workload
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["tiny", "image_tiny"])
def test_our_arm_line(workload):
    d = run_bench("--workload", workload, "--steps", "3", "--warmup", "3")
    assert COMMON | {"roofline", "clocks", "gpu_launches"} <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] in ("issue", "hbm") and r["peak"] > 1000 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    m = r["algorithmic_model"]     # SURVEY 8(d)'s figure stays beside the bound the kernel is on
    assert m["bound"] == "hbm" and m["unit"] == "GB/s" and abs(m["frac"] - m["achieved"] / m["peak"]) < 1e-9
    assert d["timed_region_s"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 3 and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert {"sm_mhz", "sm_mhz_min", "sm_max_mhz", "reasons"} <= set(d["clocks"])
