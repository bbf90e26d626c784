"""bench.py contract: both arms print exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "cpu_baseline"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"bench.py must print ONE line, got {len(lines)}"
    return json.loads(lines[0])


def test_own_arm_json_line():
    d = _run("--steps", "2", "--warmup", "3", "--spp", "64")
    assert BASE_KEYS | {"roofline", "per_config"} <= set(d)
    assert d["config"]["scene"] == 8 and d["config"]["width"] == 800 and d["kernel"]["bvh_nodes"] > 1, "the headline must run the BVH traversal (BASELINE config 3)"
    assert {"config 1", "config 2", "config 4"} <= set(d["per_config"]), d["per_config"]
    assert all(d["per_config"][k]["value"] > 0 for k in ("config 1", "config 2", "config 4"))
    assert d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["dtype"] == "f32" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["value"] > 0 and d["gpu_launches"] >= 2 * 2
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["d2h_bytes_per_step"] == 800 * 800 * 4
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and 0 < d["roofline"]["frac"] < 1.2
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_reference_arm_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mort_ref")):
        d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
        assert d["impl"] == "reference" and "unavailable" in d
        return
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    if "unavailable" in d:                       # the reference kernel itself is not ours to fix: report why, try once more
        print("reference arm unavailable:", d["unavailable"])
        d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert "unavailable" not in d, d
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
