"""bench.py's reference arm runs without a GPU (it times the CPU oracle port): check the JSON contract here so a
broken line is caught before the round-end driver runs it."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout, got %d" % len(lines)
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "gae_fwd_bwd_trajectory_steps_per_sec" and j["unit"] == "steps/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["dtype"] == "f32"
    assert j["value"] > 0 and abs(j["value"] - 1024 * 65536 / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "T=1024 B=65536" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "T=1024 B=65536" in j["config"]["workload"] and "model" not in j["config"]
    # both arms print the SAME config (`same_config` was false because the workload strings differed)
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench.config_of(1024, 65536, 1)
    assert cb["reference_origin_context"]["forward_backward_s"] == 148.7  # hpc_rll.origin itself, build-container timing
