"""CPU: the reference arm of bench.py (the reference's CPU path on the host cores) prints ONE JSON line with the
contract's keys; under a multi-rank launch only rank 0 speaks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = {**os.environ, **(extra_env or {})}
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3",
                           "--queries", "64", "--gallery-rows", "4000", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "queries/sec @top-100" and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("c3: 64 queries x 4000 gallery")


def test_reference_arm_is_silent_on_other_ranks():
    res = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""
