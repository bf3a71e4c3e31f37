"""`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the
contract's keys; runs on the host cores only -- no GPU, no /root/reference."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "cfg1_512_256px", *extra], capture_output=True, text=True, timeout=600,
                         env={**os.environ, **(env or {})})
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_single_process():
    lines = _run("--gpus", "1", "--steps", "2", "--warmup", "0")
    d = json.loads(lines[-1])
    assert d["impl"] == "reference" and d["metric"] == "megapixels/sec" and d["unit"] == "MP/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] - 0.262144 / d["value"] * 1e3) < 1e-6
    cb = d["cpu_baseline"]
    # the reference's own code when oracle/_ref (or /root/reference) is there, the cost-faithful port otherwise
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg1_512_256px"


def test_reference_arm_other_ranks_do_nothing():
    assert _run("--gpus", "2", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2"}) == []


@pytest.mark.timeout(600)
def test_reference_arm_http_workers():
    lines = _run("--gpus", "2", "--steps", "1", "--warmup", "0")
    d = json.loads(lines[-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert "aiohttp" in json.dumps(d["cpu_baseline"]) or "HTTP" in json.dumps(d["cpu_baseline"])
