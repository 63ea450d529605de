"""bench.py contract on CPU: the reference arm (oracle port on the host cores) prints ONE JSON line with
the keys the driver reads, on the same metric / unit / config object as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "synthetic:tiny",
                          "--steps", "1", "--warmup", "0", *extra], capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("streamed frames/sec") and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/longva_model.py" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert {"workload", "model", "frames_total", "sharding"} <= set(d["config"]) and d["scaling"] == "strong"


def test_reference_arm_other_ranks_are_silent():
    """Under torchrun (N > 1) only rank 0 runs and prints the reference arm."""
    assert _run("--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_uses_the_sharded_config_under_torchrun():
    """N > 1: rank 0 prints the reference line on OUR arm's config — ONE stream (strong scaling), not one per rank."""
    lines = _run("--gpus", "2", env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["frames_total"] == 160
    assert "ONE stream over 2 ranks" in d["config"]["sharding"]
    w = json.loads(_run("--gpus", "2", "--scaling", "weak", env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})[0])
    assert w["scaling"] == "weak" and w["config"]["frames_total"] == 320
