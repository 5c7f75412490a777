"""bench.py prints exactly ONE JSON line on stdout (library chatter goes to stderr); checked here with the CPU arm."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--map-size", "128", "--cascades-per-set", "1", "--sets", "2"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    # the CPU arm is the reference's own shaders compiled for the CPU (oracle/_ref) wherever that library exists
    from oracle import pyref
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if pyref.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["seconds_per_step"]["median"] > 0
    assert d["config"]["cascades_per_step_per_gpu"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
