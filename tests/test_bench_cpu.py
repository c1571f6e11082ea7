"""CPU test of bench.py's reference arm (`--impl reference`): the line contract the driver parses, on a small register
(the oracle port of QobjEvo + qutip.sesolve at QuTiP's default options -- the reference's own CPU path for this hot
path, pulser_simulation/simulation.py:729-735).  The GPU arm needs a device and is exercised on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, PB200_BENCH_ATOMS="8", PB200_REF_SAMPLE_STEPS="30", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                           "--warmup", "0"], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_line_contract():
    res = _run({})
    assert res.returncode == 0, res.stderr[-400:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "steps/s"
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * 30 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "30 consecutive" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    res = _run({"RANK": "3", "WORLD_SIZE": "8"})
    assert res.returncode == 0 and res.stdout.strip() == ""
