"""bench.py's reference arm on CPU: one JSON line with the contract's keys; under a multi-rank launch only rank 0 speaks.
(The GPU arm's line is produced on a B200 and kept under profiles/; its keys are checked here on the committed copy.)"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e")


def run_bench(extra_env, *args):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.strip()]


def test_reference_arm_line():
    lines = run_bench({}, "--impl", "reference", "--workload", "c2", "--steps", "2", "--warmup", "1")
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in BASE_KEYS:
        assert k in d, k
    assert d["metric"] == "particle_steps_per_sec" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["same_config"] is True
    rc = cb["reference_compiled"]   # the unmodified reference's own number on the sample it can hold, or why it is absent
    assert ("value" in rc and rc["kind"] == "reference") or "unavailable" in rc
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    lines = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--workload", "c2",
                      "--steps", "1", "--warmup", "1")
    assert lines == []


def test_committed_gpu_lines_carry_the_contract():
    prof = os.path.join(ROOT, "profiles")
    for name in ("r02_bench.json", "r02_bench_2gpu.json", "r02_bench_8gpu.json"):
        d = json.load(open(os.path.join(prof, name)))
        for k in BASE_KEYS + ("clocks", "gpu_launches", "roofline"):
            assert k in d, (name, k)
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, (name, k)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] > 1:
            assert d["transport_parity"]["ok"] is True
            assert "cpu_baseline" in d
