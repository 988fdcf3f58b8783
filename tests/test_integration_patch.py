"""the patch of INTEGRATION.md §2, compiled against the reference and run.

`oracle/_ref/main_patched` is the reference's own `main.cpp` + every stock translation unit, with `2DTissue.{h,cpp}` patched by
`oracle/apply_integration_patch.py` so that `_2DTissue::perform_particle_simulation()` is one `t2d_step_host` call into
lib2dtissue_b200.so (fp64, table criterion, the reference's own chart and distance matrix).  With T2D_INTEGRATION_CHECK set
the patched binary also runs the stock body from the same state after every step and prints the differences: the README
configuration (100 particles, 50 steps, --step-time 0.02) must agree step by step."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "main_patched")


@pytest.mark.gpu
def test_patched_reference_binary_readme_config(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/main_patched not built (needs the reference sources: make -C oracle patched)")
    env = dict(os.environ, T2D_INTEGRATION_CHECK="1")
    r = subprocess.run([BIN, "--particle-count", "100", "--step-count", "50", "--step-time", "0.02"], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("T2D_CHECK")]
    start = [l for l in lines if "start" in l]
    steps = [l for l in lines if "step" in l]
    assert len(start) == 1 and len(steps) == 50, r.stdout[-2000:]
    m = re.search(r"vid mismatches (\d+), max \|r3d diff\| ([0-9.eE+-]+)", start[0])
    assert int(m.group(1)) == 0 and float(m.group(2)) <= 1e-12
    heading_ties = 0
    for l in steps:
        m = re.search(r"heading mismatches (\d+), vid mismatches (\d+), colour mismatches (\d+), max \|uv diff\| ([0-9.eE+-]+), "
                      r"max \|r3d diff\| ([0-9.eE+-]+), max \|rdot diff\| ([0-9.eE+-]+)", l)
        assert m, l
        heading_ties += int(m.group(1))
        assert int(m.group(2)) == 0 and int(m.group(3)) == 0, l
        assert float(m.group(4)) <= 1e-9 and float(m.group(5)) <= 1e-9 and float(m.group(6)) <= 1e-9, l
    assert heading_ties <= 5, "%d heading mismatches over 50 steps x 100 particles" % heading_ties   # truncation ties only
    print("patched reference binary: 50 steps agree with the stock body; heading truncation ties: %d" % heading_ties)
    print(steps[-1])
