"""The C++ host driver `t2d_sim` (2dtissue_b200/host/): reference CLI + `_2DTissue` mirror above the C ABI.
CPU tier: argument handling, chart loading from the files the reference's setup writes, loud failure without a GPU.
GPU tier: a run from a saved state equals the same run through the Python host, bit for bit; CSV export format."""
import importlib
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "2dtissue_b200", "t2d_sim")
CHART = os.path.join(ROOT, "tests", "golden", "ellipsoid_x4.t2dchart")


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "2dtissue_b200", "host")])
    return SIM


def write_state(path, uv, n, step=0):
    with open(path, "wb") as f:
        f.write(b"T2DSTATE" + struct.pack("<qq", n.size, step) + np.asarray(uv, "<f8").tobytes() + np.asarray(n, "<i4").tobytes())


def read_state(path):
    b = open(path, "rb").read()
    N, step = struct.unpack("<qq", b[8:24])
    uv = np.frombuffer(b[24:24 + 16 * N], "<f8").copy()
    n = np.frombuffer(b[24 + 16 * N:24 + 20 * N], "<i4").copy()
    return uv, n, step


def test_cli_help_and_bad_arguments(sim):
    r = subprocess.run([sim, "--help"], capture_output=True, text=True)
    assert r.returncode == 0
    for flag in ("--step-count", "--save-data", "--particle-innenleben", "--optimized-monotile-boundary", "--mesh-path",
                 "--particle-count", "--step-time", "--kafka"):          # the reference's eight flags, src/main.cpp:24-52
        assert flag in r.stderr
    r = subprocess.run([sim, "--step-count"], capture_output=True, text=True)
    assert r.returncode == 1 and "Too few arguments" in r.stderr        # argparse's behaviour, main.cpp:58-63
    r = subprocess.run([sim, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1


def test_chart_from_reference_setup_files(sim, chart, tmp_path):
    """<stem>_uv.off + <stem>_open.off as SurfaceParametrization::create_uv_surface writes them -> the same arrays the
    reference holds in memory (tests/golden/ellipsoid_x4.t2dchart was exported from the compiled reference)."""
    mesh = os.path.join(ROOT, "oracle", "_ref", "mcl", "meshes", "ellipsoid_x4.off")
    if not (os.path.exists(mesh[:-4] + "_uv.off") and os.path.exists(mesh[:-4] + "_open.off")):
        pytest.skip("the reference's setup files are not present (oracle/_ref is built where /root/reference exists)")
    out = str(tmp_path / "c.t2dchart")
    subprocess.check_call([sim, "--mesh-path", mesh, "--dump-chart", out])
    t2d = importlib.import_module("2dtissue_b200")
    c = t2d.load_chart(out)
    for k in ("uv", "x3d", "faces"):
        assert np.array_equal(c[k], chart[k]), k


def test_missing_chart_is_reported(sim, tmp_path):
    r = subprocess.run([sim, "--mesh-path", str(tmp_path / "nothing.off"), "--dump-chart", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Failed to open" in r.stderr


def test_no_gpu_fails_loudly(sim):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sim, "--mesh-path", CHART, "--step-count", "1", "--particle-count", "4"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("neigh", ["table", "euclid"])
def test_driver_equals_python_host(sim, t2d, chart, hop_table, tmp_path, neigh):
    N, steps = 500, 7
    uv, n = t2d.seed_particles(N, seed=11)
    s_in, s_out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_state(s_in, uv, n)
    data_dir = tmp_path / "data"
    data_dir.mkdir()
    sigma = 0.4166666666666667 if neigh == "table" else 0.3
    r = subprocess.run([sim, "--mesh-path", CHART, "--particle-count", str(N), "--step-count", str(steps), "--step-time", "0.02",
                        "--neigh", neigh, "--sigma", repr(sigma), "--load-state", s_in, "--save-state", s_out, "--save-data",
                        "--data-dir", str(data_dir)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Step: 0" in r.stdout and "Step: %d" % (steps - 1) in r.stdout and "Time taken:" in r.stdout
    mode = t2d.NEIGH_TABLE if neigh == "table" else t2d.NEIGH_EUCLID
    ctx = t2d.Context(chart, table=hop_table if neigh == "table" else None, v0=0.02, sigma=sigma, neigh_mode=mode, capacity=N)
    ctx.set_particles(uv, n)
    assert ctx.step(steps) == 0
    ref = ctx.download()
    uv2, n2, step = read_state(s_out)
    assert step == steps and np.array_equal(uv2, ref["uv"]) and np.array_equal(n2, ref["n"])
    # --save-data: r_data_<step>.csv etc., precision 15, one row per particle (IO.h:41-69, 2DTissue.cpp:270-280)
    last = np.loadtxt(data_dir / ("r_data_%d.csv" % steps), delimiter=",")
    assert last.shape == (N, 2) and np.allclose(last[:, 0], ref["uv"][:N], rtol=1e-14, atol=1e-15)
    col = np.loadtxt(data_dir / ("particles_color_%d.csv" % steps), delimiter=",")
    assert np.array_equal(col.astype(np.int32), ref["color"])
    r3 = np.loadtxt(data_dir / ("r_data_3D_%d.csv" % steps), delimiter=",")
    assert r3.shape == (N, 3)
    # resume: 3 more steps from the saved state == 10 steps straight (counter-based RNG)
    s_out2 = str(tmp_path / "out2.bin")
    r = subprocess.run([sim, "--mesh-path", CHART, "--particle-count", str(N), "--step-count", str(steps + 3), "--step-time", "0.02",
                        "--neigh", neigh, "--sigma", repr(sigma), "--load-state", s_out, "--save-state", s_out2, "--quiet",
                        "--no-particles"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert ctx.step(3) == 0
    ref2 = ctx.download()
    uv3, n3, step3 = read_state(s_out2)
    assert step3 == steps + 3 and np.array_equal(uv3, ref2["uv"]) and np.array_equal(n3, ref2["n"])


@pytest.mark.gpu
def test_driver_export_every_k_async(sim, t2d, chart, tmp_path):
    """Row f2: --export-every k — the CSV export of every k-th step only, fetched by the asynchronous export (side stream +
    pinned ring) while the next block runs; the files equal what the Python host downloads at those steps, and the final
    state equals a straight run.  --device-seed starts from t2d_seed_particles."""
    N, steps, k = 800, 12, 4
    uv, n = t2d.seed_particles(N, seed=12)
    s_in, s_out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_state(s_in, uv, n)
    data_dir = tmp_path / "data"
    data_dir.mkdir()
    r = subprocess.run([sim, "--mesh-path", CHART, "--particle-count", str(N), "--step-count", str(steps), "--step-time", "0.02",
                        "--neigh", "euclid", "--sigma", "0.3", "--load-state", s_in, "--save-state", s_out, "--save-data",
                        "--data-dir", str(data_dir), "--export-every", str(k), "--no-particles"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Steps: 0 .. 3" in r.stdout and "Steps: 8 .. 11" in r.stdout
    ctx = t2d.Context(chart, v0=0.02, sigma=0.3, neigh_mode=t2d.NEIGH_EUCLID, capacity=N)
    ctx.set_particles(uv, n)
    for b in range(steps // k):
        assert ctx.step(k) == 0
        ref = ctx.download()
        step = (b + 1) * k
        got = np.loadtxt(data_dir / ("r_data_%d.csv" % step), delimiter=",")
        assert got.shape == (N, 2) and np.allclose(got[:, 0], ref["uv"][:N], rtol=1e-14, atol=1e-15)
        col = np.loadtxt(data_dir / ("particles_color_%d.csv" % step), delimiter=",")
        assert np.array_equal(col.astype(np.int32), ref["color"])
    assert not (data_dir / "r_data_5.csv").exists()
    uv2, n2, step = read_state(s_out)
    assert step == steps and np.array_equal(uv2, ref["uv"]) and np.array_equal(n2, ref["n"])
    r = subprocess.run([sim, "--mesh-path", CHART, "--particle-count", "300", "--step-count", "3", "--neigh", "euclid", "--sigma", "0.3",
                        "--device-seed", "--seed", "7", "--quiet", "--no-particles"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_driver_table_cache_geodesic(sim, t2d, chart, tmp_path):
    """Row f4: `t2d_sim --table FILE` — the metric geodesic table as a T2DCSR1 binary cache (TableCSR.geodesic().save());
    the driver's run equals the Python host fed the same rows."""
    N, steps, sigma = 400, 5, 0.35
    csr = t2d.TableCSR.geodesic(chart, 2.4 * sigma + 1e-9)
    cache = str(tmp_path / "geo.t2dcsr")
    csr.save(cache)
    uv, n = t2d.seed_particles(N, seed=21)
    s_in, s_out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_state(s_in, uv, n)
    r = subprocess.run([sim, "--mesh-path", CHART, "--particle-count", str(N), "--step-count", str(steps), "--step-time", "0.02",
                        "--neigh", "table", "--table", cache, "--sigma", repr(sigma), "--load-state", s_in, "--save-state", s_out,
                        "--quiet", "--no-particles"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ctx = t2d.Context(chart, table=csr, v0=0.02, sigma=sigma, neigh_mode=t2d.NEIGH_TABLE, capacity=N)
    ctx.set_particles(uv, n)
    assert ctx.step(steps) == 0
    ref = ctx.download()
    uv2, n2, step = read_state(s_out)
    assert step == steps and np.array_equal(uv2, ref["uv"]) and np.array_equal(n2, ref["n"])
