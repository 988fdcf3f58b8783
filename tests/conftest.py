import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_device_count():
    """Devices the CUDA runtime sees, without importing torch (libcudart is what the library itself links)."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            n = ctypes.c_int(0)
            return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
        except OSError:
            continue
    return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing in t2d_create
    (the library has no CPU fallback, so those tests cannot run there)."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: lib2dtissue_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def t2d_module():
    return importlib.import_module("2dtissue_b200")


@pytest.fixture(scope="session")
def t2d():
    return t2d_module()


@pytest.fixture(scope="session")
def chart():
    return t2d_module().load_chart(os.path.join(GOLD, "ellipsoid_x4.t2dchart"))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oraclebind
    oraclebind.build()
    return oraclebind


@pytest.fixture(scope="session")
def oracle(chart, oracle_mod):
    return oracle_mod.Oracle(chart)


@pytest.fixture(scope="session")
def hop_table(oracle):
    """uint8 hop-count table (pinned against the reference's table by sha256 in test_oracle_vs_reference)."""
    return oracle.build_hop_table()


def metric_table(x3d):
    """Synthetic metric table used by step_metric_N1500.npz: float32(Euclidean vertex distance)."""
    X = np.asarray(x3d, dtype=np.float64)
    D = np.zeros((len(X), len(X)), dtype=np.float32)
    for i0 in range(0, len(X), 512):
        d = X[i0:i0 + 512, None, :] - X[None, :, :]
        D[i0:i0 + 512] = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).astype(np.float32)
    return D


@pytest.fixture(scope="session")
def metric_tab(chart):
    return metric_table(chart["x3d"])


def golden(name):
    return np.load(os.path.join(GOLD, name))


STEP_FIXTURES = ["step_table_N100", "step_table_dense_N1500", "step_table_wide_N1500", "step_table_noise_N800",
                 "step_metric_N1500", "step_euclid_N2000", "step_euclid_dense_N1500"]


@pytest.fixture(scope="session")
def hd_lib():
    import ctypes
    src = os.path.join(ROOT, "tests", "support", "hd_selftest.cpp")
    out = os.path.join(ROOT, "tests", "support", "libhd_selftest.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "2dtissue_b200", "csrc", "hd_math.cuh"))):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                               src, "-o", out])
    L = ctypes.CDLL(out)
    L.hd_point_triangle_distance.restype = ctypes.c_double
    L.hd_philox_uniform.restype = ctypes.c_double
    L.hd_philox_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    L.hd_noise_deg.restype = ctypes.c_double
    L.hd_noise_deg.argtypes = [ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    L.hd_pair_fij.restype = ctypes.c_double
    L.hd_pair_fij.argtypes = [ctypes.c_double] * 3
    L.hd_inside.argtypes = [ctypes.c_double] * 2
    return L
