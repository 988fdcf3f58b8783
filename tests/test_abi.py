"""The C-ABI library loads (no GPU needed) and exports every symbol include/t2d.h declares."""
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "t2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t2d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib_mod = importlib.import_module("2dtissue_b200._lib")
    L = lib_mod.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "lib2dtissue_b200.so does not export %s" % n
    assert set(lib_mod.SYMBOLS) == set(names)
    assert L.t2d_version() >= 100


def test_create_fails_loudly_without_gpu(chart):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    t2d = importlib.import_module("2dtissue_b200")
    with pytest.raises(t2d.T2DError) as e:
        t2d.Context(chart, neigh_mode=t2d.NEIGH_EUCLID, capacity=16)
    assert "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly(monkeypatch):
    lib_mod = importlib.import_module("2dtissue_b200._lib")
    monkeypatch.setattr(lib_mod, "_lib", None)
    monkeypatch.setattr(lib_mod, "LIB_PATH", "/nonexistent/lib2dtissue_b200.so")
    with pytest.raises(RuntimeError) as e:
        lib_mod.load()
    assert "no CPU fallback" in str(e.value)


def test_chart_roundtrip_and_refine(tmp_path, chart):
    t2d = importlib.import_module("2dtissue_b200")
    import numpy as np
    p = tmp_path / "c.t2dchart"
    t2d.save_chart(str(p), chart)
    c2 = t2d.load_chart(str(p))
    for k in ("uv", "x3d", "faces", "polygon"):
        assert np.array_equal(chart[k], c2[k])
    r = t2d.refine_chart(chart, 1)
    assert len(r["faces"]) == 4 * len(chart["faces"])
    assert len(r["uv"]) == len(r["x3d"]) and len(r["uv"]) > len(chart["uv"])
    assert np.array_equal(r["uv"][:len(chart["uv"])], chart["uv"])
    # UV area is preserved, orientation of every child equals its parent's
    def area(c):
        a, b, cc = (c["uv"][c["faces"][:, i]] for i in range(3))
        return 0.5 * ((b[:, 0] - a[:, 0]) * (cc[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (cc[:, 0] - a[:, 0]))
    assert abs(area(r).sum() - area(chart).sum()) < 1e-6
    assert np.array_equal(r["uv"], r["uv"].astype(np.float32).astype(np.float64))
