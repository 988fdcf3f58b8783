"""The product's __host__ __device__ arithmetic (2dtissue_b200/csrc/hd_math.cuh) compiled for the CPU and checked
against the golden vectors / the oracle.  This is the no-GPU safety net for the kernels' math; the GPU parity
tests proper are in test_gpu_parity.py."""
import ctypes as C

import numpy as np

from conftest import golden


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_seam_reentry_matches_reference(hd_lib):
    k = golden("kat_tiling.npz")
    old, new, n = k["old"].copy(), k["new"].copy(), k["n"].copy()
    wraps = C.c_longlong(0)
    caps = hd_lib.hd_seam(n.size, _d(old), _d(new), n.ctypes.data_as(C.POINTER(C.c_int)), C.byref(wraps))
    assert caps == 0
    assert np.array_equal(new, k["out_new"]) and np.array_equal(n, k["out_n"]) and np.array_equal(old, k["out_old"])


def test_point_triangle_and_lift_match_oracle(hd_lib, oracle_mod, chart):
    L = oracle_mod.lib()
    rng = np.random.default_rng(3)
    uv, faces, x3d = chart["uv"], chart["faces"], chart["x3d"]
    for _ in range(3000):
        f = faces[rng.integers(0, len(faces))]
        a, b, c = (np.ascontiguousarray(uv[v]) for v in f)
        ctr = (a + b + c) / 3
        p = np.ascontiguousarray(ctr + rng.normal(0, 0.02, 2))
        d1 = hd_lib.hd_point_triangle_distance(_d(p), _d(a), _d(b), _d(c))
        d2 = L.t2do_point_triangle_distance(_d(p), _d(a), _d(b), _d(c))
        assert d1 == d2


def test_lift_matches_reference_golden(hd_lib, oracle, chart):
    g = golden("get_r3d_N4306.npz")
    uvp = g["uv"]
    N = uvp.size // 2
    _, _, face = oracle.get_r3d(uvp)
    uv, faces, x3d = chart["uv"], chart["faces"], chart["x3d"]
    X = np.zeros(3)
    for i in range(0, N, 7):
        f = faces[face[i]]
        p = np.array([uvp[i], uvp[N + i]])
        which = hd_lib.hd_lift(_d(p), _d(np.ascontiguousarray(uv[f[0]])), _d(np.ascontiguousarray(uv[f[1]])),
                               _d(np.ascontiguousarray(uv[f[2]])), _d(np.ascontiguousarray(x3d[f[0]])),
                               _d(np.ascontiguousarray(x3d[f[1]])), _d(np.ascontiguousarray(x3d[f[2]])), _d(X))
        assert f[which] == g["vid"][i]
        assert X[0] == g["r3d"][i] and X[1] == g["r3d"][N + i] and X[2] == g["r3d"][2 * N + i]


def test_philox_matches_oracle(hd_lib, oracle_mod):
    L = oracle_mod.lib()
    for seed, step, pid in [(0, 0, 0), (1234, 7, 99), (2 ** 40 + 5, 2 ** 33 + 1, 2 ** 31 + 3), (42, 1, 1)]:
        a = hd_lib.hd_philox_uniform(seed, step, pid)
        b = L.t2do_philox_uniform(seed, step, pid)
        assert a == b and 0.0 <= a < 1.0
        assert hd_lib.hd_noise_deg(0.3 * 360.0, seed, step, pid) == L.t2do_noise_deg(0.3, seed, step, pid)
    # Philox4x32-10 known answer (Random123 kat_vectors: counter 0, key 0)
    u = hd_lib.hd_philox_uniform(0, 0, 0)
    bits = int(u * 2 ** 53)
    assert bits == ((0x6627e8d5 << 32) | 0xe169c58d) >> 11


def test_inside_and_fij(hd_lib, oracle_mod):
    z = golden("inside_pins.npz")
    uv, ins = z["uv"], z["inside"]
    N = ins.size
    got = np.array([hd_lib.hd_inside(float(uv[i]), float(uv[N + i])) for i in range(N)])
    assert np.array_equal(got, ins)
    out = np.zeros(2)
    oracle_mod.lib().t2do_repulsive_adhesion(10.0, 1.4166666666666667, 0.953489, 1.0, 0.75, 1.0, 0.0, _d(out))
    assert hd_lib.hd_pair_fij(10.0, 2 * 1.4166666666666667, 0.953489) / 0.953489 == out[0] or \
        abs(hd_lib.hd_pair_fij(10.0, 2 * 1.4166666666666667, 0.953489) * (1.0 / 0.953489) - out[0]) < 1e-15


def test_mean_angle_correctly_rounded_matches_glibc_pipeline(hd_lib):
    """The fp64 heading path rebuilds atan2 correctly rounded (hd_math.cuh mean_angle_degrees_cr) because the
    reference truncates the mean angle to int and aligned / isolated particles sit exactly on integer degrees.
    Exhaustive over the tie families + 2M random neighbour sets: the truncated heading must equal the glibc
    pipeline's (what the reference executes) in every case."""
    hd_lib.hd_cr_selfcheck.restype = C.c_longlong
    for family in (0, 1, 2):
        sets, ties, val = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        bad = hd_lib.hd_cr_selfcheck(family, C.byref(sets), C.byref(ties), C.byref(val))
        print("family %d: %d sets, %d ties, %d angle doubles differ, %d headings differ" %
              (family, sets.value, ties.value, val.value, bad))
        assert bad == 0
        if family < 2:
            assert ties.value > 0.5 * sets.value


def test_fast_path_pair_algebra_matches_the_reference_formula(oracle_mod):
    """The fp32 step kernel evaluates F_ij / d as ONE FMA on 1/d:  (-k)/d + k/(2 sigma)  (k_step_euclid_fast, pair_term).
    Same number as the reference's  [-k (2 sigma - d) / (2 sigma)] / d  (ForceHelper.cpp:84-104, via the oracle's
    t2do_repulsive_adhesion), including the d == 0 -> 0.001 rule (ForceHelper.cpp:59-62), to fp32 accuracy."""
    L = oracle_mod.lib()
    rng = np.random.default_rng(4)
    k, sigma = 1.0, 0.05
    d = np.concatenate([rng.random(2000) * 2 * sigma, [1e-7, 1e-5, 0.001, 2 * sigma * (1 - 1e-7)]])
    duv = rng.normal(size=(d.size, 2)) * 1e-3
    for dist, (ux, uy) in zip(d, duv):
        out = (C.c_double * 2)()
        # oracle: F_ij * dist_v / dist  with dist_v = (ux, uy)
        L.t2do_repulsive_adhesion(C.c_double(k), C.c_double(sigma), C.c_double(dist), C.c_double(1.0), C.c_double(0.75),
                                  C.c_double(ux), C.c_double(uy), out)
        rinv = np.float32(1.0) / np.float32(dist)
        g = np.float32(rinv * np.float32(-k) + np.float32(k / (2 * sigma)))
        fx, fy = float(g * np.float32(ux)), float(g * np.float32(uy))
        scale = max(abs(out[0]), abs(out[1]), 1e-12)
        assert abs(fx - out[0]) <= 2e-5 * scale + 1e-9 and abs(fy - out[1]) <= 2e-5 * scale + 1e-9
    # d == 0: the rule gives 1/d = 1000
    g0 = np.float32(1000.0) * np.float32(-k) + np.float32(k / (2 * sigma))
    ref = (-k * (2 * sigma - 0.001) / (2 * sigma)) / 0.001
    assert abs(float(g0) - ref) <= 1e-5 * abs(ref)
