"""Known-answer tests of the CPU oracle against the golden vectors held by the reference's OWN tests
(SURVEY.md §4 / §8c).  File:line citations are under /root/reference."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def test_angles_to_unit_vectors_kat(oracle_mod):
    # tests/simulation/test_LinearAlgebra.cpp:13-40
    L = oracle_mod.lib()
    n = np.array([0, 45, 90, 180, 270, 360], dtype=np.int32)
    out = np.zeros(12)
    L.t2do_angles_to_unit_vectors(6, _i(n), _d(out))
    exp = np.array([[1, 0], [np.sqrt(2) / 2, np.sqrt(2) / 2], [0, 1], [-1, 0], [0, -1], [1, 0]])
    assert np.allclose(out[:6], exp[:, 0], atol=1e-9)
    assert np.allclose(out[6:], exp[:, 1], atol=1e-9)


def test_euclidean_tiling_kat(oracle):
    # tests/simulation/test_EuclideanTiling.cpp:44-72 (crashes in the stock build: empty borders)
    old = np.array([0.5, 0.5, 0.9, 0.5, 0.5, 0.4])
    new = np.array([2.5, 1.3, -0.3, 0.5, 1.2, 0.7])
    n = np.array([80, 120, 42], dtype=np.int32)
    _, uv, nn = oracle.tiling(old, new, n)
    expected = np.array([0.5, 0.7, 0.7, 0.5, 0.8, 0.3])   # column-major of {0.5,0.5},{0.7,0.8},{0.7,0.3}
    assert np.allclose(uv, expected, atol=1e-9)
    # headings produced by the compiled reference with populated borders (tests/golden/kat_tiling.npz)
    k = golden("kat_tiling.npz")
    assert np.array_equal(nn, k["kat_out_n"]) and list(nn) == [-280, -60, -48]
    assert np.array_equal(uv, k["kat_out_new"])


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_repulsive_adhesion_kat(oracle_mod, sign):
    # tests/simulation/test_Locomotion.cpp:42-60 (commented out upstream, vectors still valid)
    L = oracle_mod.lib()
    out = np.zeros(2)
    L.t2do_repulsive_adhesion(10.0, 1.4166666666666667, 0.953489, 1.0, 0.75, sign * 0.0203217, sign * 0.010791, _d(out))
    assert abs(out[0] - (-sign * 0.141406)) < 1e-5
    assert abs(out[1] - (-sign * 0.075088)) < 1e-5


def test_mean_angle_kats(oracle_mod):
    # tests/simulation/test_Locomotion.cpp:63-82
    L = oracle_mod.lib()
    empty = C.c_int(0)
    r = L.t2do_mean_angle_deg(None, 0, C.byref(empty))
    assert empty.value == 1 and np.isnan(r)          # reference throws std::invalid_argument
    a = np.array([0.0, 180.0, 90.0])
    assert abs(L.t2do_mean_angle_deg(_d(a), 3, None) - 90.0) < 1e-5
    a = np.array([-45.0, -90, -135, -180, -225, -270, -315])
    assert abs(L.t2do_mean_angle_deg(_d(a), 7, None) - 180.0) < 1e-5


def test_symmetrize_kats(oracle_mod):
    # tests/simulation/test_Locomotion.cpp:87-136
    L = oracle_mod.lib()
    for src, exp in [([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1, 2, 3], [2, 5, 6], [3, 6, 9]]),
                     ([[1, 0, 3], [4, 5, 6], [7, 8, 0]], [[1, 0, 3], [0, 5, 6], [3, 6, 0]]),
                     ([[0, 0, 0]] * 3, [[0, 0, 0]] * 3)]:
        A = np.array(src, dtype=np.float64)
        L.t2do_symmetrize_min(3, _d(A))
        assert np.array_equal(A, np.array(exp, dtype=np.float64))


def test_get_dist_vect_kats(oracle_mod):
    # tests/simulation/test_Locomotion.cpp:141-263
    L = oracle_mod.lib()

    def run(r):
        r = np.asarray(r, dtype=np.float64)
        N = len(r)
        col = np.concatenate([r[:, 0], r[:, 1]])
        dx, dy = np.zeros(N * N), np.zeros(N * N)
        L.t2do_get_dist_vect(N, _d(col), _d(dx), _d(dy))
        return dx.reshape(N, N), dy.reshape(N, N)

    dx, dy = run([[1, 2], [3, 4], [5, 6]])
    e = np.array([[0, -2, -4], [2, 0, -2], [4, 2, 0]], dtype=float)
    assert np.array_equal(dx, e) and np.array_equal(dy, e)
    dx, dy = run([[0, 0]] * 3)
    assert not dx.any() and not dy.any()
    dx, dy = run([[1, 2]])
    assert dx.shape == (1, 1) and dx[0, 0] == 0 and dy[0, 0] == 0
    dx, dy = run([[1.0, 2.0], [3.0, 4.0]])
    assert np.allclose(dx, [[0, -2], [2, 0]], atol=1e-5) and np.allclose(dy, [[0, -2], [2, 0]], atol=1e-5)
    r10 = [[0.448453, 0.365021], [0.252378, 0.477139], [0.0309307, 0.327166], [0.903785, 0.160117], [0.257268, 0.436529],
           [0.289008, 0.49713], [0.844151, 0.209798], [0.268783, 0.352196], [0.968116, 0.673458], [0.188375, 0.567491]]
    dx, dy = run(r10)
    row0x = [0, 0.196075, 0.417522, -0.455332, 0.191185, 0.159445, -0.395698, 0.17967, -0.519663, 0.260078]
    row0y = [0, -0.112118, 0.0378543, 0.204904, -0.0715087, -0.132109, 0.155223, 0.0128243, -0.308438, -0.202471]
    assert np.allclose(dx[0], row0x, atol=1e-5) and np.allclose(dy[0], row0y, atol=1e-5)
    assert np.allclose(dx, -dx.T) and np.allclose(dy, -dy.T)


def test_average_n_within_distance_kat(oracle_mod, chart):
    # tests/simulation/test_Locomotion.cpp:279-335: 10 particles, sigma = 1.41667, tolerance +-2 degrees.
    # Rows 2 and 4 are an antipodal pair (290 vs 110): their expected value (21) is numerical noise.
    dl = np.array([
        [0, 1.94061, 5.60903, 8.28046, 8.47736, 11.0131, 14.2291, 6.0693, 12.7292, 10.761],
        [1.94061, 0, 4.51284, 6.31737, 7.23179, 11.6579, 12.934, 5.8962, 10.894, 10.6601],
        [5.60903, 4.51284, 0, 5.55914, 2.81003, 9.00323, 10.8524, 10.1511, 7.80554, 6.77901],
        [8.28046, 6.31737, 5.55914, 0, 5.75308, 14.5522, 6.5117, 7.98363, 5.19999, 11.8869],
        [8.47736, 7.23179, 2.81003, 5.75308, 0, 9.30527, 9.01998, 12.5087, 5.71498, 6.17054],
        [11.0131, 11.6579, 9.00323, 14.5522, 9.30527, 0, 9.45922, 14.2346, 12.4246, 3.59292],
        [14.2291, 12.934, 10.8524, 6.5117, 9.01998, 9.45922, 0, 10.0873, 3.28317, 12.8653],
        [6.0693, 5.8962, 10.1511, 7.98363, 12.5087, 14.2346, 10.0873, 0, 11.829, 16.5712],
        [12.7292, 10.894, 7.80554, 5.19999, 5.71498, 12.4246, 3.28317, 11.829, 0, 11.2371],
        [10.761, 10.6601, 6.77901, 11.8869, 6.17054, 3.59292, 12.8653, 16.5712, 11.2371, 0]])
    n = np.array([168, 154, 290, 83, 110, 46, 48, 144, 227, 48], dtype=np.int32)
    expected = np.array([161, 161, 21, 83, 21, 46, 48, 144, 227, 48])
    sigma = 1.4166666666666667
    L = oracle_mod.lib()
    got = np.zeros(10, dtype=int)
    for i in range(10):
        ang = np.array([float(n[j]) for j in range(10) if dl[i, j] < 2 * sigma])
        got[i] = int(L.t2do_mean_angle_deg(_d(ang), len(ang), None))
    keep = [0, 1, 3, 5, 6, 7, 8, 9]
    assert np.all(np.abs(got[keep] - expected[keep]) <= 2)
    # the oracle's step uses the same function through a table: feed dl as a 10-vertex "table"
    o = oracle_mod.Oracle(dict(uv=chart["uv"], x3d=chart["x3d"], faces=chart["faces"]))
    Dfull = np.full((o.V, o.V), 1e9)
    np.fill_diagonal(Dfull, 0.0)
    Dfull[:10, :10] = dl
    o.set_table(Dfull)
    uv = np.concatenate([chart["uv"][:10, 0], chart["uv"][:10, 1]]) * 0.5 + 0.25
    r3d, _, _ = o.get_r3d(uv)
    res = o.step(uv, n, np.arange(10, dtype=np.int32), r3d, 0.0, 10.0, sigma, 0.0, mode=0, brute=True)
    assert np.all(np.abs(res["n"][keep] - expected[keep]) <= 2)
    res2 = o.step(uv, n, np.arange(10, dtype=np.int32), r3d, 0.0, 10.0, sigma, 0.0, mode=0, brute=False)
    assert np.array_equal(res["n"], res2["n"]) and np.array_equal(res["F"], res2["F"])


def test_inside_predicate_pins(oracle_mod):
    # MeshCartographyLib/tests/test_SurfaceParametrization.cpp:59-123: check_point_in_polygon == closed unit square;
    # inside_pins.npz holds the compiled polygon test's answers on random + boundary points
    L = oracle_mod.lib()
    z = golden("inside_pins.npz")
    uv, ins = z["uv"], z["inside"]
    N = ins.size
    got = np.array([L.t2do_inside(float(uv[i]), float(uv[N + i])) for i in range(N)])
    assert np.array_equal(got, ins)


def test_barycentric_lift_extension(chart):
    """T2D_LIFT_BARYCENTRIC in the oracle: the lifted point lies in the plane of its 3-D triangle, the
    weights reproduce the corners, and the default mode is untouched (the golden tests above run with it)."""
    from oracle import oraclebind
    orc = oraclebind.Oracle(chart)
    rng = np.random.default_rng(3)
    N = 5000
    uv = rng.random(2 * N)
    r_ref, vid_ref, face_ref = orc.get_r3d(uv)
    orc.set_lift_mode(1)
    r_b, vid_b, face_b = orc.get_r3d(uv)
    orc.set_lift_mode(0)
    assert np.array_equal(face_b, face_ref)                       # point location does not depend on the lift
    faces, x3d, uvv = chart["faces"], chart["x3d"], chart["uv"]
    A, B, Cc = (x3d[faces[face_b][:, k]] for k in range(3))
    P = np.stack([r_b[:N], r_b[N:2 * N], r_b[2 * N:]], 1)
    nrm = np.cross(B - A, Cc - A)
    dist = np.abs(np.einsum("ij,ij->i", P - A, nrm)) / np.linalg.norm(nrm, axis=1)
    assert dist.max() < 1e-9                                      # in the triangle's plane
    # independent barycentric evaluation in UV
    a, b, c = (uvv[faces[face_b][:, k]] for k in range(3))
    p = np.stack([uv[:N], uv[N:]], 1)
    T = np.stack([b - a, c - a], 2)
    lam = np.linalg.solve(T, (p - a)[:, :, None])[:, :, 0]
    Q = A + lam[:, :1] * (B - A) + lam[:, 1:] * (Cc - A)
    assert np.max(np.abs(Q - P)) < 1e-9
    # the reference's lift is a different map (it pulls points towards the middle of the face)
    Pr = np.stack([r_ref[:N], r_ref[N:2 * N], r_ref[2 * N:]], 1)
    assert np.max(np.linalg.norm(Pr - P, axis=1)) > 1e-3
    # corners map to corners
    v = faces[100]
    cu = np.concatenate([uvv[v][:, 0], uvv[v][:, 1]])
    orc.set_lift_mode(1)
    rc, _, _ = orc.get_r3d(cu)
    orc.set_lift_mode(0)
    Pc = np.stack([rc[:3], rc[3:6], rc[6:]], 1)
    assert min(np.min(np.linalg.norm(x3d - Pc[k], axis=1)) for k in range(3)) < 1e-9
