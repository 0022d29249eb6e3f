"""Pins the oracle (oracle/*.h, *.cpp — CPU restatement of the reference) against the reference's own
known-answer material.  The reference's tests for this path are property tests in f64
(util/src/tests.rs:243-683, `type T = f64` under cfg(test)), a BVH superset test
(mesh_util/src/bounding_volume_hierarchy.rs:280-323) and kernel values
(gpu/src/kernels/test.rs:33-52); each is restated here with the same constants.
The reference's RNG (ChaCha8) is replaced by numpy's Philox with fixed seeds.
"""
import ctypes as C

import numpy as np
import pytest

import oracle.oracle as orc
from squishy_volumes_b200 import cstructs as cs

L = orc.lib()
dp = C.POINTER(C.c_double)


def d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def cm(F):
    """row-major numpy 3x3 -> column-major flat (nalgebra layout)."""
    return np.ascontiguousarray(np.asarray(F, dtype=np.float64).T.reshape(9))


def from_cm(flat):
    return np.asarray(flat, dtype=np.float64).reshape(3, 3).T


# the nine hand-picked position gradients of util/src/tests.rs:176-224
CANONICAL_F = [
    np.eye(3),
    [[0, -1, 0], [1, 0, 0], [0, 0, 1]],
    [[0, 0, -1], [0, 1, 0], [1, 0, 0]],
    [[1, 0, 0], [0, 0, -1], [0, 1, 0]],
    [[3, 0, 0], [0, 2, 0], [0, 0, 1]],
    [[1, 0, 0], [0, 2, 0], [0, 0, 1]],
    [[0, -1, 0], [1, 0, 0], [0, 0, 2]],
    [[0, -1, 0], [2, 0, 0], [0, 0, 1]],
    [[0, -2, 0], [1, 0, 0], [0, 0, 1]],
]


def position_gradients(n, seed=7):
    """util/src/tests.rs:226-240: random F with 0.1 < |det| < 10, sign flipped to det > 0."""
    rng = np.random.Generator(np.random.Philox(seed))
    out = [np.asarray(f, dtype=np.float64) for f in CANONICAL_F]
    while len(out) < n + len(CANONICAL_F):
        F = rng.random((3, 3))
        dt = abs(np.linalg.det(F))
        if not (1e-1 < dt < 1e1):
            continue
        if np.linalg.det(F) < 0:
            F = -F
        out.append(F)
    return out


LAME = [(10000.0, 0.3), (1000000.0, 0.3), (10000.0, 0.0), (0.0, 0.4)]   # util/src/elastic.rs:679-689
INVISCID = [(100.0, 2), (1000.0, 2), (100.0, 7), (1000.0, 7)]            # util/src/elastic.rs:691-693


def lame(E, nu):
    return L.svo_lame_mu(E, nu), L.svo_lame_lambda(E, nu)


def test_kernel_known_answers():
    # gpu/src/kernels/test.rs:33-36 (values -1, -0.5, 0, 0.5, 1) + cpu/src/kernels.rs:10-38 closed forms
    xs = [-1.0, -0.5, 0.0, 0.5, 1.0]
    quad = [0.125, 0.5, 0.75, 0.5, 0.125]
    lin = [0.0, 0.5, 1.0, 0.5, 0.0]
    cub = [1 / 6, 0.5 * 0.125 - 0.25 + 2 / 3, 2 / 3, 0.5 * 0.125 - 0.25 + 2 / 3, 1 / 6]
    for x, q, l, c in zip(xs, quad, lin, cub):
        assert L.svo_kernel_quadratic(x) == pytest.approx(q, abs=1e-7)
        assert L.svo_kernel_linear(x) == pytest.approx(l, abs=1e-7)
        assert L.svo_kernel_cubic(x) == pytest.approx(c, abs=1e-6)
    assert L.svo_kernel_quadratic(1.5) == 0.0 and L.svo_kernel_quadratic(-1.5) == 0.0
    assert L.svo_kernel_quadratic(7.0) == 0.0


def test_kernel_partition_of_unity_and_shift():
    # cpu/src/kernels.rs:46-49,86-92: the 3 stencil weights from the base node sum to 1
    rng = np.random.Generator(np.random.Philox(42))
    pos = ((rng.random((1000, 3), dtype=np.float32) * 4 - 2) % 4).astype(np.float32)
    h = np.float32(0.37)
    shift = orc.shift_quadratic(pos, float(h))
    norm = pos / h
    assert np.array_equal(shift, np.floor(norm - np.float32(0.5)).astype(np.int32))
    for p in range(0, 1000, 50):
        for a in range(3):
            s = sum(L.svo_kernel_quadratic(float(np.float32(shift[p, a] + i) - norm[p, a])) for i in range(3))
            assert s == pytest.approx(1.0, abs=2e-6)


def test_collider_bits_truth_table():
    # util/src/collider_bits.rs:9-38
    assert L.svo_bits_get(0, 3) == -1
    b = L.svo_bits_set(0, 3, 1)
    assert b == (0x00010001 << 3) and L.svo_bits_get(b, 3) == 1
    b = L.svo_bits_set(b, 3, 0)
    assert b == (0x00010000 << 3) and L.svo_bits_get(b, 3) == 0
    assert L.svo_bits_set(b, 3, -1) == 0
    near0_a, near0_b = L.svo_bits_set(0, 0, 0), L.svo_bits_set(0, 0, 1)
    assert L.svo_bits_compatible(near0_a, near0_a) == 1
    assert L.svo_bits_compatible(near0_a, near0_b) == 0      # opposite sides of collider 0
    assert L.svo_bits_compatible(near0_a, 0) == 1            # "far" is compatible with both sides
    assert L.svo_bits_compatible(0, near0_b) == 1
    c5 = L.svo_bits_set(near0_a, 5, 1)
    assert L.svo_bits_compatible(c5, near0_a) == 1 and L.svo_bits_compatible(c5, near0_b) == 0
    for c in range(16):   # exhaustive per collider
        for s1 in (-1, 0, 1):
            for s2 in (-1, 0, 1):
                a, b2 = L.svo_bits_set(0, c, s1), L.svo_bits_set(0, c, s2)
                assert L.svo_bits_compatible(a, b2) == (0 if (s1 >= 0 and s2 >= 0 and s1 != s2) else 1)


def test_lame_parameters():
    # util/src/elastic.rs:52-64
    mu, lam = lame(10000.0, 0.3)
    assert mu == pytest.approx(10000 / 2 / 1.3) and lam == pytest.approx(10000 * 0.3 / 1.3 / 0.4)
    assert lame(0.0, 0.4) == (0.0, 0.0)


def fd_gradient(energy, F, h):
    g = np.zeros((3, 3))
    for r in range(3):
        for c in range(3):
            Fp, Fm = F.copy(), F.copy()
            Fp[r, c] += h
            Fm[r, c] -= h
            g[r, c] = (energy(Fp) - energy(Fm)) / (2 * h)
    return g


def test_first_piola_neo_hookean_is_energy_gradient():
    # util/src/tests.rs:310-329
    for E, nu in LAME:
        mu, lam = lame(E, nu)
        for F in position_gradients(60):
            if np.linalg.det(F) <= 0:
                continue
            P = np.zeros(9)
            L.svo_stress_neo_hookean(mu, lam, d(cm(F))[1], d(P)[1])
            Pm = np.zeros(9)
            a, ap = d(Pm)
            L.svo_stress_neo_hookean(mu, lam, d(cm(F))[1], ap)
            g = fd_gradient(lambda X: L.svo_energy_neo_hookean(mu, lam, d(cm(X))[1]), np.asarray(F, float), 1e-6)
            assert np.allclose(from_cm(a), g, rtol=1e-4, atol=1e-4 * max(1.0, mu + lam))


def test_first_piola_inviscid_is_energy_gradient():
    # util/src/tests.rs (inviscid counterpart :560-600)
    for K, ex in INVISCID:
        for F in position_gradients(60):
            if np.linalg.det(F) <= 0:
                continue
            a, ap = d(np.zeros(9))
            L.svo_stress_inviscid(K, ex, d(cm(F))[1], ap)
            g = fd_gradient(lambda X: L.svo_energy_inviscid(K, ex, d(cm(X))[1]), np.asarray(F, float), 1e-6)
            scale = max(1.0, float(np.max(np.abs(g))))
            assert np.allclose(from_cm(a), g, rtol=1e-4, atol=1e-5 * scale)


def oracle_svd(F):
    U, S, V = np.zeros(9), np.zeros(3), np.zeros(9)
    L.svo_svd3(d(cm(F))[1], U.ctypes.data_as(dp), S.ctypes.data_as(dp), V.ctypes.data_as(dp))
    return from_cm(U), S, from_cm(V)


def test_svd_matches_lapack():
    # gpu/src/test_svd/test.rs:15-58 compares U V^T and sigma against nalgebra; here against LAPACK
    for F in position_gradients(300):
        F = np.asarray(F, float)
        U, S, V = oracle_svd(F)
        assert np.all(S[:-1] >= S[1:]) and np.all(S >= 0)
        assert np.allclose(U @ np.diag(S) @ V.T, F, atol=1e-12)
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-12) and np.allclose(V.T @ V, np.eye(3), atol=1e-12)
        u, s, vt = np.linalg.svd(F)
        assert np.allclose(S, s, rtol=1e-12, atol=1e-13)
        assert np.allclose(U @ V.T, u @ vt, atol=1e-9)


def test_svd_form_equals_direct_stress():
    # util/src/tests.rs:331-356 and :638-665: P(F) == U diag(P_hat(sigma)) V^T, eps 1e-5
    for E, nu in LAME:
        mu, lam = lame(E, nu)
        for F in position_gradients(100):
            F = np.asarray(F, float)
            if np.linalg.det(F) <= 0:
                continue
            U, S, V = oracle_svd(F)
            a, ap = d(np.zeros(9))
            L.svo_stress_neo_hookean(mu, lam, d(cm(F))[1], ap)
            sd, sdp = d(np.zeros(3))
            L.svo_stress_neo_hookean_svd_diag(mu, lam, d(S)[1], sdp)
            assert np.allclose(U @ np.diag(sd) @ V.T, from_cm(a), rtol=1e-5, atol=1e-5 * max(1.0, mu + lam))
    for K, ex in INVISCID:
        for F in position_gradients(100):
            F = np.asarray(F, float)
            if np.linalg.det(F) <= 0:
                continue
            U, S, V = oracle_svd(F)
            a, ap = d(np.zeros(9))
            L.svo_stress_inviscid(K, ex, d(cm(F))[1], ap)
            sd, sdp = d(np.zeros(3))
            L.svo_stress_inviscid_svd_diag(K, ex, d(S)[1], sdp)
            scale = max(1.0, float(np.max(np.abs(from_cm(a)))))
            assert np.allclose(U @ np.diag(sd) @ V.T, from_cm(a), rtol=1e-5, atol=1e-7 * scale)


def test_second_derivatives_are_jacobians():
    # util/src/tests.rs:405-430 (neo-hookean) and :667-683 (inviscid): d P_hat / d sigma by finite differences
    def jac(first, S, h=1e-6):
        J = np.zeros((3, 3))
        for c in range(3):
            sp, sm = S.copy(), S.copy()
            sp[c] += h
            sm[c] -= h
            J[:, c] = (first(sp) - first(sm)) / (2 * h)
        return J

    for E, nu in LAME:
        mu, lam = lame(E, nu)
        for F in position_gradients(60):
            S = np.linalg.svd(np.asarray(F, float), compute_uv=False)

            def first(s):
                o, op = d(np.zeros(3))
                L.svo_stress_neo_hookean_svd_diag(mu, lam, d(s)[1], op)
                return o.copy()
            H, Hp = d(np.zeros(9))
            L.svo_second_neo_hookean_svd_diag(mu, lam, d(S)[1], Hp)
            J = jac(first, S)
            assert np.allclose(from_cm(H), J, rtol=1e-5, atol=1e-5 * max(1.0, float(np.max(np.abs(J)))))
    for K, ex in INVISCID:
        for F in position_gradients(60):
            S = np.linalg.svd(np.asarray(F, float), compute_uv=False)

            def first(s):
                o, op = d(np.zeros(3))
                L.svo_stress_inviscid_svd_diag(K, ex, d(s)[1], op)
                return o.copy()
            H, Hp = d(np.zeros(9))
            L.svo_second_inviscid_svd_diag(K, ex, d(S)[1], Hp)
            J = jac(first, S)
            assert np.allclose(from_cm(H), J, rtol=1e-5, atol=1e-5 * max(1.0, float(np.max(np.abs(J)))))


def test_viscosity_stress():
    # util/src/elastic.rs:669-688 over the 8x8 parameter grid of :695-697
    rng = np.random.Generator(np.random.Philox(3))
    for i in range(8):
        for j in range(8):
            Cm = rng.random((3, 3)) - 0.5
            o, op = d(np.zeros(9))
            L.svo_viscosity_stress(float(i), float(j), d(cm(Cm))[1], op)
            want = 2 * i * 0.5 * (Cm + Cm.T) + j * np.trace(Cm) * np.eye(3)
            assert np.allclose(from_cm(o), want, atol=1e-12)


def random_triangles(n, seed=420):
    # mesh_util/src/bounding_volume_hierarchy.rs:245-278: n triangles of size ~1 in [-20, 20]^3
    rng = np.random.Generator(np.random.Philox(seed))
    base = rng.random((n, 1, 3)) * 40 - 20
    return (base + rng.random((n, 3, 3)) * 2 - 1).astype(np.float32)


def test_bvh_query_is_superset_of_brute_force():
    # mesh_util/src/bounding_volume_hierarchy.rs:280-323 (leaf_size 1, margin 1, threshold 4, 1000 queries in +-25)
    tris = random_triangles(1000)
    leaf, margin = 1.0, 1.0
    bvh = L.svo_bvh_build(1000, cs.fptr(np.ascontiguousarray(tris.reshape(-1))), C.c_float(leaf), C.c_float(margin), 4)
    try:
        mn = np.floor((tris.min(axis=1) - margin) / leaf)
        mx = np.ceil((tris.max(axis=1) + margin) / leaf)
        rng = np.random.Generator(np.random.Philox(666))
        out = np.zeros(1000, np.uint32)
        for _ in range(1000):
            p = rng.random(3) * 50 - 25
            q = np.floor(p / leaf).astype(np.int32)
            subset = np.nonzero(np.all(mn * leaf <= p, axis=1) & np.all(p <= mx * leaf, axis=1))[0]
            k = L.svo_bvh_query(bvh, q.ctypes.data_as(cs.c_i32p), cs.uptr(out), 1000)
            got = set(out[:k].tolist())
            assert set(subset.tolist()) <= got
    finally:
        L.svo_bvh_destroy(bvh)


def brute_distance(p, a, b, c, samples=60):
    u = np.linspace(0, 1, samples)
    U, V = np.meshgrid(u, u)
    keep = U + V <= 1
    pts = a + U[keep][:, None] * (b - a) + V[keep][:, None] * (c - a)
    return float(np.min(np.linalg.norm(pts - p, axis=1)))


def test_distance_to_triangle():
    # mesh_util/src/mesh.rs:277-309 against dense sampling of the triangle
    rng = np.random.Generator(np.random.Philox(11))
    for _ in range(200):
        a, b, c = (rng.random(3).astype(np.float32) * 2 - 1 for _ in range(3))
        n = np.cross(b - a, c - a)
        if np.linalg.norm(n) < 1e-2:
            continue
        n = (n / np.linalg.norm(n)).astype(np.float32)
        p = (rng.random(3) * 4 - 2).astype(np.float32)
        got = L.svo_distance_to_triangle(cs.fptr(p), cs.fptr(a), cs.fptr(b), cs.fptr(c), cs.fptr(n))
        want = brute_distance(p.astype(float), a.astype(float), b.astype(float), c.astype(float))
        assert got <= want + 1e-5 and got >= want - 0.05
