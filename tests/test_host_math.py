"""CPU checks of the PRODUCT's math header and host code (squishy_volumes_b200/csrc/svb_math.cuh,
svb_host.h) compiled for the host through tests/native/math_shim.cpp: the same source the CUDA
kernels inline.  Compared with numpy/LAPACK and with the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle.oracle as orc
from squishy_volumes_b200 import cstructs as cs, scenes
from squishy_volumes_b200.types import ParticleFlags
from tests.test_oracle_pins import INVISCID, LAME, position_gradients, random_triangles

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "math_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, os.path.join(HERE, "native", "math_shim.cpp")], check=True)
    S = C.CDLL(out)
    S.shim_kernel_quadratic.restype = C.c_float
    S.shim_kernel_quadratic.argtypes = [C.c_float]
    S.shim_bits_set.restype = C.c_uint32
    S.shim_bits_set.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    S.shim_bits_get.argtypes = [C.c_uint32, C.c_uint32]
    S.shim_bits_compatible.argtypes = [C.c_uint32, C.c_uint32]
    S.shim_svd3.argtypes = [cs.c_f32p] * 4
    S.shim_return_map.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, cs.c_f32p, cs.c_f32p]
    S.shim_stress.argtypes = [C.c_int, C.c_float, C.c_float, cs.c_f32p, cs.c_f32p]
    S.shim_viscous.argtypes = [C.c_float, C.c_float, cs.c_f32p, cs.c_f32p]
    S.shim_limits.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, cs.c_f32p, C.c_float, cs.c_f32p]
    S.shim_mesh_build.restype = C.c_void_p
    S.shim_mesh_build.argtypes = [C.c_uint32, cs.c_u32p, cs.c_u32p, cs.c_u32p, cs.c_f32p, cs.c_f32p, C.c_float, C.c_float, C.c_uint32, C.POINTER(C.c_int)]
    S.shim_mesh_destroy.argtypes = [C.c_void_p]
    S.shim_mesh_topology.argtypes = [C.c_void_p, cs.c_u32p, cs.c_u32p, cs.c_u32p, cs.c_u32p]
    S.shim_mesh_level.argtypes = [C.c_void_p]
    S.shim_mesh_query.restype = C.c_uint64
    S.shim_mesh_query.argtypes = [C.c_void_p, cs.c_i32p, cs.c_u32p, C.c_uint64]
    return S


def f32cm(F):
    return np.ascontiguousarray(np.asarray(F, dtype=np.float32).T.reshape(9))


def test_kernel_and_bits_match_oracle(shim):
    L = orc.lib()
    for x in np.linspace(-2.5, 2.5, 1001, dtype=np.float32):
        assert shim.shim_kernel_quadratic(float(x)) == L.svo_kernel_quadratic(float(x))
    rng = np.random.Generator(np.random.Philox(5))
    for _ in range(2000):
        a, b = (int(v) for v in rng.integers(0, 2 ** 32, 2, dtype=np.uint64))
        c = int(rng.integers(0, 16))
        s = int(rng.integers(-1, 2))
        assert shim.shim_bits_compatible(a, b) == L.svo_bits_compatible(a, b)
        assert shim.shim_bits_set(a, c, s) == L.svo_bits_set(a, c, s)
        assert shim.shim_bits_get(a, c) == L.svo_bits_get(a, c)


def test_device_svd_against_lapack(shim):
    # f32 one-sided Jacobi: sigma, U V^T and the recomposition within a few f32 ulps of LAPACK (f64)
    worst = 0.0
    for F in position_gradients(500):
        Ff = np.asarray(F, np.float32)
        U, S, V = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(9, np.float32)
        shim.shim_svd3(cs.fptr(f32cm(Ff)), cs.fptr(U), cs.fptr(S), cs.fptr(V))
        U, V = U.reshape(3, 3).T.astype(float), V.reshape(3, 3).T.astype(float)
        u, s, vt = np.linalg.svd(Ff.astype(float))
        assert np.all(S[:-1] >= S[1:])
        assert np.allclose(S, s, rtol=2e-6, atol=2e-6 * s[0])
        assert np.allclose(U @ np.diag(S.astype(float)) @ V.T, Ff, atol=4e-6 * s[0])
        assert np.allclose(U.T @ U, np.eye(3), atol=5e-6) and np.allclose(V.T @ V, np.eye(3), atol=5e-6)
        if s[1] - s[2] > 1e-3 * s[0] or True:
            worst = max(worst, float(np.max(np.abs(U @ V.T - u @ vt))))
    assert worst < 2e-4   # U V^T is ill-conditioned only when sigma_min -> 0 (|det| >= 0.1 here)


def test_stress_matches_oracle(shim):
    L = orc.lib()
    dp = C.POINTER(C.c_double)
    for E, nu in LAME:
        mu, lam = L.svo_lame_mu(E, nu), L.svo_lame_lambda(E, nu)
        for F in position_gradients(100):
            if np.linalg.det(F) <= 0.05:
                continue
            P = np.zeros(9, np.float32)
            shim.shim_stress(0, mu, lam, cs.fptr(f32cm(F)), cs.fptr(P))
            Fd = np.ascontiguousarray(np.asarray(F, np.float32).astype(np.float64).T.reshape(9))
            ref = np.zeros(9)
            L.svo_stress_neo_hookean(float(np.float32(mu)), float(np.float32(lam)), Fd.ctypes.data_as(dp), ref.ctypes.data_as(dp))
            assert np.allclose(P, ref, rtol=2e-5, atol=2e-5 * max(1.0, mu + lam))
    for K, ex in INVISCID:
        for F in position_gradients(100):
            if np.linalg.det(F) <= 0.2:
                continue
            P = np.zeros(9, np.float32)
            shim.shim_stress(1, K, float(ex), cs.fptr(f32cm(F)), cs.fptr(P))
            Fd = np.ascontiguousarray(np.asarray(F, np.float32).astype(np.float64).T.reshape(9))
            ref = np.zeros(9)
            L.svo_stress_inviscid(K, ex, Fd.ctypes.data_as(dp), ref.ctypes.data_as(dp))
            assert np.allclose(P, ref, rtol=1e-4, atol=1e-4 * max(1.0, float(np.max(np.abs(ref)))))


def nalgebra_like_return_map(F, mu, lam, alpha):
    """advance_particles.rs:50-71 in f64 with LAPACK's SVD."""
    u, s, vt = np.linalg.svd(F)
    e = np.log(s)
    tr = e.sum()
    eh = e - tr / 3
    n = np.linalg.norm(eh)
    if tr < 0 and n > 0:
        dg = n + (3 * lam + 2 * mu) / 2 / mu * tr * alpha
        if dg > 0:
            return u @ np.diag(np.exp(e - dg / n * eh)) @ vt
        return F
    return u @ vt


def test_sand_and_fluid_return_mapping(shim):
    # gpu/src/sand/test.rs:16-80 and gpu/src/fluid/test.rs: 1 % by norm against nalgebra; here 1e-4 against f64
    mu, lam = 3846.1538, 5769.2308
    for F in position_gradients(400):
        F = np.asarray(F, np.float32).astype(float)
        buf = f32cm(F)
        e = np.zeros(1, np.float32)
        ok = shim.shim_return_map(ParticleFlags.IS_SOLID | ParticleFlags.USE_SAND_ALPHA, mu, lam, 0.3, cs.fptr(buf), cs.fptr(e))
        want = nalgebra_like_return_map(F, mu, lam, 0.3)
        got = buf.reshape(3, 3).T.astype(float)
        assert np.linalg.norm(got - want) <= 2e-4 * max(1.0, np.linalg.norm(want))
        assert ok == (1 if np.linalg.det(got) > 0 else 0)
        buf = f32cm(F)
        shim.shim_return_map(ParticleFlags.IS_FLUID, 1000.0, 7.0, 0.0, cs.fptr(buf), cs.fptr(e))
        u, s, vt = np.linalg.svd(F)
        want = np.cbrt(np.prod(s)) * (u @ vt)
        got = buf.reshape(3, 3).T.astype(float)
        assert np.linalg.norm(got - want) <= 2e-4 * max(1.0, np.linalg.norm(want))


def test_topology_and_bvh_match_oracle(shim):
    L = orc.lib()
    scene = scenes.mixed(side=8, brick=2)     # plane (open fans) + torus (closed manifold)
    fi = scene.frame_input
    o = orc.OracleState.from_io_state(scene.io_state, fi)
    o._sync_keyframes(fi)
    nv, nt, flat = cs.topology_arrays(fi)
    va = np.ascontiguousarray(fi.a().vertex_positions, np.float32)
    vb = np.ascontiguousarray(fi.b().vertex_positions, np.float32)
    h = fi.consts.scaled_grid_node_size()
    err = C.c_int(0)
    m = shim.shim_mesh_build(len(fi.colliders), cs.uptr(nv), cs.uptr(nt), cs.uptr(flat), cs.fptr(va), cs.fptr(vb), C.c_float(np.float32(h) * np.float32(2.2)),
                             C.c_float(fi.consts.leaf_size), fi.consts.leaf_threshold, C.byref(err))
    try:
        assert err.value == 0
        T, V = int(nt.sum()), int(nv.sum())
        tri, opp, col, fan = (np.zeros(T * 3, np.uint32), np.zeros(T * 3, np.uint32), np.zeros(T, np.uint32), np.zeros(V, np.uint32))
        shim.shim_mesh_topology(m, cs.uptr(tri), cs.uptr(opp), cs.uptr(col), cs.uptr(fan))
        tri_o, opp_o, col_o, fan_o = (np.zeros(T * 3, np.uint32), np.zeros(T * 3, np.uint32), np.zeros(T, np.uint32), np.zeros(V, np.uint32))
        L.svo_topology_get(o._h, cs.uptr(tri_o), cs.uptr(opp_o), cs.uptr(col_o), cs.uptr(fan_o))
        assert np.array_equal(tri, tri_o) and np.array_equal(opp, opp_o) and np.array_equal(col, col_o) and np.array_equal(fan, fan_o)
        assert fan[:4].sum() == 0 and np.all(fan[4:] == 6)   # plane corners are open fans; torus vertices have valence 6
        rng = np.random.Generator(np.random.Philox(9))
        a, b = np.zeros(4096, np.uint32), np.zeros(4096, np.uint32)
        lo = np.floor(va.min(axis=0) / fi.consts.leaf_size).astype(int) - 3
        hi = np.ceil(va.max(axis=0) / fi.consts.leaf_size).astype(int) + 3
        hits = 0
        for _ in range(3000):
            q = np.array([rng.integers(lo[k], hi[k] + 1) for k in range(3)], dtype=np.int32)
            ka = shim.shim_mesh_query(m, q.ctypes.data_as(cs.c_i32p), cs.uptr(a), 4096)
            kb = L.svo_handle_bvh_query(o._h, q.ctypes.data_as(cs.c_i32p), cs.uptr(b), 4096)
            assert ka == kb and np.array_equal(a[:ka], b[:kb])
            hits += ka > 0
        assert hits > 50
    finally:
        shim.shim_mesh_destroy(m)


def test_bvh_superset_property_product(shim):
    # the reference's own BVH test (mesh_util/src/bounding_volume_hierarchy.rs:280-323) on the product builder
    tris = random_triangles(1000)
    nv = np.array([3000], np.uint32)
    nt = np.array([1000], np.uint32)
    flat = np.arange(3000, dtype=np.uint32)
    va = np.ascontiguousarray(tris.reshape(-1, 3))
    err = C.c_int(0)
    m = shim.shim_mesh_build(1, cs.uptr(nv), cs.uptr(nt), cs.uptr(flat), cs.fptr(va), None, C.c_float(1.0), C.c_float(1.0), 4, C.byref(err))
    try:
        assert err.value == 0
        mn = np.floor((tris.min(axis=1) - 1.0) / 1.0)
        mx = np.ceil((tris.max(axis=1) + 1.0) / 1.0)
        rng = np.random.Generator(np.random.Philox(666))
        out = np.zeros(1000, np.uint32)
        for _ in range(1000):
            p = rng.random(3) * 50 - 25
            q = np.floor(p).astype(np.int32)
            subset = set(np.nonzero(np.all(mn <= p, axis=1) & np.all(p <= mx, axis=1))[0].tolist())
            k = shim.shim_mesh_query(m, q.ctypes.data_as(cs.c_i32p), cs.uptr(out), 1000)
            assert subset <= set(out[:k].tolist())
    finally:
        shim.shim_mesh_destroy(m)


def test_time_step_limits_match_oracle(shim):
    """particle_time_step_limits (device header) vs the oracle's LimitTimeStepBeforeForce on 1-particle scenes."""
    from squishy_volumes_b200.types import RunParameters
    rng = np.random.Generator(np.random.Philox(21))
    for kind in ("solid", "fluid"):
        for trial in range(12):
            sc = scenes.jelly_collision(side=1) if kind == "solid" else scenes.dam_break(nx=1, ny=1, nz=1)
            p = sc.io_state.particles.select(np.array([0]))
            F = np.eye(3) + (rng.random((3, 3)) - 0.5) * (0.0 if trial == 0 else 0.3)
            p.position_gradients[0] = F.T.astype(np.float32)
            sc.io_state.particles = p
            sc.frame_input.keyframes = [type(k)(gravity=k.gravity) for k in sc.frame_input.keyframes]
            sc.frame_input.colliders = []
            h = sc.frame_input.consts.scaled_grid_node_size()
            o = orc.OracleState.from_io_state(sc.io_state, sc.frame_input)
            o.produce_next_state(None, sc.frame_input, RunParameters(target_time=1e-9, max_time_step=1.0, adaptive_time_steps=True))
            out = np.zeros(2, np.float32)
            shim.shim_limits(int(kind == "fluid"), float(p.mu_or_bulk_modulus[0]), float(p.lambda_or_exponent[0]), float(p.mass[0]), float(p.initial_volume[0]),
                             cs.fptr(np.ascontiguousarray(p.position_gradients[0].reshape(9))), C.c_float(h), cs.fptr(out))
            # the oracle's first substep time step is min(max, by_sound, by_isolated) (by_velocity/deformation not yet set before scatter)
            assert o.time == pytest.approx(min(1.0, float(out.min())), rel=2e-4) or o.time <= min(1.0, float(out.min())) * (1 + 2e-4)
