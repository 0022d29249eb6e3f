"""Whole-substep properties of the oracle (the reference holds NO test of the CPU substep, SURVEY.md
§4; these pin the restatement through physics the algorithm must satisfy) and the golden fixtures
that the GPU tests replay."""
import os

import numpy as np
import pytest

import oracle.oracle as orc
from squishy_volumes_b200 import scenes
from squishy_volumes_b200.types import ParticleFlags, RunParameters
from tests import golden_scenes

HERE = os.path.dirname(os.path.abspath(__file__))


def run(scene, n, adaptive=False, store_grid=True):
    o = orc.OracleState.from_io_state(scene.io_state, scene.frame_input)
    dt = scene.time_step
    st, err = o.produce_next_state(None, scene.frame_input, RunParameters(scene.io_state.time + (n - 0.5) * dt if not adaptive else n * dt, dt, adaptive, store_grid))
    return o, st, err


def test_p2g_conserves_mass_and_momentum():
    # scatter_momentum.rs: sum_nodes m = sum_p m ; with F = I (zero stress) sum_nodes p = sum_p m (v + g dt)
    sc = scenes.jelly_collision(side=8)
    o, st, err = run(sc, 1)
    assert err is None
    g = st.grid_nodes
    p = sc.io_state.particles
    assert g.masses.sum() == pytest.approx(p.mass.sum(), rel=1e-5)
    mom_grid = (g.velocities.astype(np.float64) * g.masses[:, None]).sum(axis=0)
    mom_part = (p.velocities.astype(np.float64) * p.mass[:, None]).sum(axis=0)
    assert np.allclose(mom_grid, mom_part, atol=1e-5 * np.abs(p.mass.sum()))


def test_apic_reproduces_affine_velocity_field():
    # collect_velocity.rs: with v(x) = A x + b carried by (v, C) the round trip P2G -> G2P returns C = A
    # on interior particles (quadratic B-splines reproduce affine fields; APIC transfers are exact for them)
    sc = scenes.jelly_collision(side=14)
    keep = np.nonzero(sc.io_state.particles.positions[:, 0] < 0)[0]      # one block only
    sc.io_state.particles = sc.io_state.particles.select(keep)
    for k in sc.frame_input.keyframes:
        k.particle_flags = None
        k.particle_goal_positions = None
    p = sc.io_state.particles
    A = np.array([[0.0, -0.3, 0.1], [0.3, 0.0, 0.2], [-0.1, -0.2, 0.0]], np.float32)   # divergence free
    b = np.array([0.1, -0.2, 0.3], np.float32)
    p.velocities[:] = p.positions @ A.T + b
    p.velocity_gradients[:] = A.T[None]            # array of columns
    p.mu_or_bulk_modulus[:] = 0                    # no stress: pure transfer
    p.lambda_or_exponent[:] = 0
    sc.time_step = 1e-6
    o, st, err = run(sc, 1)
    q = st.particles
    x0 = sc.io_state.particles.positions
    h = 0.04
    interior = np.all((x0 > x0.min(axis=0) + 2 * h) & (x0 < x0.max(axis=0) - 2 * h), axis=1)   # full stencil surrounded by material
    assert interior.sum() > 50
    want_v = x0 @ A.T + b
    assert np.allclose(q.velocities[interior], want_v[interior], atol=2e-5)
    assert np.allclose(q.velocity_gradients[interior], A.T[None], atol=2e-3)


def test_free_fall_and_fixed_time_step():
    sc = scenes.elastic_cube(side=6, h=0.1)
    sc.frame_input.colliders = []
    for k in sc.frame_input.keyframes:
        k.vertex_positions = None
        k.triangle_frictions = None
        k.triangle_dampings = None
    o, st, err = run(sc, 10)
    assert o.substeps == 10 and o.time == pytest.approx(10 * 1e-3, rel=1e-6)
    # rigid free fall: v = g t for every particle, F stays I
    assert np.allclose(st.particles.velocities[:, 2], -9.8 * 10e-3, rtol=1e-4)
    assert np.allclose(st.particles.position_gradients, np.eye(3)[None], atol=1e-5)


def test_cull_tombstones_and_freezes():
    # cull_particles.rs:31-39: strictly inside the (scaled) domain box or TOMBSTONED; tombstoned particles stop moving
    sc = scenes.elastic_cube(side=4, h=0.1)
    z0 = sc.io_state.particles.positions[:, 2].copy()
    cut = float(np.median(z0))
    sc.frame_input.consts.domain_min = (-100.0, -100.0, cut)
    o, st, err = run(sc, 1)
    tomb = (st.particles.flags & ParticleFlags.TOMBSTONED) != 0
    assert 0 < tomb.sum() < sc.n
    assert np.array_equal(tomb, ~(st.particles.positions[:, 2] > np.float32(cut)))
    o2, st2, _ = run(sc, 3)
    tomb2 = (st2.particles.flags & ParticleFlags.TOMBSTONED) != 0
    assert np.all(tomb2[tomb])
    assert np.array_equal(st2.particles.positions[tomb], st.particles.positions[tomb])


def test_energy_error_is_returned_with_state():
    sc = scenes.jelly_collision(side=4)
    sc.io_state.particles.position_gradients[5] = -np.eye(3, dtype=np.float32)
    o, st, err = run(sc, 1)
    assert err is not None and err.status == 8
    assert st.particles.flags[5] & ParticleFlags.FAILED


def test_wrong_frame_is_fatal():
    from squishy_volumes_b200.types import FatalError
    sc = scenes.jelly_collision(side=4)
    sc.io_state.time = 0.5   # frame 12, but frame 0 is loaded
    o = orc.OracleState.from_io_state(sc.io_state, sc.frame_input)
    with pytest.raises(FatalError):
        o.produce_next_state(None, sc.frame_input, RunParameters(0.6, 1e-3))


@pytest.mark.parametrize("name", sorted(golden_scenes.GOLDEN))
def test_oracle_reproduces_golden(name):
    """tests/golden/*.npz were produced by tests/golden_scenes.py from this oracle (the reference's Rust
    path cannot run here); a drift of the oracle shows up as a diff against the committed vectors."""
    path = os.path.join(HERE, "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    gold = np.load(path)
    st, o = golden_scenes.run_oracle(name)
    assert o.substeps == int(gold["substeps"])
    assert np.array_equal(st.particles.flags, gold["flags"]) and np.array_equal(st.particles.collider_bits, gold["collider_bits"])
    for f in ("positions", "velocities", "position_gradients", "velocity_gradients"):
        a, b = getattr(st.particles, f), gold[f]
        assert np.allclose(a, b, rtol=1e-5, atol=1e-5 * max(1e-6, float(np.abs(b).max()))), f
