"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI
(squishy_volumes_b200.state.B200State -> lib/libsvb200.so), against the oracle on the same seeded
inputs and against the committed golden vectors.

Bars: cell keys, per-cell membership, active (block, layer) set, grid-node set,
flags and collider bits: bit-exact.  x, v, F, C, energy: tests/parity.py tolerances (stated there)."""
import os

import numpy as np
import pytest

from squishy_volumes_b200 import scenes
from squishy_volumes_b200.types import FatalError, Harness, ParticleFlags, RunParameters
from tests import golden_scenes, parity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def B200State():
    from squishy_volumes_b200.state import B200State as S
    return S


def h_of(scene):
    return scene.frame_input.consts.scaled_grid_node_size()


def test_library_loaded_and_device_listed():
    from squishy_volumes_b200 import abi, state
    abi.load()
    gpus = state.available_gpus()
    assert gpus and "B200" in gpus[0]


@pytest.mark.parametrize("maker,kw", [
    (scenes.elastic_cube, dict(side=14, h=0.1)),
    (scenes.jelly_collision, dict(side=12)),
    (scenes.sand_torus, dict(side=16)),
    (scenes.dam_break, dict(nx=16, ny=10, nz=8, viscous=True)),
    (scenes.mixed, dict(side=24, brick=4)),
])
def test_single_substep_parity(maker, kw):
    """One substep from identical state: every integer stage exact, floats at per-substep tolerance."""
    import oracle.oracle as orc
    scene = maker(**kw)
    (o, ro, eo), (g, rg, eg) = parity.run_both(scene, 1, store_grid=True)
    assert eo is None and eg is None and o.substeps == g.substeps == 1 and o.time == g.time
    parity.compare_states(rg, ro, rtol=parity.RTOL_STEP, h=h_of(scene))
    # binning: the device's cell of every particle == the oracle's formula on the device's own positions,
    # and == the oracle's cells (positions agree to the last bit after one substep from the same state here)
    cells_g = parity.cells_by_original(g)
    assert np.array_equal(cells_g, orc.shift_quadratic(rg.particles.positions, h_of(scene)))
    sm, cells = g.binning()
    assert np.array_equal(np.sort(sm), np.arange(scene.n, dtype=np.uint32))          # a permutation
    # block activation + grid node set
    ids, bits = g.active_blocks()
    got_blocks = set(map(tuple, np.concatenate([ids, bits[:, None].astype(np.int64)], axis=1).tolist()))
    assert got_blocks == parity.node_blocks(ro.grid_nodes)
    og, gg = ro.grid_nodes, rg.grid_nodes
    keep = og.contributor_counts > 0
    want_nodes = {tuple(r) for r in np.concatenate([og.node_ids[keep], og.collider_bits[keep][:, None].astype(np.int64)], axis=1).tolist()}
    got_nodes = [tuple(r) for r in np.concatenate([gg.node_ids, gg.collider_bits[:, None].astype(np.int64)], axis=1).tolist()]
    assert len(got_nodes) == len(set(got_nodes)) and set(got_nodes) == want_nodes
    # grid values (melded velocity, mass) on the common node set
    order_o = {k: i for i, k in enumerate(map(tuple, np.concatenate([og.node_ids, og.collider_bits[:, None].astype(np.int64)], axis=1).tolist()))}
    idx = np.array([order_o[k] for k in got_nodes])
    assert np.allclose(gg.masses, og.masses[idx], rtol=1e-5, atol=1e-6 * float(og.masses.max()))
    vmax = max(float(np.abs(og.velocities).max()), 1e-9)
    assert np.allclose(gg.velocities, og.velocities[idx], rtol=1e-4, atol=2e-5 * vmax)


def _runs(labels):
    """number of maximal runs of equal consecutive rows"""
    change = np.any(labels[1:] != labels[:-1], axis=1)
    return 1 + int(np.count_nonzero(change))


def test_binned_order_groups_tiles_and_cells():
    """After a substep the resident order is a counting sort on (tile, cell): every 4x4x4 block is one
    contiguous run of particles and inside it every cell is one contiguous run in ascending cell order —
    per-cell membership is exact (the reference's own intra-cell order is unspecified: unstable sort)."""
    import oracle.oracle as orc
    scene = scenes.jelly_collision(side=12)
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    sm, cells = g.binning()   # before any substep: identity order
    assert np.array_equal(sm, np.arange(scene.n))
    st = g.to_io_state()
    g.advance(None, scene.frame_input, RunParameters(1e-9, 1e-9))      # one (tiny) substep = one re-bin of `st`
    sm, _ = g.binning()
    assert np.array_equal(np.sort(sm), np.arange(scene.n, dtype=np.uint32))
    cells0 = orc.shift_quadratic(st.particles.positions, h_of(scene))[sm]      # cells at binning time, resident order
    blocks = cells0 >> 2
    assert _runs(blocks) == len(np.unique(blocks, axis=0))
    assert _runs(cells0) == len(np.unique(cells0, axis=0))
    local = ((cells0[:, 0] & 3) << 4) | ((cells0[:, 1] & 3) << 2) | (cells0[:, 2] & 3)
    same_block = ~np.any(blocks[1:] != blocks[:-1], axis=1)
    assert np.all(np.diff(local)[same_block] >= 0)


@pytest.mark.parametrize("name", sorted(golden_scenes.GOLDEN))
def test_golden_runs(name):
    """Tens of substeps (colliders, friction, moving meshes, goals, sand, viscous fluid, adaptive dt,
    several collider-bit layers per block) against the committed oracle vectors."""
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    scene, calls = golden_scenes.build(name)
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    st, err = golden_scenes.run_state(g, scene, calls)
    assert err is None
    assert g.substeps == int(gold["substeps"])
    assert g.time == pytest.approx(float(gold["time"]), rel=1e-6)
    p = st.particles
    live = (gold["flags"] & ParticleFlags.TOMBSTONED) == 0
    mism_flags = int(np.count_nonzero(p.flags != gold["flags"]))
    mism_bits = int(np.count_nonzero(p.collider_bits[live] != gold["collider_bits"][live]))
    # collider side decisions are thresholds on floats: after tens of substeps a particle sitting within
    # float noise of accept_distance / a face may legitimately flip one substep early or late
    assert mism_flags == 0
    assert mism_bits <= max(0, int(2e-3 * p.n)), mism_bits
    hh = h_of(scene)
    vmax = max(float(np.abs(gold["velocities"]).max()), 1e-9)
    ok = live & (p.collider_bits == gold["collider_bits"])
    for f, scale_floor in (("positions", 0.0), ("velocities", 0.0), ("position_gradients", 0.0), ("velocity_gradients", vmax / hh)):
        err_abs, scale = parity.field_error(getattr(p, f), gold[f], ok)
        scale = max(scale, scale_floor)
        assert err_abs <= parity.ATOL + parity.RTOL_RUN * scale, (f, err_abs, scale)
    # every live particle — the few with a flipped side bit included — under the per-particle percentile bounds
    rep = parity.assert_percentiles(p, gold, hh, parity.PCT_RUN, live, label=name)
    print(name, "bit mismatches", mism_bits, "percentiles", rep)
    ids, bits = g.active_blocks()
    got_blocks = np.unique(np.concatenate([ids, bits[:, None].astype(np.int32)], axis=1), axis=0)
    if mism_bits == 0:
        assert np.array_equal(got_blocks, gold["active_blocks"])


def test_energy_error_returned_with_valid_state():
    scene = scenes.jelly_collision(side=6)
    scene.io_state.particles.position_gradients[7] = -np.eye(3, dtype=np.float32)
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    st, err = g.produce_next_state(None, scene.frame_input, RunParameters(0.5e-3, 1e-3))
    assert err is not None and err.status & 8
    assert st.particles.flags[7] & ParticleFlags.FAILED
    # like the reference, the inverted particle's log(det F) poisons its neighbourhood with NaN; the state
    # is still returned (and stored by the caller, core/src/compute_thread.rs:165-169)
    import oracle.oracle as orc
    o = orc.OracleState.from_io_state(scene.io_state, scene.frame_input)
    _, eo = o.produce_next_state(None, scene.frame_input, RunParameters(0.5e-3, 1e-3))
    assert eo is not None and eo.status == 8


def test_fatal_errors():
    scene = scenes.jelly_collision(side=4)
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    hz = Harness()
    hz.cancel()
    with pytest.raises(FatalError) as e:
        g.produce_next_state(hz, scene.frame_input, RunParameters(1e-3, 1e-3))
    assert e.value.status == -1
    scene.io_state.time = 0.5            # frame 12 while frame 0 is loaded (xpu/src/frame_input.rs:266-278)
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    with pytest.raises(FatalError) as e:
        g.produce_next_state(None, scene.frame_input, RunParameters(0.6, 1e-3))
    assert e.value.status == -3
    scene.io_state.time = 0.0
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    with pytest.raises(FatalError) as e:
        g.produce_next_state(None, scene.frame_input, RunParameters(1e-3, 0.0))
    assert e.value.status == -2


def test_empty_and_single_particle_and_tombstoned():
    scene = scenes.jelly_collision(side=3)
    empty = scene.io_state.particles.select(np.zeros(0, dtype=np.int64))
    for k in scene.frame_input.keyframes:
        k.particle_flags = None
        k.particle_goal_positions = None
    from squishy_volumes_b200.types import IoState
    g = B200State().from_io_state(IoState(0.0, empty), scene.frame_input)
    st, err = g.produce_next_state(None, scene.frame_input, RunParameters(2.5e-3, 1e-3))
    assert err is None and st.particles.n == 0 and g.substeps == 3
    one = scene.io_state.particles.select(np.array([0]))
    (o, ro, _), (g, rg, _) = parity.run_both(type(scene)(scene.name, IoState(0.0, one), scene.frame_input, 1e-3, ""), 5)
    parity.compare_states(rg, ro, rtol=parity.RTOL_STEP, h=h_of(scene))
    # all-tombstoned input: nothing moves, nothing is binned
    dead = scene.io_state.particles.copy()
    dead.flags |= ParticleFlags.TOMBSTONED
    g = B200State().from_io_state(IoState(0.0, dead), scene.frame_input)
    st, err = g.produce_next_state(None, scene.frame_input, RunParameters(2.5e-3, 1e-3))
    assert err is None and np.array_equal(st.particles.positions, dead.positions)
    ids, _ = g.active_blocks()
    assert len(ids) == 0


def test_cull_matches_oracle_exactly():
    scene = scenes.elastic_cube(side=8, h=0.1)
    cut = float(np.median(scene.io_state.particles.positions[:, 2]))
    scene.frame_input.consts.domain_min = (-100.0, -100.0, cut)
    (o, ro, _), (g, rg, _) = parity.run_both(scene, 4)
    assert np.array_equal(rg.particles.flags, ro.particles.flags)
    assert 0 < np.count_nonzero(rg.particles.flags & ParticleFlags.TOMBSTONED) < scene.n
    parity.compare_states(rg, ro, rtol=parity.RTOL_RUN, h=h_of(scene))


def test_snapshot_restore_is_deterministic_in_integers():
    scene = scenes.jelly_collision(side=10)
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    g.snapshot()
    a, _ = g.produce_next_state(None, scene.frame_input, RunParameters(4.5e-3, 1e-3))
    g.restore()
    b, _ = g.produce_next_state(None, scene.frame_input, RunParameters(4.5e-3, 1e-3))
    assert np.array_equal(a.particles.flags, b.particles.flags)
    assert np.allclose(a.particles.positions, b.particles.positions, atol=1e-6)   # float atomics may reorder sums


def test_upload_restarts_a_session_handle():
    """svb_upload = from_io_state into an existing handle: a different state (other particle count, other buffer
    parity) goes in, the clock and the error words restart, and the run equals the one of a fresh handle / the oracle."""
    first = scenes.elastic_cube(side=10, h=0.1)
    scene = scenes.elastic_cube(side=14, h=0.1)       # more particles than the first state; same collider topology (ground plane)
    g = B200State().from_io_state(first.io_state, first.frame_input)
    g.produce_next_state(None, first.frame_input, RunParameters(2.5e-3, 1e-3))    # odd substep count: the live buffer is the second one
    g.upload(scene.io_state)
    assert g.substeps == 0 and g.time == scene.io_state.time
    params = RunParameters(4.5e-3, 1e-3)
    got, err = g.produce_next_state(None, scene.frame_input, params)
    assert err is None and g.substeps == 5
    fresh = B200State().from_io_state(scene.io_state, scene.frame_input)
    want, _ = fresh.produce_next_state(None, scene.frame_input, params)
    assert np.array_equal(got.particles.flags, want.particles.flags)
    assert np.array_equal(got.particles.collider_bits, want.particles.collider_bits)
    parity.compare_states(got, want, rtol=parity.RTOL_STEP, h=h_of(scene))
    import oracle.oracle as orc
    ro, _ = orc.OracleState.from_io_state(scene.io_state, scene.frame_input).produce_next_state(None, scene.frame_input, params)
    parity.compare_states(got, ro, rtol=parity.RTOL_RUN, h=h_of(scene))
    # a failed run does not poison the next upload
    bad = scenes.elastic_cube(side=8, h=0.1)
    bad.io_state.particles.position_gradients[0] = -np.eye(3, dtype=np.float32)
    g.upload(bad.io_state)
    _, err = g.produce_next_state(None, bad.frame_input, RunParameters(1.5e-3, 1e-3))
    assert err is not None and err.status & 8
    g.upload(scene.io_state)
    again, err2 = g.produce_next_state(None, scene.frame_input, params)
    assert err2 is None
    parity.compare_states(again, want, rtol=parity.RTOL_STEP, h=h_of(scene))


def test_million_particle_properties():
    """BASELINE configs[1] at full size (1.02 M particles): size-independent properties instead of the oracle —
    sort_map is a permutation, the resident order is sorted by bin key, mass is conserved on the grid,
    momentum is conserved by P2G->G2P without gravity, every live particle lands in its cell."""
    scene = scenes.jelly_collision(side=80)
    p0 = scene.io_state.particles
    g = B200State().from_io_state(scene.io_state, scene.frame_input)
    st, err = g.produce_next_state(None, scene.frame_input, RunParameters(2.5e-3, 1e-3, store_grid=True))
    assert err is None and g.substeps == 3
    sm, cells = g.binning()
    assert np.array_equal(np.sort(sm), np.arange(scene.n, dtype=np.uint32))
    mom0 = (p0.velocities.astype(np.float64) * p0.mass[:, None]).sum(axis=0)
    mom1 = (st.particles.velocities.astype(np.float64) * st.particles.mass[:, None]).sum(axis=0)
    assert np.allclose(mom0, mom1, atol=1e-4 * float(np.abs(p0.mass).sum()))
    grid = st.grid_nodes
    assert grid.masses.sum(dtype=np.float64) == pytest.approx(float(p0.mass.sum(dtype=np.float64)), rel=1e-5)
    assert np.isfinite(st.particles.position_gradients).all()
