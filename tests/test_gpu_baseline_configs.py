"""BASELINE.json configs at the sizes it names, against the oracle:

* config 2 — 1 024 000-particle jelly collision, 100 substeps: EVERY substep the device's binning keys equal
  floor(x/h - 1/2) of the positions it binned, the resident order groups every block and every cell into one run
  (per-cell membership), and cell keys / membership equal the oracle's wherever the two position sets agree bit for bit
  (cpu/src/phase/sort.rs:29-33) — plus per-particle error percentiles after 100 substeps of contact;
* config 1 — 46^3 = 97 336-particle elastic cube on the ground-plane collider, 240 substeps with fixed dt and one run with
  adaptive time steps (the reference's default mode, python/src/squishy_volumes_extension/panels/panel_simulate.py:61).

The oracle does 1 M particles in ~0.25 s per substep on 16 cores, so these finish in a few minutes on the GPU box."""
import numpy as np
import pytest

from squishy_volumes_b200 import scenes
from squishy_volumes_b200.types import ParticleFlags, RunParameters
from tests import parity

pytestmark = pytest.mark.gpu


def _states(scene):
    import oracle.oracle as orc
    from squishy_volumes_b200.state import B200State
    return orc.OracleState.from_io_state(scene.io_state, scene.frame_input), B200State.from_io_state(scene.io_state, scene.frame_input)


def _key(cells):
    c = cells.astype(np.int64) + (1 << 20)
    return (c[:, 0] << 42) | (c[:, 1] << 21) | c[:, 2]


def _runs(keys):
    return 1 + int(np.count_nonzero(keys[1:] != keys[:-1]))


def test_config2_million_jelly_binning_every_substep():
    import oracle.oracle as orc
    scene = scenes.jelly_collision(side=80)
    assert scene.n == 1_024_000
    scene.io_state.particles.velocities[:, 0] *= 8.0   # the 2h gap closes after ~5 substeps: 95 of the 100 substeps are in contact
    scene.frame_input.consts.frames_per_second = 1     # one long frame: no keyframe reload inside the loop
    h = scene.frame_input.consts.scaled_grid_node_size()
    dt = scene.time_step
    o, g = _states(scene)
    x_g = scene.io_state.particles.positions.copy()
    x_o = x_g.copy()
    n_sub = 100
    bit_equal_min, moved_cells, cell_mismatch_where_unequal = 1.0, 0, 0
    for k in range(1, n_sub + 1):
        params = RunParameters(target_time=(k - 0.5) * dt, max_time_step=dt)
        cells_before = orc.shift_quadratic(x_g, h)        # the reference's key of the positions this substep bins (kernels.rs:46-49)
        assert g.advance(None, scene.frame_input, params) is None
        ro, eo = o.produce_next_state(None, scene.frame_input, params)
        assert eo is None and g.substeps == o.substeps == k
        sm, cells_dev = g.binning()                        # resident order after the substep; keys of the ADVANCED positions
        st = g.to_io_state()
        x_new = st.particles.positions
        # (1) the resident order is a permutation that groups the binned positions by block and by cell: per-cell membership
        assert np.array_equal(np.sort(sm), np.arange(scene.n, dtype=np.uint32))
        cb = cells_before[sm]
        kc, kb = _key(cb), _key(cb >> 2)
        assert _runs(kb) == np.unique(kb).size, f"substep {k}: a block is split over several runs"
        assert _runs(kc) == np.unique(kc).size, f"substep {k}: a cell is split over several runs"
        # (2) the device's keys of the advanced positions == the reference formula on those positions, bit for bit
        cells_formula = orc.shift_quadratic(x_new, h)
        dev_by_orig = np.empty_like(cells_dev)
        dev_by_orig[sm] = cells_dev
        assert np.array_equal(dev_by_orig, cells_formula), f"substep {k}: device cell keys differ from floor(x/h - 1/2)"
        # (3) against the oracle: wherever the two position sets agree bit for bit the keys (hence the membership) are equal;
        #     elsewhere (float atomics reorder sums) they may differ only for a particle within rounding of a cell face
        x_o = ro.particles.positions
        same = np.all(x_new.view(np.uint32) == x_o.view(np.uint32), axis=1)
        cells_o = orc.shift_quadratic(x_o, h)
        assert np.array_equal(cells_formula[same], cells_o[same])
        bit_equal_min = min(bit_equal_min, float(same.mean()))
        cell_mismatch_where_unequal += int(np.count_nonzero(np.any(cells_formula[~same] != cells_o[~same], axis=1)))
        moved_cells += int(np.count_nonzero(np.any(cells_formula != cells_before, axis=1)))
        x_g = x_new
    assert moved_cells > 100_000, "the run must actually move particles across cells"
    # positions that differ (by ~1e-6 h: float atomics reorder the grid sums) may fall on different sides of a cell face: expected
    # 3 axes x 2e-6 per particle-substep, measured 894 over the 1.02e8 particle-substeps of this run
    assert cell_mismatch_where_unequal <= 3e-5 * scene.n * n_sub, cell_mismatch_where_unequal
    rep = parity.compare_states(st, ro, rtol=parity.RTOL_RUN, h=h)
    print("config 2, 100 substeps: fraction of bit-equal positions (min over substeps)", bit_equal_min, "cell changes", moved_cells,
          "key mismatches among unequal positions", cell_mismatch_where_unequal, "percentiles", rep["percentiles"])


@pytest.mark.parametrize("adaptive", [False, True])
def test_config1_elastic_cube_full(adaptive):
    scene = scenes.elastic_cube(side=46)
    assert scene.n == 97_336
    h = scene.frame_input.consts.scaled_grid_node_size()
    # fixed: 240 substeps of 1e-3.  adaptive: the same 0.24 s with max_time_step = 2e-2, so that the four limits of
    # limit_time_step.rs:25-223 (sound speed / isolated particle first, the velocity limit after the impact) decide every step
    dt = 2e-2 if adaptive else scene.time_step
    o, g = _states(scene)
    n_sub = 240
    fps = scene.frame_input.consts.frames_per_second
    target = n_sub * scene.time_step if adaptive else (n_sub - 0.5) * dt
    ro = rg = None
    frame = 0
    while True:   # like the compute thread: one produce_next_state per output frame (core/src/compute_thread.rs:126-192)
        end = min(target, (frame + 1) / fps)
        scene.frame_input.load(frame)
        params = RunParameters(target_time=end, max_time_step=dt, adaptive_time_steps=adaptive)
        rg, eg = g.produce_next_state(None, scene.frame_input, params)
        ro, eo = o.produce_next_state(None, scene.frame_input, params)
        assert eg is None and eo is None
        if end >= target:
            break
        frame += 1
    scene.frame_input.load(0)
    if adaptive:
        assert g.substeps == o.substeps and 20 <= g.substeps < 100 and g.time == pytest.approx(o.time, rel=1e-5)
        assert g.allowed_time_step == pytest.approx(o.allowed_time_step, rel=1e-4) and g.allowed_time_step < 0.5 * dt
    else:
        assert g.substeps == o.substeps == n_sub and g.time == o.time
    # the cube has landed on the plane: collider bits set, contact forces at work
    live = (ro.particles.flags & ParticleFlags.TOMBSTONED) == 0
    assert np.count_nonzero(ro.particles.collider_bits) > 1000
    mism = int(np.count_nonzero(rg.particles.collider_bits[live] != ro.particles.collider_bits[live]))
    assert mism <= 2e-3 * scene.n, mism   # side decisions are float thresholds: a particle within rounding of the accept distance may flip a substep apart
    assert np.array_equal(rg.particles.flags, ro.particles.flags)
    rep = parity.assert_percentiles(rg.particles, ro.particles, h, parity.PCT_RUN, live, label=f"config 1 adaptive={adaptive}")
    for name in parity.FIELDS:
        err, scale = parity.field_error(getattr(rg.particles, name), getattr(ro.particles, name), live)
        assert err <= parity.ATOL + 5 * parity.RTOL_RUN * max(scale, 1e-3), (name, err, scale)
    print("config 1", "adaptive" if adaptive else "fixed", "substeps", g.substeps, o.substeps, "bit mismatches", mism, "percentiles", rep)
