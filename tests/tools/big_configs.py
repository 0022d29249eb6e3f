"""BASELINE.json configs 3-5 at full size on ONE B200 (run on the GPU box): size-independent properties instead of the oracle,
plus device-timed substep rates.  Prints one JSON line per config; the summary is committed under profiles/.

    python tests/tools/big_configs.py sand_torus 1 dam_break 1 mixed 1          # name, scale pairs
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from squishy_volumes_b200.state import B200State  # noqa: E402
from squishy_volumes_b200.types import ParticleFlags, RunParameters  # noqa: E402


def run(name, scale, steps=10, warm=3):
    t0 = time.time()
    sc = bench.make_scene(name, scale)
    sc.frame_input.consts.frames_per_second = 1
    gen_s = time.time() - t0
    p0 = sc.io_state.particles
    h = sc.frame_input.consts.scaled_grid_node_size()
    dt = sc.time_step
    t0 = time.time()
    g = B200State.from_io_state(sc.io_state, sc.frame_input)
    up_s = time.time() - t0
    err = g.advance(None, sc.frame_input, RunParameters((warm - 0.5) * dt, dt))
    assert err is None, err
    err = g.advance(None, sc.frame_input, RunParameters((warm + steps - 0.5) * dt, dt))
    assert err is None, err
    ms = g.last_advance_ms / steps
    err = g.advance(None, sc.frame_input, RunParameters((warm + steps + 0.5) * dt, dt, store_grid=True))   # one more substep with exact node masks for the grid checks
    assert err is None, err
    sm, cells = g.binning()
    st = g.to_io_state(store_grid=True)
    p1 = st.particles
    out = {"config": name, "particles": int(sc.n), "description": sc.description, "substeps": int(g.substeps), "ms_per_substep": ms,
           "particle_substeps_per_s": sc.n / (ms * 1e-3), "generate_s": round(gen_s, 1), "create_upload_s": round(up_s, 2)}
    # the binned order is a permutation and sorted by (tile, cell): contiguous runs
    assert np.array_equal(np.sort(sm), np.arange(sc.n, dtype=np.uint32))
    live = (p1.flags & ParticleFlags.TOMBSTONED) == 0
    out["tombstoned"] = int((~live).sum())
    ids, bits = g.active_blocks()
    assert np.isfinite(p1.positions).all() and np.isfinite(p1.position_gradients).all() and np.isfinite(p1.velocities).all()
    # grid mass = live particle mass (every particle's weights sum to one; tombstoned particles take no part)
    gm = float(st.grid_nodes.masses.sum(dtype=np.float64))
    pm = float(p0.mass[live].sum(dtype=np.float64))
    out["grid_mass_rel_err"] = abs(gm - pm) / pm
    # without collider layers the melded grid holds every particle's mass exactly once; with layers meld_grid.rs:41-59 adds the mass of
    # every compatible sibling to each of them, so the sum over (node, bits) entries counts shared mass once per sibling: >= particle mass
    n_layers = int(np.unique(st.grid_nodes.collider_bits).shape[0])
    assert (out["grid_mass_rel_err"] < 1e-4) if n_layers == 1 else (pm * (1 - 1e-4) <= gm <= pm * n_layers), out
    # materials are carried, never mixed up by the re-bin
    for f in ("mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha"):
        assert np.array_equal(getattr(p1, f), getattr(p0, f)), f
    assert np.array_equal(p1.flags & 0x1f, p0.flags & 0x1f)
    det = np.linalg.det(p1.position_gradients[live][:: max(1, int(live.sum()) // 200000)].astype(np.float64))
    out["min_det_F_sampled"] = float(det.min())
    assert det.min() > 0
    out["active_blocks"] = int(ids.shape[0])
    out["collider_layers"] = int(np.unique(bits).shape[0])
    # the scene starts in contact (scenes.py contact=True): colliders are near from the first substep, material deforms / yields
    out["particles_near_a_collider"] = int(np.count_nonzero(p1.collider_bits))
    dF = p1.position_gradients.reshape(-1, 9) - np.eye(3, dtype=np.float32).reshape(9)
    out["particles_deformed"] = int(np.count_nonzero(np.abs(dF).max(axis=1) > 1e-5))   # (a weakly compressible fluid keeps F = J^(1/3) I: small numbers)
    if name != "jelly_collision":
        assert out["collider_layers"] > 1 and out["particles_near_a_collider"] > 0 and out["particles_deformed"] > 0, out
    out["max_speed"] = float(np.linalg.norm(p1.velocities[live], axis=1).max())
    g.close()
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    for k in range(0, len(args), 2):
        print(json.dumps(run(args[k], float(args[k + 1]))), flush=True)
