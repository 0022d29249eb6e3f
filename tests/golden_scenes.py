"""Golden vectors: small seeded scenes covering every material / collider / dt mode of the path.
`python -m tests.golden_scenes` regenerates tests/golden/*.npz from the ORACLE (the reference is Rust
and cannot be executed in this environment; see DESIGN.md §5 — parity is unpinned end-to-end, these
fixtures pin the oracle against drift and let the GPU tests run against committed numbers)."""
from __future__ import annotations

import os

import numpy as np

from squishy_volumes_b200 import scenes
from squishy_volumes_b200.types import ParticleFlags, RunParameters

HERE = os.path.dirname(os.path.abspath(__file__))


def _cube_drop():
    sc = scenes.elastic_cube(side=10, h=0.1)
    sc.io_state.particles.velocities[:, 2] = -3.0      # reaches the ground plane within the run
    return sc, 40, False


def _cube_adaptive():
    sc = scenes.elastic_cube(side=10, h=0.1)
    sc.io_state.particles.velocities[:, 2] = -3.0
    sc.time_step = 2e-3
    return sc, 12, True


def _jelly():
    sc = scenes.jelly_collision(side=10)
    sc.io_state.particles.velocities[:, 0] *= 4.0      # the blocks touch after ~10 substeps
    return sc, 30, False


def _sand():
    sc = scenes.sand_torus(side=14)
    sc.io_state.particles.velocities[:, 2] = -8.0      # hits the torus within the run
    sc.time_step = 2e-4
    return sc, 40, False


def _dam():
    sc = scenes.dam_break(nx=16, ny=8, nz=8, viscous=True)
    return sc, 30, False


def _mixed_adaptive():
    sc = scenes.mixed(side=16, brick=4)
    sc.io_state.particles.velocities[:, 2] = -4.0
    sc.time_step = 2e-4
    return sc, 15, True


def _goals_moving_collider():
    sc = scenes.elastic_cube(side=8, h=0.1, n_keyframes=3)
    n = sc.n
    top = sc.io_state.particles.positions[:, 2] > np.percentile(sc.io_state.particles.positions[:, 2], 85)
    for f, k in enumerate(sc.frame_input.keyframes):
        k.particle_flags = np.where(top, ParticleFlags.HAS_GOAL, 0).astype(np.uint32)
        k.particle_goal_positions = (sc.io_state.particles.positions + np.array([0.02 * f, 0, 0.01 * f], np.float32)).astype(np.float32)
        k.vertex_positions = k.vertex_positions + np.array([0, 0, 0.05 * f], np.float32)   # plane rises
        k.gravity = (0.0, 0.0, -9.8 + f)
    sc.frame_input.consts.frames_per_second = 50
    return sc, 30, False


def _split_layers():
    """A plane and a torus cutting THROUGH a jelly block: particles on both sides of each collider, so
    single grid blocks carry several collider-bit layers and meld_grid has work to do."""
    sc = scenes.jelly_collision(side=12)
    p = sc.io_state.particles
    p.velocities[:, 0] *= 2.0
    p.velocities[:, 2] = np.where(p.positions[:, 2] > 0.013, -1.5, 1.5).astype(np.float32)
    plane = scenes.plane_mesh(0.013, 2.0)
    torus = scenes.torus_mesh(0.16, 0.05, 24, 12, center=(0.0, 0.0, 0.05))
    sc.frame_input = scenes._frame_input(sc.frame_input.consts, [plane, torus], p.n, (0, 0, -9.8), [0.5, 0.2], [0.3, 0.0], 3, moving=(0.0, 0.0, 0.004))
    return sc, 25, False


GOLDEN = {
    "split_layers": _split_layers,
    "cube_drop": _cube_drop,
    "cube_adaptive": _cube_adaptive,
    "jelly": _jelly,
    "sand_torus": _sand,
    "dam_viscous": _dam,
    "mixed_adaptive": _mixed_adaptive,
    "goals_moving_collider": _goals_moving_collider,
}


def build(name):
    """-> (scene, [RunParameters per produce_next_state call]).  Like core/src/compute_thread.rs:126-192 the
    run is cut at output-frame boundaries: frame f is loaded, then the state advances to (f+1)/fps."""
    scene, n, adaptive = GOLDEN[name]()
    dt = scene.time_step
    t0 = scene.io_state.time
    target = t0 + (n * dt if adaptive else (n - 0.5) * dt)
    fps = scene.frame_input.consts.frames_per_second
    calls = []
    frame = int(np.floor(t0 * fps))
    while True:
        end = (frame + 1) / fps
        calls.append((frame, RunParameters(target_time=min(target, end), max_time_step=dt, adaptive_time_steps=adaptive, store_grid=True)))
        if end >= target:
            break
        frame += 1
    return scene, calls


def run_state(state, scene, calls):
    """Drive any back end (oracle or B200) through the calls; returns (IoState, error)."""
    st, err = None, None
    for frame, params in calls:
        scene.frame_input.load(frame)
        st, err = state.produce_next_state(None, scene.frame_input, params)
        if err is not None:
            break
    scene.frame_input.load(calls[0][0])
    return st, err


def run_oracle(name):
    import oracle.oracle as orc
    scene, calls = build(name)
    o = orc.OracleState.from_io_state(scene.io_state, scene.frame_input)
    st, err = run_state(o, scene, calls)
    assert err is None, err
    return st, o


def main():
    os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
    for name in GOLDEN:
        st, o = run_oracle(name)
        p = st.particles
        g = st.grid_nodes
        keep = g.contributor_counts > 0
        blocks = np.unique(np.concatenate([g.node_ids[keep] >> 2, g.collider_bits[keep][:, None].astype(np.int32)], axis=1), axis=0)
        np.savez_compressed(os.path.join(HERE, "golden", name + ".npz"), substeps=o.substeps, time=o.time, flags=p.flags, collider_bits=p.collider_bits,
                            positions=p.positions, velocities=p.velocities, position_gradients=p.position_gradients, velocity_gradients=p.velocity_gradients,
                            elastic_energies=p.elastic_energies, active_blocks=blocks.astype(np.int32))
        print(name, "substeps", o.substeps, "n", p.n, "nonzero bits", int(np.count_nonzero(p.collider_bits)), "tomb", int(np.count_nonzero(p.flags & ParticleFlags.TOMBSTONED)),
              "distinct bits", len(np.unique(p.collider_bits)), "max|v|", float(np.abs(p.velocities).max()))


if __name__ == "__main__":
    main()
