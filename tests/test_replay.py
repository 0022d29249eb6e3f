"""svb_replay (csrc/svb_replay.cpp): the compute thread's frame loop (core/src/compute_thread.rs:60-190) over the two C ABIs —
recorded input file in, frame files out."""
import os
import subprocess

import numpy as np
import pytest

from squishy_volumes_b200 import abi, files, scenes
from squishy_volumes_b200.types import RunParameters
from tests import parity


def record(scene, cache_dir, E=1e4, nu=0.3):
    """Record a synthetic scene (unit transforms: F = I) the way the Blender add-on records its input."""
    p = scene.io_state.particles
    n = p.n
    t = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    t[:, 3, :3] = p.positions
    size = np.cbrt(p.initial_volume.astype(np.float64)).astype(np.float32)
    topo = scene.frame_input.colliders
    objects = {"cube": ("particles", n)}
    for c, tp in enumerate(topo):
        objects[f"mesh{c}"] = ("collider", tp.num_vertices, tp.triangles.shape[0])
    w = files.InputWriter(os.path.join(cache_dir, "simulation_input.bin"), scene.frame_input.consts, objects)
    for k in scene.frame_input.keyframes:
        colliders, v0, t0 = {}, 0, 0
        for c, tp in enumerate(topo):
            nt = tp.triangles.shape[0]
            colliders[f"mesh{c}"] = {"vertex_positions": k.vertex_positions[v0:v0 + tp.num_vertices], "triangle_indices": tp.triangles,
                                     "triangle_frictions": k.triangle_frictions[t0:t0 + nt], "triangle_dampings": k.triangle_dampings[t0:t0 + nt]}
            v0 += tp.num_vertices
            t0 += nt
        w.record_frame(k.gravity, {"cube": {"flags": p.flags, "transforms": t, "sizes": size, "densities": p.mass / (size * size * size),
                                            "youngs_moduluses": np.full(n, E, np.float32), "poissons_ratios": np.full(n, nu, np.float32),
                                            "initial_velocities": p.velocities}}, colliders)
    w.finish()


def test_replay_fails_loudly_without_a_device(tmp_path):
    """No CPU fallback: without a CUDA device the frame loop stops at from_io_state — after frame 0 (the initial state, which
    needs no device) has been stored like compute_thread.rs:81-89 does."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    abi.build()
    sc = scenes.elastic_cube(side=5, h=0.1, n_keyframes=2)
    record(sc, str(tmp_path))
    r = subprocess.run([abi.REPLAY_PATH, str(tmp_path), "3", "0.002"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "svb_create failed" in r.stderr
    st = files.read_frame(files.frame_path(str(tmp_path), 0))
    assert st.time == 0.0 and np.array_equal(st.particles.positions, sc.io_state.particles.positions)
    assert not os.path.exists(files.frame_path(str(tmp_path), 1))
    r = subprocess.run([abi.REPLAY_PATH, str(tmp_path / "nowhere"), "3", "0.002"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "failed to open" in r.stderr


@pytest.mark.gpu
def test_replay_writes_the_frames_of_the_public_api(tmp_path):
    """Three output frames of a cube dropping on a ground-plane collider: every stored frame equals what the host mirror
    (B200State.produce_next_state per frame) returns and, at the end, the oracle; then a resume from the last checkpoint."""
    from squishy_volumes_b200.state import B200State
    sc = scenes.elastic_cube(side=8, h=0.1, n_keyframes=3)
    record(sc, str(tmp_path))
    dt = 2e-3
    r = subprocess.run([abi.REPLAY_PATH, str(tmp_path), "4", str(dt)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("substeps") == 3
    inp = files.InputFile(os.path.join(str(tmp_path), "simulation_input.bin"))
    fi = inp.frame_input()
    init = inp.initialize_io_state()
    h = fi.consts.scaled_grid_node_size()
    g = B200State.from_io_state(init, fi)
    import oracle.oracle as orc
    o = orc.OracleState.from_io_state(init, fi)
    for frame in (1, 2, 3):
        fi.load(frame - 1)
        params = RunParameters(target_time=frame / fi.consts.frames_per_second, max_time_step=dt, store_grid=True)
        want, err = g.produce_next_state(None, fi, params)
        assert err is None
        got = files.read_frame(files.frame_path(str(tmp_path), frame))
        assert got.time == pytest.approx(want.time, abs=1e-12)
        assert np.array_equal(got.particles.flags, want.particles.flags) and np.array_equal(got.particles.collider_bits, want.particles.collider_bits)
        parity.compare_states(got, want, rtol=parity.RTOL_STEP, h=h)
        assert got.grid_nodes is not None and got.grid_nodes.masses.shape[0] == want.grid_nodes.masses.shape[0] > 0
        ro, eo = o.produce_next_state(None, fi, params)
        assert eo is None
    parity.compare_states(got, ro, rtol=parity.RTOL_RUN, h=h)
    # resume: frames 4 from the checkpoint of frame 3 (compute_thread.rs:90-96)
    r = subprocess.run([abi.REPLAY_PATH, str(tmp_path), "5", str(dt), "--next-frame", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    fi.load(3)
    want, _ = g.produce_next_state(None, fi, RunParameters(target_time=4 / fi.consts.frames_per_second, max_time_step=dt, store_grid=True))
    got = files.read_frame(files.frame_path(str(tmp_path), 4))
    parity.compare_states(got, want, rtol=parity.RTOL_RUN, h=h)
