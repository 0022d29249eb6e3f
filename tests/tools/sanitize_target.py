"""compute-sanitizer target (GPU box): a few substeps of every code path on small scenes — collider + layers (mesh path: collide,
k_bin, meld), meshless binning-ahead path, adaptive steps with the device clock, sand / fluid return mapping, cull, grid download.
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tests/tools/sanitize_target.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
from squishy_volumes_b200 import scenes  # noqa: E402
from squishy_volumes_b200.state import B200State  # noqa: E402
from squishy_volumes_b200.types import RunParameters  # noqa: E402
from tests import golden_scenes  # noqa: E402


def run(name, sc, n, adaptive=False, store_grid=False):
    sc.frame_input.consts.frames_per_second = 1
    dt = sc.time_step
    g = B200State.from_io_state(sc.io_state, sc.frame_input)
    target = n * dt if adaptive else (n - 0.5) * dt
    st, err = g.produce_next_state(None, sc.frame_input, RunParameters(target, dt * (4 if adaptive else 1), adaptive_time_steps=adaptive, store_grid=store_grid))
    assert err is None and np.isfinite(st.particles.positions).all()
    print(name, sc.n, "particles", g.substeps, "substeps", g.kernel_launches, "launches", flush=True)
    g.close()


run("cube on plane (mesh path)", scenes.elastic_cube(side=16, h=0.1), 5, store_grid=True)
run("jelly (binning ahead)", scenes.jelly_collision(side=10), 6)
run("jelly adaptive (device clock)", scenes.jelly_collision(side=10), 6, adaptive=True)
run("split layers (meld)", golden_scenes.GOLDEN["split_layers"]()[0], 4, store_grid=True)
run("sand on torus (return mapping, long triangle runs)", scenes.sand_torus(side=14, contact=True), 4)
run("mixed adaptive", scenes.mixed(side=16, brick=4, contact=True), 4, adaptive=True)
sc = scenes.elastic_cube(side=8, h=0.1)
sc.frame_input.consts.domain_min = (-100.0, -100.0, float(np.median(sc.io_state.particles.positions[:, 2])))
run("cull (tombstoned rows)", sc, 4)
