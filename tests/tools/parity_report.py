"""Per-particle relative error percentiles of the CUDA path against the oracle (GPU box): one line per scene.
    python tests/tools/parity_report.py > gpurun_out/parity_percentiles.txt        (committed under profiles/)"""
import sys

import numpy as np

sys.path.insert(0, ".")
from squishy_volumes_b200 import scenes  # noqa: E402
from squishy_volumes_b200.types import ParticleFlags  # noqa: E402
from tests import golden_scenes, parity  # noqa: E402


def line(label, got, ref, h):
    live = (ref.flags & ParticleFlags.TOMBSTONED) == 0
    rep = parity.error_percentiles(got, ref, h, live)
    mism = int(np.count_nonzero(got.collider_bits[live] != ref.collider_bits[live]))
    print(f"{label:34s} n={got.n:8d} bit-mismatch={mism:4d} " + "  ".join(f"{k[:9]}: " + "/".join(f"{v:.1e}" for v in vals) for k, vals in rep.items()), flush=True)


print("# per-particle relative error e_i = |got_i - ref_i| / max(|ref_i|, floor) (tests/parity.py), P50/P99/P99.9/max over live particles")
for name, mk, n in (("cube 1 substep", lambda: scenes.elastic_cube(side=14, h=0.1), 1), ("jelly 1 substep", lambda: scenes.jelly_collision(side=12), 1),
                    ("sand 1 substep", lambda: scenes.sand_torus(side=16), 1), ("dam 1 substep", lambda: scenes.dam_break(nx=16, ny=10, nz=8, viscous=True), 1),
                    ("mixed 1 substep", lambda: scenes.mixed(side=24, brick=4), 1), ("sand contact 14 substeps", lambda: scenes.sand_torus(side=40, contact=True), 14),
                    ("dam contact 14 substeps", lambda: scenes.dam_break(nx=48, ny=24, nz=24, contact=True), 14), ("mixed contact 14 substeps", lambda: scenes.mixed(side=48, brick=8, contact=True), 14)):
    sc = mk()
    (o, ro, eo), (g, rg, eg) = parity.run_both(sc, n)
    line(name, rg.particles, ro.particles, sc.frame_input.consts.scaled_grid_node_size())
for name in sorted(golden_scenes.GOLDEN):
    import oracle.oracle as orc
    from squishy_volumes_b200.state import B200State
    sc, calls = golden_scenes.build(name)
    g = B200State.from_io_state(sc.io_state, sc.frame_input)
    st, err = golden_scenes.run_state(g, sc, calls)
    o = orc.OracleState.from_io_state(sc.io_state, sc.frame_input)
    so, _ = golden_scenes.run_state(o, sc, calls)
    line(f"golden {name} ({g.substeps} substeps)", st.particles, so.particles, sc.frame_input.consts.scaled_grid_node_size())
