"""Development driver (GPU box): oracle vs CUDA on small scenes, prints errors."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, '.')
from squishy_volumes_b200 import scenes
from squishy_volumes_b200.types import ParticleFlags
from tests import parity

def report(name, scene, n_sub, adaptive=False):
    print(f"== {name}: n={scene.n} substeps={n_sub} adaptive={adaptive}", flush=True)
    try:
        (o, ro, eo), (g, rg, eg) = parity.run_both(scene, n_sub, adaptive=adaptive, store_grid=True)
    except Exception as e:
        traceback.print_exc()
        return
    print("  oracle substeps", o.substeps, "time", o.time, "| gpu substeps", g.substeps, "time", g.time, "errs", eo, eg)
    gp, rp = rg.particles, ro.particles
    print("  flags equal:", np.array_equal(gp.flags, rp.flags), " bits equal:", np.array_equal(gp.collider_bits, rp.collider_bits),
          " nonzero bits:", int(np.count_nonzero(rp.collider_bits)), " tomb:", int(np.count_nonzero(rp.flags & ParticleFlags.TOMBSTONED)))
    for f in parity.FIELDS + ("elastic_energies",):
        err, scale = parity.field_error(getattr(gp, f), getattr(rp, f))
        print(f"  {f:20s} err {err:.3e} scale {scale:.3e} rel {err/scale:.3e}")
    co = np.zeros((scene.n, 3), np.int32); 
    sm = o.sort_map()
    import oracle.oracle as orc
    cells_o = orc.shift_quadratic(rp.positions, scene.frame_input.consts.scaled_grid_node_size())
    cells_g = parity.cells_by_original(g)
    cg = orc.shift_quadratic(gp.positions, scene.frame_input.consts.scaled_grid_node_size())
    print("  gpu cells == oracle formula on gpu positions:", np.array_equal(cells_g, cg), " cells equal across impls:", np.array_equal(cells_g, cells_o))
    gb = parity.node_blocks(ro.grid_nodes)
    ids, bits = g.active_blocks()
    ab = set(map(tuple, np.concatenate([ids, bits[:, None].astype(np.int64)], axis=1).tolist()))
    print("  active blocks: oracle", len(gb), "gpu", len(ab), "equal", gb == ab)
    gg = rg.grid_nodes
    og = ro.grid_nodes
    keep = og.contributor_counts > 0
    so = set(map(tuple, np.concatenate([og.node_ids[keep], og.collider_bits[keep][:, None].astype(np.int64)], axis=1).tolist()))
    sg = set(map(tuple, np.concatenate([gg.node_ids, gg.collider_bits[:, None].astype(np.int64)], axis=1).tolist()))
    print("  grid nodes: oracle", len(so), "gpu", len(sg), "equal", so == sg)
    print("  stage launches", g.kernel_launches)

report("cube", scenes.elastic_cube(side=12, h=0.1), 1)
report("cube", scenes.elastic_cube(side=12, h=0.1), 20)
report("jelly", scenes.jelly_collision(side=10), 10)
report("sand", scenes.sand_torus(side=16), 10)
report("dam", scenes.dam_break(nx=16, ny=8, nz=8), 10)
report("dam-visc", scenes.dam_break(nx=16, ny=8, nz=8, viscous=True), 10)
report("mixed", scenes.mixed(side=24, brick=4), 10)
report("cube-adaptive", scenes.elastic_cube(side=12, h=0.1), 10, adaptive=True)
report("sand-adaptive", scenes.sand_torus(side=16), 10, adaptive=True)
