import sys, numpy as np
sys.path.insert(0, '.')
from squishy_volumes_b200 import scenes
from tests import parity
for name, sc in (("jelly", scenes.jelly_collision(side=12)), ("cube", scenes.elastic_cube(side=14, h=0.1))):
    (o, ro, eo), (g, rg, eg) = parity.run_both(sc, 1, store_grid=True)
    dv = np.abs(rg.particles.velocities - ro.particles.velocities).max(axis=1)
    bad = np.nonzero(dv > 1e-5)[0]
    print(name, "n", sc.n, "bad", len(bad), "max", dv.max())
    if len(bad):
        print(" bad idx", bad[:20], "v got", rg.particles.velocities[bad[:3]], "ref", ro.particles.velocities[bad[:3]])
        print(" bits", rg.particles.collider_bits[bad[:10]], "pos", ro.particles.positions[bad[:5]])
        print(" grid nodes got/ref", len(rg.grid_nodes.masses), len(ro.grid_nodes.masses))
