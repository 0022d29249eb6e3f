"""ncu target: a few substeps of the bench scene (run under ncu on the GPU box)."""
import sys
sys.path.insert(0, '.')
from squishy_volumes_b200 import scenes
from squishy_volumes_b200.state import B200State
from squishy_volumes_b200.types import RunParameters
name = sys.argv[1] if len(sys.argv) > 1 else "jelly_collision"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
import bench
scene = bench.make_scene(name, float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
scene.frame_input.consts.frames_per_second = 1
g = B200State.from_io_state(scene.io_state, scene.frame_input)
g.advance(None, scene.frame_input, RunParameters((steps - 0.5) * scene.time_step, scene.time_step))
print("substeps", g.substeps, "launches", g.kernel_launches, "ms", g.last_advance_ms)
