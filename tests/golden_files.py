"""Golden fixtures of the file layer: `python -m tests.golden_files` writes tests/golden/frame_small.bin and
tests/golden/input_small.bin with the format ORACLE (oracle/bincode_ref.py) from seeded inputs.  They guard both the
oracle and csrc/svb_files.cpp against drift; they are NOT bytes produced by the reference (no Rust toolchain here)."""
import os

import numpy as np

from oracle import bincode_ref as ref
from tests import test_files as tf

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONSTS = dict(tf.CONSTS, simulation_scale=2.0, grid_node_size=0.25)
OBJECTS = {"jelly": ("particles", 12), "water": ("particles", 5), "floor": ("collider", 4, 2), "ball": ("collider", 6, 8)}


def frame_inputs():
    return 1.0 / 24.0, tf.random_particles(37, seed=21), tf.random_grid(11, seed=22)


def input_frames():
    rng = np.random.default_rng(23)
    frames = []
    for k in range(2):
        frames.append({"gravity": (0.0, 0.0, -9.8 + k), "particles": {"jelly": tf.particles_input(12, rng), "water": tf.particles_input(5, rng, with_everything=(k == 0))},
                       "colliders": {"floor": tf.collider_input(4, 2, rng), "ball": tf.collider_input(6, 8, rng)}})
    return frames


def main():
    os.makedirs(HERE, exist_ok=True)
    t, p, g = frame_inputs()
    with open(os.path.join(HERE, "frame_small.bin"), "wb") as f:
        f.write(ref.encode_io_state(t, p, g))
    with open(os.path.join(HERE, "input_small.bin"), "wb") as f:
        f.write(ref.encode_input_file(CONSTS, OBJECTS, input_frames()))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
