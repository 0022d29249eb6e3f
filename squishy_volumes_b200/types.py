"""Host-side mirrors of the reference's boundary types for the MPM substep path.

Nothing here computes; these are the plain containers that cross the back-end boundary
(`CpuState`/`GpuState::from_io_state` + `produce_next_state`), restated with numpy arrays:

* ``InputConsts``      file_input/src/header.rs:10-71
* ``ParticleFlags``    file_frame/src/particles.rs:23-33
* ``Particles``        file_frame/src/particles.rs:93-109 (+ flattened ``ParticleParameters`` :61-91)
* ``GridNodes``        file_frame/src/grid_nodes.rs:9-15
* ``IoState``          file_frame/src/io_state.rs:11-18
* ``Keyframe``         xpu/src/frame_input.rs:54-66 (``InputInterpolationPoint``)
* ``ColliderTopology`` mesh_util/src/mesh.rs:22-27 (``TopologyInput``)
* ``FrameInput``       xpu/src/frame_input.rs:33-52 (the parts the back end reads)
* ``RunParameters``    cpu/src/cpu_state.rs:138-143 (``CpuRunParameters``)
* ``Harness``          xpu/src/harness.rs:52-144 (cancel flag + progress)

(paths relative to the reference's ``rust/crates``).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np


class ParticleFlags:
    IS_SOLID = 1 << 0
    IS_FLUID = 1 << 1
    USE_VISCOSITY = 1 << 2
    USE_SAND_ALPHA = 1 << 3
    HAS_GOAL = 1 << 4
    TOMBSTONED = 1 << 5
    FAILED = 1 << 6


@dataclass
class InputConsts:
    grid_node_size: float = 0.5
    leaf_size: float = 1.0
    leaf_threshold: int = 16
    simulation_scale: float = 1.0
    frames_per_second: int = 24
    domain_min: Sequence[float] = (-100.0, -100.0, -100.0)
    domain_max: Sequence[float] = (100.0, 100.0, 100.0)

    def scaled_grid_node_size(self) -> float:
        return float(np.float32(self.grid_node_size) / np.float32(self.simulation_scale))

    def seconds_per_frame(self) -> float:
        return 1.0 / float(self.frames_per_second)


def _f32(a, shape_tail=()):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_tail:
        a = a.reshape((-1,) + tuple(shape_tail))
    return a


@dataclass
class Particles:
    """Struct of arrays, one row per particle, ORIGINAL particle order.

    ``position_gradients`` / ``velocity_gradients`` are (n, 3, 3) with the LAST-BUT-ONE axis the
    column index, i.e. ``F[p, c, r]`` — an array of columns exactly like the wire format
    ``[[f32; 3]; 3]`` of a column-major nalgebra ``Matrix3`` (use ``F[p].T`` for the math matrix).

    The Rust ``ParticleParameters`` enum is flattened; which fields are live follows the flags:
    IS_SOLID -> (mu, lambda[, sand_alpha if USE_SAND_ALPHA]); IS_FLUID -> (bulk_modulus, exponent);
    USE_VISCOSITY -> (viscosity_dynamic, viscosity_bulk).
    """

    flags: np.ndarray
    mass: np.ndarray
    initial_volume: np.ndarray
    mu_or_bulk_modulus: np.ndarray
    lambda_or_exponent: np.ndarray
    sand_alpha: np.ndarray
    viscosity_dynamic: np.ndarray
    viscosity_bulk: np.ndarray
    initial_positions: np.ndarray
    positions: np.ndarray
    position_gradients: np.ndarray
    velocities: np.ndarray
    velocity_gradients: np.ndarray
    elastic_energies: np.ndarray
    collider_bits: np.ndarray

    @property
    def n(self) -> int:
        return int(self.flags.shape[0])

    @staticmethod
    def empty(n: int) -> "Particles":
        z = lambda *s: np.zeros((n,) + s, dtype=np.float32)
        F = z(3, 3)
        F[:, 0, 0] = F[:, 1, 1] = F[:, 2, 2] = 1.0
        return Particles(
            flags=np.zeros(n, dtype=np.uint32), mass=z(), initial_volume=z(), mu_or_bulk_modulus=z(),
            lambda_or_exponent=z(), sand_alpha=z(), viscosity_dynamic=z(), viscosity_bulk=z(),
            initial_positions=z(3), positions=z(3), position_gradients=F, velocities=z(3),
            velocity_gradients=z(3, 3), elastic_energies=z(), collider_bits=np.zeros(n, dtype=np.uint32))

    def normalized(self) -> "Particles":
        """Contiguous arrays of the exact dtypes/shapes the C-ABI expects."""
        return Particles(
            flags=np.ascontiguousarray(self.flags, dtype=np.uint32),
            mass=_f32(self.mass), initial_volume=_f32(self.initial_volume),
            mu_or_bulk_modulus=_f32(self.mu_or_bulk_modulus), lambda_or_exponent=_f32(self.lambda_or_exponent),
            sand_alpha=_f32(self.sand_alpha), viscosity_dynamic=_f32(self.viscosity_dynamic),
            viscosity_bulk=_f32(self.viscosity_bulk),
            initial_positions=_f32(self.initial_positions, (3,)), positions=_f32(self.positions, (3,)),
            position_gradients=_f32(self.position_gradients, (3, 3)), velocities=_f32(self.velocities, (3,)),
            velocity_gradients=_f32(self.velocity_gradients, (3, 3)), elastic_energies=_f32(self.elastic_energies),
            collider_bits=np.ascontiguousarray(self.collider_bits, dtype=np.uint32))

    def copy(self) -> "Particles":
        return Particles(**{f.name: getattr(self, f.name).copy() for f in dataclasses.fields(self)})

    def select(self, idx) -> "Particles":
        return Particles(**{f.name: np.ascontiguousarray(getattr(self, f.name)[idx]) for f in dataclasses.fields(self)})

    @staticmethod
    def concatenate(parts: Sequence["Particles"]) -> "Particles":
        return Particles(**{f.name: np.concatenate([getattr(p, f.name) for p in parts], axis=0)
                            for f in dataclasses.fields(Particles)})


@dataclass
class GridNodes:
    node_ids: np.ndarray        # (g, 3) int32
    collider_bits: np.ndarray   # (g,) uint32
    masses: np.ndarray          # (g,) float32
    velocities: np.ndarray      # (g, 3) float32  (the reference spells the field `velocites`)
    contributor_counts: Optional[np.ndarray] = None  # oracle-only diagnostic


@dataclass
class IoState:
    time: float
    particles: Particles
    grid_nodes: Optional[GridNodes] = None


@dataclass
class Keyframe:
    """One input interpolation point; positions already divided by ``simulation_scale``."""

    gravity: Sequence[float] = (0.0, 0.0, -9.8)
    particle_flags: Optional[np.ndarray] = None            # (n,) uint32, original particle order
    particle_goal_positions: Optional[np.ndarray] = None   # (n, 3)
    vertex_positions: Optional[np.ndarray] = None          # (V, 3) all colliders concatenated
    triangle_frictions: Optional[np.ndarray] = None        # (T,)
    triangle_dampings: Optional[np.ndarray] = None         # (T,)


@dataclass
class ColliderTopology:
    num_vertices: int
    triangles: np.ndarray  # (t, 3) uint32, LOCAL vertex indices of this collider


@dataclass
class FrameInput:
    """What the back end reads from ``xpu::FrameInput``: consts, collider topology (from the
    collider inputs of frame 0) and keyframes ``a`` (= ``frame``) / ``b`` (= ``frame + 1`` or None).
    ``keyframes`` plays the role of the recorded input file; ``load`` mirrors
    ``FrameInput::load`` (xpu/src/frame_input.rs:206-232, clamping to the last frame :289-293)."""

    consts: InputConsts
    colliders: List[ColliderTopology] = field(default_factory=list)
    keyframes: List[Keyframe] = field(default_factory=lambda: [Keyframe()])
    frame: int = 0

    def load(self, frame: int) -> None:
        self.frame = int(frame)

    def a_index(self) -> int:
        return min(self.frame, len(self.keyframes) - 1)

    def a(self) -> Keyframe:
        return self.keyframes[self.a_index()]

    def b(self) -> Optional[Keyframe]:
        i = self.a_index()
        return self.keyframes[i + 1] if i + 1 < len(self.keyframes) else None

    def num_vertices(self) -> int:
        return sum(c.num_vertices for c in self.colliders)

    def num_triangles(self) -> int:
        return sum(int(np.asarray(c.triangles).reshape(-1, 3).shape[0]) for c in self.colliders)


@dataclass
class RunParameters:
    target_time: float
    max_time_step: float
    adaptive_time_steps: bool = False
    store_grid: bool = False


class Harness:
    """Cancel flag + progress callback (xpu/src/harness.rs:52-144, reduced to what the back end uses)."""

    def __init__(self, progress: Optional[Callable[[int], None]] = None):
        self._cancel = np.zeros(1, dtype=np.int32)
        self.progress = progress

    def cancel(self) -> None:
        self._cancel[0] = 1

    @property
    def cancelled(self) -> bool:
        return bool(self._cancel[0])

    def cancel_pointer(self):
        return self._cancel.ctypes.data


class SimulationError(Exception):
    """Inner (simulation-level) error: the returned state is still valid and is stored by the caller
    (cpu/src/cpu_state.rs:178-184, core/src/compute_thread.rs:165-169)."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class FatalError(Exception):
    """Outer error: cancelled / frame-input / zero time step (cpu/src/errors.rs:9-36)."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status
