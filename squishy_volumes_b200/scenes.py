"""Synthetic scenes for the five BASELINE.json configs, parametrised by size so the
same generators feed the parity tests (tiny) and bench.py (full).  Pure numpy; no device code.

Modelled on the reference's generators: ``gpu/src/test_data/particles.rs:114-187`` (``Neat(spacing)``
lattice = 8 particles per cell at spacing h/2), ``core/src/initialization.rs:202-259`` (volume =
size^3, mass = volume*density, mu/lambda from E/nu) and ``gpu_cli/src/main.rs:83-91`` (seed 1234).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .types import (ColliderTopology, FrameInput, InputConsts, IoState, Keyframe, ParticleFlags, Particles)

SEED = 1234


def lame(E: float, nu: float) -> Tuple[np.float32, np.float32]:
    """util/src/elastic.rs:52-64 in f32."""
    E = np.float32(E)
    nu = np.float32(nu)
    one = np.float32(1.0)
    two = np.float32(2.0)
    mu = E / two / (one + nu)
    lam = E * nu / (one + nu) / (one - two * nu)
    return mu, lam


def lattice(counts: Sequence[int], spacing: float, origin: Sequence[float], jitter: float = 0.1,
            seed: int = SEED) -> np.ndarray:
    """Regular lattice, k fastest, jittered by +-jitter*spacing with a counter-based generator."""
    nx, ny, nz = (int(c) for c in counts)
    i = np.arange(nx, dtype=np.float32)
    j = np.arange(ny, dtype=np.float32)
    k = np.arange(nz, dtype=np.float32)
    p = np.empty((nx, ny, nz, 3), dtype=np.float32)
    p[..., 0] = (np.float32(origin[0]) + (i + np.float32(0.5)) * np.float32(spacing))[:, None, None]
    p[..., 1] = (np.float32(origin[1]) + (j + np.float32(0.5)) * np.float32(spacing))[None, :, None]
    p[..., 2] = (np.float32(origin[2]) + (k + np.float32(0.5)) * np.float32(spacing))[None, None, :]
    p = p.reshape(-1, 3)
    if jitter:
        rng = np.random.Generator(np.random.Philox(seed))
        p += (rng.random(p.shape, dtype=np.float32) * np.float32(2) - np.float32(1)) * np.float32(jitter * spacing)
    return p


@dataclass
class Material:
    kind: str = "solid"            # "solid" | "sand" | "fluid"
    density: float = 1000.0
    youngs_modulus: float = 1e4
    poissons_ratio: float = 0.3
    sand_alpha: float = 0.3        # gpu/src/sand/test.rs:95
    bulk_modulus: float = 1000.0   # util/src/elastic.rs:691
    exponent: int = 7
    viscosity: Optional[Tuple[float, float]] = None  # (dynamic, bulk)


def make_particles(positions: np.ndarray, spacing: float, material: Material,
                   velocity: Sequence[float] = (0.0, 0.0, 0.0)) -> Particles:
    n = positions.shape[0]
    p = Particles.empty(n)
    p.positions[:] = positions
    p.initial_positions[:] = positions
    p.velocities[:] = np.asarray(velocity, dtype=np.float32)
    vol = np.float32(spacing) ** 3
    p.initial_volume[:] = vol
    p.mass[:] = vol * np.float32(material.density)
    flags = 0
    if material.kind in ("solid", "sand"):
        mu, lam = lame(material.youngs_modulus, material.poissons_ratio)
        flags |= ParticleFlags.IS_SOLID
        p.mu_or_bulk_modulus[:] = mu
        p.lambda_or_exponent[:] = lam
        if material.kind == "sand":
            flags |= ParticleFlags.USE_SAND_ALPHA
            p.sand_alpha[:] = material.sand_alpha
    elif material.kind == "fluid":
        flags |= ParticleFlags.IS_FLUID
        p.mu_or_bulk_modulus[:] = material.bulk_modulus
        p.lambda_or_exponent[:] = float(material.exponent)
    else:
        raise ValueError(material.kind)
    if material.viscosity is not None:
        flags |= ParticleFlags.USE_VISCOSITY
        p.viscosity_dynamic[:] = material.viscosity[0]
        p.viscosity_bulk[:] = material.viscosity[1]
    p.flags[:] = flags
    return p


# ----------------------------------------------------------------------------- collider meshes
def plane_mesh(z: float, half: float) -> Tuple[np.ndarray, np.ndarray]:
    """Two triangles, normal +z (the ground plane of config 1 is a mesh collider)."""
    v = np.array([[-half, -half, z], [half, -half, z], [half, half, z], [-half, half, z]], dtype=np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    return v, t


def box_mesh(lo: Sequence[float], hi: Sequence[float], inward: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Closed 12-triangle box; normals point inward (a container) or outward."""
    lo = np.asarray(lo, dtype=np.float32)
    hi = np.asarray(hi, dtype=np.float32)
    v = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]],
                 dtype=np.float32)
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (3, 0, 4, 7)]  # outward
    t = []
    for a, b, c, d in quads:
        t += [[a, b, c], [a, c, d]]
    t = np.array(t, dtype=np.uint32)
    if inward:
        t = t[:, ::-1].copy()
    return v, t


def torus_mesh(R: float, r: float, nu: int, nv: int, center: Sequence[float] = (0, 0, 0)) -> Tuple[np.ndarray, np.ndarray]:
    """Closed manifold torus around the z axis, outward normals (stand-in for gpu/src/torus.rs)."""
    u = np.arange(nu, dtype=np.float64) * (2 * np.pi / nu)
    w = np.arange(nv, dtype=np.float64) * (2 * np.pi / nv)
    U, W = np.meshgrid(u, w, indexing="ij")
    x = (R + r * np.cos(W)) * np.cos(U) + center[0]
    y = (R + r * np.cos(W)) * np.sin(U) + center[1]
    z = r * np.sin(W) + center[2]
    v = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    t = []
    for i in range(nu):
        for j in range(nv):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            t += [[a, b, c], [a, c, d]]
    return v, np.array(t, dtype=np.uint32)


@dataclass
class Scene:
    name: str
    io_state: IoState
    frame_input: FrameInput
    time_step: float
    description: str

    @property
    def n(self) -> int:
        return self.io_state.particles.n


def _frame_input(consts: InputConsts, meshes: List[Tuple[np.ndarray, np.ndarray]], n: int, gravity, frictions,
                 dampings, n_keyframes: int = 2, moving: Optional[Sequence[float]] = None) -> FrameInput:
    colliders = [ColliderTopology(num_vertices=v.shape[0], triangles=t) for v, t in meshes]
    verts = np.concatenate([v for v, _ in meshes], axis=0) if meshes else np.zeros((0, 3), np.float32)
    fr = np.concatenate([np.full(t.shape[0], f, np.float32) for (_, t), f in zip(meshes, frictions)]) if meshes else np.zeros(0, np.float32)
    da = np.concatenate([np.full(t.shape[0], d, np.float32) for (_, t), d in zip(meshes, dampings)]) if meshes else np.zeros(0, np.float32)
    kfs = []
    for f in range(n_keyframes):
        vp = verts.copy()
        if moving is not None:
            vp += np.asarray(moving, np.float32) * np.float32(f)
        kfs.append(Keyframe(gravity=tuple(gravity), particle_flags=np.zeros(n, np.uint32),
                            particle_goal_positions=np.zeros((n, 3), np.float32), vertex_positions=vp,
                            triangle_frictions=fr.copy(), triangle_dampings=da.copy()))
    return FrameInput(consts=consts, colliders=colliders, keyframes=kfs, frame=0)


def _consts(h: float, extent: float) -> InputConsts:
    # leaf_size = 2*grid_node_size, leaf_threshold 16: input_capture.py:45-49
    return InputConsts(grid_node_size=h, leaf_size=2.0 * h, leaf_threshold=16, simulation_scale=1.0,
                       frames_per_second=24, domain_min=(-extent,) * 3, domain_max=(extent,) * 3)


def elastic_cube(side: int = 46, h: float = 0.04, n_keyframes: int = 12) -> Scene:
    """Config 1: ~100 k-particle Neo-Hookean cube dropped on a ground-plane mesh collider."""
    sp = h / 2
    pos = lattice((side,) * 3, sp, (-side * sp / 2, -side * sp / 2, -1.0 + 1.5 * h))
    p = make_particles(pos, sp, Material("solid", 1000.0, 1e4, 0.3))
    fi = _frame_input(_consts(h, 100.0), [plane_mesh(-1.0, 4.0)], p.n, (0, 0, -9.8), [0.5], [0.0], n_keyframes)
    return Scene("elastic_cube", IoState(0.0, p), fi, 1e-3,
                 f"{p.n}-particle elastic cube (E=1e4, nu=0.3) on a 2-triangle ground plane, h={h}")


def jelly_collision(side: int = 80, h: float = 0.04, n_keyframes: int = 4, length: int = 1) -> Scene:
    """Config 2: two Neo-Hookean blocks (2*side^3 particles; side=80 -> 1.02 M) approaching at +-1 m/s.
    `length` stretches both blocks along x (length*side x side x side each): the weak-scaling workload of
    the slab decomposition, 1.02 M particles and a constant 80 x 80-cell cut face per GPU."""
    sp = h / 2
    gap = 2 * h
    sx = side * length
    a = make_particles(lattice((sx, side, side), sp, (-sx * sp - gap / 2, -side * sp / 2, -side * sp / 2), seed=SEED),
                       sp, Material("solid", 1000.0, 1e4, 0.3), velocity=(1.0, 0.0, 0.0))
    b = make_particles(lattice((sx, side, side), sp, (gap / 2, -side * sp / 2, -side * sp / 2), seed=SEED + 1),
                       sp, Material("solid", 1000.0, 1e5, 0.3), velocity=(-1.0, 0.0, 0.0))
    p = Particles.concatenate([a, b])
    fi = _frame_input(_consts(h, 100.0), [], p.n, (0, 0, 0), [], [], n_keyframes)
    return Scene("jelly_collision", IoState(0.0, p), fi, 1e-3,
                 f"{p.n}-particle two-block Neo-Hookean collision (E=1e4 / 1e5), h={h}, no collider")


def sand_torus(side: int = 200, h: float = 0.04, n_keyframes: int = 4, contact: bool = False) -> Scene:
    """Config 3: sand block (Drucker-Prager return mapping) falling onto a torus mesh collider.
    `contact`: the block starts half a cell INTO the top of the torus at 8 m/s, so that the first substeps already
    classify sides, push particles out and yield (the full-size measurements are short runs)."""
    sp = h / 2
    L = side * sp
    pos = lattice((side,) * 3, sp, (-L / 2, -L / 2, -3.0 * h if contact else 0.0))
    p = make_particles(pos, sp, Material("sand", 1600.0, 1e6, 0.3, sand_alpha=0.3), velocity=(0, 0, -8.0 if contact else -1.0))
    torus = torus_mesh(0.35 * L, 0.12 * L, 48, 24, center=(0, 0, -0.12 * L - 2.5 * h))
    fi = _frame_input(_consts(h, 100.0), [torus], p.n, (0, 0, -9.8), [0.4], [0.0], n_keyframes)
    return Scene("sand_torus", IoState(0.0, p), fi, 1e-4,
                 f"{p.n}-particle sand block (E=1e6, alpha=0.3) over a torus collider, h={h}")


def dam_break(nx: int = 400, ny: int = 200, nz: int = 200, h: float = 0.04, n_keyframes: int = 4,
              viscous: bool = False, contact: bool = False) -> Scene:
    """Config 4: weakly compressible fluid column inside a closed box collider.
    `contact`: the box hugs the column (0.1 h instead of 3 h of clearance) and the column moves at (2, 0, -6) m/s, so that
    the walls are within the collider's accept distance from the first substep and particles cross them within ten."""
    sp = h / 2
    lo = (-nx * sp, -ny * sp / 2, 0.0)
    pos = lattice((nx, ny, nz), sp, lo)
    mat = Material("fluid", 1000.0, bulk_modulus=1000.0, exponent=7, viscosity=(0.5, 0.1) if viscous else None)
    p = make_particles(pos, sp, mat, velocity=(2.0, 0.0, -6.0) if contact else (0.0, 0.0, 0.0))
    m = (0.1 if contact else 3.0) * h
    box = box_mesh((lo[0] - m, lo[1] - m, lo[2] - m), (lo[0] + 2 * nx * sp + m, lo[1] + ny * sp + m, nz * sp * 1.5), inward=True)
    fi = _frame_input(_consts(h, 100.0), [box], p.n, (0, 0, -9.8), [0.0], [0.0], n_keyframes)
    return Scene("dam_break", IoState(0.0, p), fi, 2e-4,
                 f"{p.n}-particle weakly compressible dam break (K=1000, gamma=7) in a box collider, h={h}")


def mixed(side: int = 400, h: float = 0.04, brick: int = 16, n_keyframes: int = 4, contact: bool = False) -> Scene:
    """Config 5: interleaved bricks, half Neo-Hookean, quarter sand, quarter fluid; plane + torus colliders.
    `contact`: the block starts 0.1 h above the plane (overlapping the top of the torus) at 4 m/s."""
    sp = h / 2
    L = side * sp
    pos = lattice((side,) * 3, sp, (-L / 2, -L / 2, -3.15 * h if contact else 0.0))
    n = pos.shape[0]
    ii = (np.arange(side) // brick)
    code = (ii[:, None, None] + ii[None, :, None] + ii[None, None, :]).reshape(-1) % 4  # 0,1 solid; 2 sand; 3 fluid
    parts = []
    order = []
    for c, mat in ((0, Material("solid", 1000.0, 1e4, 0.3)), (1, Material("solid", 1000.0, 1e5, 0.3)),
                   (2, Material("sand", 1600.0, 1e5, 0.3, sand_alpha=0.3)), (3, Material("fluid", 1000.0, bulk_modulus=1000.0, exponent=7))):
        sel = np.nonzero(code == c)[0]
        order.append(sel)
        parts.append(make_particles(pos[sel], sp, mat, velocity=(0, 0, -4.0 if contact else -0.5)))
    p = Particles.concatenate(parts)
    inv = np.argsort(np.concatenate(order), kind="stable")
    p = p.select(inv)  # back to lattice order so the input is spatially coherent like a Blender capture
    plane = plane_mesh(-3 * h, 2 * L)
    torus = torus_mesh(0.3 * L, 0.08 * L, 48, 24, center=(0, 0, -0.08 * L - 2.5 * h + 0.0))
    fi = _frame_input(_consts(h, 200.0), [plane, torus], n, (0, 0, -9.8), [0.5, 0.3], [0.0, 0.0], n_keyframes)
    return Scene("mixed", IoState(0.0, p), fi, 1e-4,
                 f"{n}-particle mixed solid/sand/fluid bricks with plane + torus colliders, h={h}")


SCENES = {
    "elastic_cube": elastic_cube,
    "jelly_collision": jelly_collision,
    "sand_torus": sand_torus,
    "dam_break": dam_break,
    "mixed": mixed,
}
