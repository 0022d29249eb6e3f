"""ctypes mirrors of the plain-C structs declared in include/svb200.h, plus marshalling helpers
from the numpy containers in ``types.py``.  No compute, no library loading."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .types import FrameInput, GridNodes, InputConsts, Keyframe, Particles

c_f32p = C.POINTER(C.c_float)
c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)


class SvbConsts(C.Structure):
    _fields_ = [("grid_node_size", C.c_float), ("leaf_size", C.c_float), ("leaf_threshold", C.c_uint32),
                ("simulation_scale", C.c_float), ("frames_per_second", C.c_uint32),
                ("domain_min", C.c_float * 3), ("domain_max", C.c_float * 3)]


class SvbParticles(C.Structure):
    _fields_ = [("n", C.c_uint64), ("flags", c_u32p), ("mass", c_f32p), ("initial_volume", c_f32p),
                ("mu_or_bulk_modulus", c_f32p), ("lambda_or_exponent", c_f32p), ("sand_alpha", c_f32p),
                ("viscosity_dynamic", c_f32p), ("viscosity_bulk", c_f32p), ("initial_positions", c_f32p),
                ("positions", c_f32p), ("position_gradients", c_f32p), ("velocities", c_f32p),
                ("velocity_gradients", c_f32p), ("elastic_energies", c_f32p), ("collider_bits", c_u32p)]


class SvbKeyframe(C.Structure):
    _fields_ = [("gravity", C.c_float * 3), ("particle_flags", c_u32p), ("particle_goal_positions", c_f32p),
                ("vertex_positions", c_f32p), ("triangle_frictions", c_f32p), ("triangle_dampings", c_f32p)]


class SvbGrid(C.Structure):
    _fields_ = [("n", C.c_uint64), ("node_ids", c_i32p), ("collider_bits", c_u32p), ("masses", c_f32p),
                ("velocities", c_f32p), ("contributor_counts", c_u32p)]


def fptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_f32p) if a is not None else None


def uptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_u32p) if a is not None else None


def consts_struct(c: InputConsts) -> SvbConsts:
    s = SvbConsts()
    s.grid_node_size = c.grid_node_size
    s.leaf_size = c.leaf_size
    s.leaf_threshold = int(c.leaf_threshold)
    s.simulation_scale = c.simulation_scale
    s.frames_per_second = int(c.frames_per_second)
    for i in range(3):
        s.domain_min[i] = c.domain_min[i]
        s.domain_max[i] = c.domain_max[i]
    return s


def particles_struct(p: Particles) -> SvbParticles:
    """``p`` must be ``Particles.normalized()`` and be kept alive by the caller."""
    s = SvbParticles()
    s.n = p.n
    s.flags = uptr(p.flags)
    s.collider_bits = uptr(p.collider_bits)
    for name in ("mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha",
                 "viscosity_dynamic", "viscosity_bulk", "initial_positions", "positions", "position_gradients",
                 "velocities", "velocity_gradients", "elastic_energies"):
        setattr(s, name, fptr(getattr(p, name)))
    return s


def keyframe_struct(k: Keyframe, n_particles: int, n_vertices: int, n_triangles: int):
    """Returns (struct, keepalive list)."""
    s = SvbKeyframe()
    keep = []
    for i in range(3):
        s.gravity[i] = float(k.gravity[i])

    def arr(a, dtype, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype).reshape(shape)
        keep.append(a)
        return a

    fl = arr(k.particle_flags, np.uint32, (n_particles,))
    gp = arr(k.particle_goal_positions, np.float32, (n_particles, 3))
    vp = arr(k.vertex_positions, np.float32, (n_vertices, 3)) if n_vertices else None
    fr = arr(k.triangle_frictions, np.float32, (n_triangles,)) if n_triangles else None
    da = arr(k.triangle_dampings, np.float32, (n_triangles,)) if n_triangles else None
    if n_vertices and vp is None:
        raise ValueError("keyframe lacks vertex_positions for a scene with colliders")
    if n_triangles and fr is None:
        fr = arr(np.zeros(n_triangles), np.float32, (n_triangles,))
    if n_triangles and da is None:
        da = arr(np.zeros(n_triangles), np.float32, (n_triangles,))
    s.particle_flags = uptr(fl)
    s.particle_goal_positions = fptr(gp)
    s.vertex_positions = fptr(vp)
    s.triangle_frictions = fptr(fr)
    s.triangle_dampings = fptr(da)
    return s, keep


def topology_arrays(fi: FrameInput):
    nv = np.array([c.num_vertices for c in fi.colliders], dtype=np.uint32)
    tris = [np.ascontiguousarray(c.triangles, dtype=np.uint32).reshape(-1, 3) for c in fi.colliders]
    nt = np.array([t.shape[0] for t in tris], dtype=np.uint32)
    flat = np.concatenate(tris, axis=0) if tris else np.zeros((0, 3), dtype=np.uint32)
    return nv, nt, np.ascontiguousarray(flat)


def alloc_grid(n: int, with_counts: bool = False):
    g = GridNodes(node_ids=np.zeros((n, 3), dtype=np.int32), collider_bits=np.zeros(n, dtype=np.uint32),
                  masses=np.zeros(n, dtype=np.float32), velocities=np.zeros((n, 3), dtype=np.float32),
                  contributor_counts=np.zeros(n, dtype=np.uint32) if with_counts else None)
    s = SvbGrid()
    s.n = n
    s.node_ids = g.node_ids.ctypes.data_as(c_i32p)
    s.collider_bits = uptr(g.collider_bits)
    s.masses = fptr(g.masses)
    s.velocities = fptr(g.velocities)
    s.contributor_counts = uptr(g.contributor_counts)
    return g, s
