"""Frame files, input files and scene set-up: the host mirror of include/svb_files.h.

    frame files   IoState::write / read               rust/crates/file_frame/src/io_state.rs:31-75
    input files   InputReader / InputWriter           rust/crates/file_input/src/{reading,writing}.rs
    set-up        initialize_io_state                 rust/crates/core/src/initialization.rs:84-277
    keyframes     InputInterpolationPoint::new        rust/crates/xpu/src/frame_input.rs:66-135

Everything here calls the C++ implementation in lib/libsvb200.so (csrc/svb_files.cpp); there is no Python re-implementation
of the formats in the product (the independent one used as the checker lives in oracle/bincode_ref.py).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi
from . import cstructs as cs
from .types import ColliderTopology, FrameInput, GridNodes, InputConsts, IoState, Keyframe, Particles

OBJECT_PARTICLES, OBJECT_COLLIDER = 0, 1


class FileError(Exception):
    """A negative SVBF_* status (include/svb_files.h) with the library's message."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[{status}] {message}")
        self.status = status
        self.message = message


class SvbfObjectDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", C.c_int32), ("count", C.c_uint64), ("count2", C.c_uint64)]


class SvbfParticlesInput(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n", C.c_uint64), ("flags", cs.c_u32p), ("transforms", cs.c_f32p), ("sizes", cs.c_f32p), ("densities", cs.c_f32p),
                ("youngs_moduluses", cs.c_f32p), ("poissons_ratios", cs.c_f32p), ("initial_positions", cs.c_f32p), ("initial_velocities", cs.c_f32p),
                ("viscosities_dynamic", cs.c_f32p), ("viscosities_bulk", cs.c_f32p), ("exponents", cs.c_u32p), ("bulk_moduluses", cs.c_f32p),
                ("sand_alphas", cs.c_f32p), ("goal_positions", cs.c_f32p)]


class SvbfColliderInput(C.Structure):
    _fields_ = [("name", C.c_char_p), ("num_vertices", C.c_uint64), ("num_triangles", C.c_uint64), ("vertex_positions", cs.c_f32p),
                ("triangle_indices", cs.c_u32p), ("triangle_frictions", cs.c_f32p), ("triangle_dampings", cs.c_f32p)]


_bound = False


def _lib():
    global _bound
    L = abi.load()
    if _bound:
        return L
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.svbf_last_error.restype = C.c_char_p
    L.svbf_default_version.restype = C.c_char_p
    L.svbf_frame_write.restype = C.c_int32
    L.svbf_frame_write.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.POINTER(cs.SvbParticles), C.POINTER(cs.SvbGrid), u64p]
    L.svbf_frame_path.restype = C.c_int32
    L.svbf_frame_path.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_size_t]
    L.svbf_frame_open.restype = C.c_int32
    L.svbf_frame_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.svbf_frame_time.restype = C.c_double
    L.svbf_frame_time.argtypes = [vp]
    L.svbf_frame_particle_count.restype = C.c_uint64
    L.svbf_frame_particle_count.argtypes = [vp]
    L.svbf_frame_grid_count.restype = C.c_int64
    L.svbf_frame_grid_count.argtypes = [vp]
    L.svbf_frame_copy.restype = C.c_int32
    L.svbf_frame_copy.argtypes = [vp, C.POINTER(cs.SvbParticles), C.POINTER(cs.SvbGrid)]
    L.svbf_frame_close.restype = None
    L.svbf_frame_close.argtypes = [vp]
    L.svbf_input_open.restype = C.c_int32
    L.svbf_input_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.svbf_input_close.restype = None
    L.svbf_input_close.argtypes = [vp]
    for name in ("svbf_input_size", "svbf_input_frame_count", "svbf_input_total_particles", "svbf_input_total_vertices", "svbf_input_total_triangles"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [vp]
    L.svbf_input_object_count.restype = C.c_uint32
    L.svbf_input_object_count.argtypes = [vp]
    L.svbf_input_consts.restype = C.c_int32
    L.svbf_input_consts.argtypes = [vp, C.POINTER(cs.SvbConsts)]
    L.svbf_input_object.restype = C.c_int32
    L.svbf_input_object.argtypes = [vp, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_int32), u64p, u64p, u64p, u64p]
    L.svbf_input_topology.restype = C.c_int32
    L.svbf_input_topology.argtypes = [vp, C.POINTER(C.c_uint32), cs.c_u32p, cs.c_u32p, cs.c_u32p]
    L.svbf_input_initialize.restype = C.c_int32
    L.svbf_input_initialize.argtypes = [vp, C.POINTER(cs.SvbParticles)]
    L.svbf_input_keyframe.restype = C.c_int32
    L.svbf_input_keyframe.argtypes = [vp, C.c_uint64, C.POINTER(C.c_float * 3), cs.c_u32p, cs.c_f32p, cs.c_f32p, cs.c_f32p, cs.c_f32p]
    L.svbf_input_writer_open.restype = C.c_int32
    L.svbf_input_writer_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(cs.SvbConsts), C.POINTER(SvbfObjectDesc), C.c_uint32, C.POINTER(vp)]
    L.svbf_input_writer_frame.restype = C.c_int32
    L.svbf_input_writer_frame.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(SvbfParticlesInput), C.c_uint32, C.POINTER(SvbfColliderInput), C.c_uint32]
    L.svbf_input_writer_finish.restype = C.c_int32
    L.svbf_input_writer_finish.argtypes = [vp]
    _bound = True
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise FileError(rc, _lib().svbf_last_error().decode(errors="replace"))


def _ver(version: Optional[str]):
    return version.encode() if version is not None else None


def default_version() -> str:
    return _lib().svbf_default_version().decode()


# ---------------------------------------------------------------------------------------------------- frame files
def frame_path(cache_dir: str, frame: int) -> str:
    buf = C.create_string_buffer(4096)
    _check(_lib().svbf_frame_path(str(cache_dir).encode(), int(frame), buf, len(buf)))
    return buf.value.decode()


def write_frame(path: str, io_state: IoState, version: Optional[str] = None) -> int:
    """IoState::write (io_state.rs:32-65): `<dir>/temp.bin`, then rename.  Returns the bytes written."""
    L = _lib()
    p = io_state.particles.normalized()
    ps = cs.particles_struct(p)
    written = C.c_uint64(0)
    g = io_state.grid_nodes
    if g is None:
        _check(L.svbf_frame_write(str(path).encode(), _ver(version), C.c_double(io_state.time), C.byref(ps), None, C.byref(written)))
    else:
        ids = np.ascontiguousarray(g.node_ids, dtype=np.int32).reshape(-1, 3)
        bits = np.ascontiguousarray(g.collider_bits, dtype=np.uint32)
        masses = np.ascontiguousarray(g.masses, dtype=np.float32)
        vel = np.ascontiguousarray(g.velocities, dtype=np.float32).reshape(-1, 3)
        gs = cs.SvbGrid()
        gs.n = ids.shape[0]
        gs.node_ids = ids.ctypes.data_as(cs.c_i32p)
        gs.collider_bits = cs.uptr(bits)
        gs.masses = cs.fptr(masses)
        gs.velocities = cs.fptr(vel)
        gs.contributor_counts = None
        _check(L.svbf_frame_write(str(path).encode(), _ver(version), C.c_double(io_state.time), C.byref(ps), C.byref(gs), C.byref(written)))
    return int(written.value)


def read_frame(path: str, version: Optional[str] = None) -> IoState:
    """IoState::read (io_state.rs:67-75)."""
    L = _lib()
    h = C.c_void_p()
    _check(L.svbf_frame_open(str(path).encode(), _ver(version), C.byref(h)))
    try:
        n = int(L.svbf_frame_particle_count(h))
        p = Particles.empty(n)
        ps = cs.particles_struct(p)
        ng = int(L.svbf_frame_grid_count(h))
        grid = None
        if ng >= 0:
            grid = GridNodes(np.zeros((ng, 3), np.int32), np.zeros(ng, np.uint32), np.zeros(ng, np.float32), np.zeros((ng, 3), np.float32))
            gs = cs.SvbGrid()
            gs.n = ng
            gs.node_ids = grid.node_ids.ctypes.data_as(cs.c_i32p)
            gs.collider_bits = cs.uptr(grid.collider_bits)
            gs.masses = cs.fptr(grid.masses)
            gs.velocities = cs.fptr(grid.velocities)
            gs.contributor_counts = None
            _check(L.svbf_frame_copy(h, C.byref(ps), C.byref(gs)))
        else:
            _check(L.svbf_frame_copy(h, C.byref(ps), None))
        return IoState(float(L.svbf_frame_time(h)), p, grid)
    finally:
        L.svbf_frame_close(h)


# ---------------------------------------------------------------------------------------------------- input files
@dataclasses.dataclass
class InputObject:
    name: str
    kind: int          # OBJECT_PARTICLES | OBJECT_COLLIDER
    count: int         # particles | vertices
    count2: int        # triangles
    start: int         # first particle | first vertex
    start2: int        # first triangle


class InputFile:
    """InputReader + the parts of FrameInput / initialize_io_state that turn a recorded input into the back end's inputs."""

    def __init__(self, path: str, version: Optional[str] = None):
        L = _lib()
        self._h = C.c_void_p()
        _check(L.svbf_input_open(str(path).encode(), _ver(version), C.byref(self._h)))
        c = cs.SvbConsts()
        _check(L.svbf_input_consts(self._h, C.byref(c)))
        self.consts = InputConsts(grid_node_size=float(c.grid_node_size), leaf_size=float(c.leaf_size), leaf_threshold=int(c.leaf_threshold),
                                  simulation_scale=float(c.simulation_scale), frames_per_second=int(c.frames_per_second),
                                  domain_min=tuple(float(v) for v in c.domain_min), domain_max=tuple(float(v) for v in c.domain_max))
        self.size = int(L.svbf_input_size(self._h))
        self.n_frames = int(L.svbf_input_frame_count(self._h))
        self.total_particles = int(L.svbf_input_total_particles(self._h))
        self.total_vertices = int(L.svbf_input_total_vertices(self._h))
        self.total_triangles = int(L.svbf_input_total_triangles(self._h))
        self.objects: List[InputObject] = []
        for i in range(int(L.svbf_input_object_count(self._h))):
            name = C.create_string_buffer(1024)
            kind = C.c_int32()
            a, b, s, s2 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            _check(L.svbf_input_object(self._h, i, name, len(name), C.byref(kind), C.byref(a), C.byref(b), C.byref(s), C.byref(s2)))
            self.objects.append(InputObject(name.value.decode(), int(kind.value), int(a.value), int(b.value), int(s.value), int(s2.value)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib().svbf_input_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def topology(self) -> List[ColliderTopology]:
        """Collider topology from frame 0, colliders in name order (xpu/src/frame_input.rs:168-184)."""
        L = _lib()
        nc = C.c_uint32(0)
        _check(L.svbf_input_topology(self._h, C.byref(nc), None, None, None))
        k = int(nc.value)
        nv = np.zeros(max(k, 1), np.uint32)
        nt = np.zeros(max(k, 1), np.uint32)
        _check(L.svbf_input_topology(self._h, C.byref(nc), cs.uptr(nv), cs.uptr(nt), None))
        tris = np.zeros((max(int(nt[:k].sum()), 1), 3), np.uint32)
        _check(L.svbf_input_topology(self._h, C.byref(nc), cs.uptr(nv), cs.uptr(nt), cs.uptr(tris)))
        out, at = [], 0
        for c in range(k):
            out.append(ColliderTopology(num_vertices=int(nv[c]), triangles=tris[at:at + int(nt[c])].copy()))
            at += int(nt[c])
        return out

    def initialize_io_state(self) -> IoState:
        """initialize_io_state (core/src/initialization.rs:84-277): time 0, an empty grid."""
        p = Particles.empty(self.total_particles)
        ps = cs.particles_struct(p)
        _check(_lib().svbf_input_initialize(self._h, C.byref(ps)))
        empty = GridNodes(np.zeros((0, 3), np.int32), np.zeros(0, np.uint32), np.zeros(0, np.float32), np.zeros((0, 3), np.float32))
        return IoState(0.0, p, empty)

    def keyframe(self, frame: int) -> Keyframe:
        """InputInterpolationPoint::new (xpu/src/frame_input.rs:66-135) of recorded frame `frame`."""
        g = (C.c_float * 3)()
        n, nv, nt = self.total_particles, self.total_vertices, self.total_triangles
        flags = np.zeros(n, np.uint32)
        goals = np.zeros((n, 3), np.float32)
        verts = np.zeros((nv, 3), np.float32)
        fric = np.zeros(nt, np.float32)
        damp = np.zeros(nt, np.float32)
        _check(_lib().svbf_input_keyframe(self._h, int(frame), C.byref(g), cs.uptr(flags), cs.fptr(goals), cs.fptr(verts), cs.fptr(fric), cs.fptr(damp)))
        return Keyframe(gravity=tuple(float(v) for v in g), particle_flags=flags, particle_goal_positions=goals, vertex_positions=verts,
                        triangle_frictions=fric, triangle_dampings=damp)

    def frame_input(self, frame: int = 0) -> FrameInput:
        """FrameInput::new (xpu/src/frame_input.rs:158-204) with every recorded keyframe resident."""
        return FrameInput(consts=self.consts, colliders=self.topology(), keyframes=[self.keyframe(f) for f in range(self.n_frames)], frame=int(frame))


_PARTICLE_ATTRS = [("transforms", np.float32, 16), ("sizes", np.float32, 1), ("densities", np.float32, 1), ("youngs_moduluses", np.float32, 1),
                   ("poissons_ratios", np.float32, 1), ("initial_positions", np.float32, 3), ("initial_velocities", np.float32, 3),
                   ("viscosities_dynamic", np.float32, 1), ("viscosities_bulk", np.float32, 1), ("exponents", np.uint32, 1), ("bulk_moduluses", np.float32, 1),
                   ("sand_alphas", np.float32, 1), ("goal_positions", np.float32, 3)]


class InputWriter:
    """InputWriter (file_input/src/writing.rs:21-73).  objects: name -> ("particles", n) | ("collider", vertices, triangles)."""

    def __init__(self, path: str, consts: InputConsts, objects: Dict[str, Tuple], version: Optional[str] = None):
        L = _lib()
        self._objects = dict(objects)
        descs = (SvbfObjectDesc * max(len(objects), 1))()
        self._names = []
        for i, (name, o) in enumerate(objects.items()):
            b = name.encode()
            self._names.append(b)
            descs[i].name = b
            descs[i].kind = OBJECT_PARTICLES if o[0] == "particles" else OBJECT_COLLIDER
            descs[i].count = int(o[1])
            descs[i].count2 = int(o[2]) if len(o) > 2 else 0
        c = cs.consts_struct(consts)
        self._h = C.c_void_p()
        _check(L.svbf_input_writer_open(str(path).encode(), _ver(version), C.byref(c), descs, len(objects), C.byref(self._h)))

    def record_frame(self, gravity: Sequence[float], particles: Optional[Dict[str, dict]] = None, colliders: Optional[Dict[str, dict]] = None) -> None:
        """particles: name -> {"flags": .., "transforms": .. | None, ...}; colliders: name -> {"vertex_positions", "triangle_indices",
        "triangle_frictions", "triangle_dampings"}.  Verified against the header like InputFrame::verify (frame.rs:66-187)."""
        particles = particles or {}
        colliders = colliders or {}
        keep = []
        ps = (SvbfParticlesInput * max(len(particles), 1))()
        for i, (name, d) in enumerate(particles.items()):
            b = name.encode()
            keep.append(b)
            flags = np.ascontiguousarray(d["flags"], dtype=np.uint32)
            keep.append(flags)
            ps[i].name = b
            ps[i].n = flags.shape[0]
            ps[i].flags = cs.uptr(flags)
            for attr, dtype, k in _PARTICLE_ATTRS:
                a = d.get(attr)
                if a is None:
                    continue
                a = np.ascontiguousarray(a, dtype=dtype).reshape(-1)
                if a.shape[0] != flags.shape[0] * k:
                    raise FileError(-28, f"'{name}': {attr} holds {a.shape[0] // k} rows, flags {flags.shape[0]}")
                keep.append(a)
                setattr(ps[i], attr, cs.uptr(a) if dtype is np.uint32 else cs.fptr(a))
        cl = (SvbfColliderInput * max(len(colliders), 1))()
        for i, (name, d) in enumerate(colliders.items()):
            b = name.encode()
            v = np.ascontiguousarray(d["vertex_positions"], dtype=np.float32).reshape(-1, 3)
            t = np.ascontiguousarray(d["triangle_indices"], dtype=np.uint32).reshape(-1, 3)
            f = np.ascontiguousarray(d["triangle_frictions"], dtype=np.float32).reshape(-1)
            da = np.ascontiguousarray(d["triangle_dampings"], dtype=np.float32).reshape(-1)
            if f.shape[0] != t.shape[0] or da.shape[0] != t.shape[0]:
                raise FileError(-28, f"'{name}': frictions / dampings do not match the triangle count")
            keep += [b, v, t, f, da]
            cl[i].name = b
            cl[i].num_vertices = v.shape[0]
            cl[i].num_triangles = t.shape[0]
            cl[i].vertex_positions = cs.fptr(v)
            cl[i].triangle_indices = cs.uptr(t)
            cl[i].triangle_frictions = cs.fptr(f)
            cl[i].triangle_dampings = cs.fptr(da)
        g = (C.c_float * 3)(*[float(x) for x in gravity])
        _check(_lib().svbf_input_writer_frame(self._h, C.byref(g), ps, len(particles), cl, len(colliders)))

    def finish(self) -> None:
        """InputWriter::flush: frame index + its offset; the writer is gone afterwards."""
        if self._h:
            h, self._h = self._h, None
            _check(_lib().svbf_input_writer_finish(h))
