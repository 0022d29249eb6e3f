"""TEST INFRASTRUCTURE — format oracle for the reference's files (only tests/ may import this).

A plain-Python restatement of how serde + bincode 1.3.3 (default options: little endian, fixed-width integers,
u64 lengths, u8 Option tag, u32 enum variant index, arrays without length, usize as u64) lay out the reference's
types, written independently of squishy_volumes_b200/csrc/svb_files.cpp so the two can be compared byte for byte:

  IoState / Particles / ParticleParameters / GridNodes   rust/crates/file_frame/src/{io_state,particles,grid_nodes}.rs
  InputHeader / InputConsts / InputObject                 rust/crates/file_input/src/header.rs:10-90
  InputFrame / ParticlesInput / ColliderInput             rust/crates/file_input/src/{frame.rs:10-49, collider_inputs.rs:9-18}
  container (magic, version, frame index)                 rust/crates/file_util/src/lib.rs:26-101, file_input/src/writing.rs:27-68
  initialize_io_state                                     rust/crates/core/src/initialization.rs:84-277
  InputInterpolationPoint::new                            rust/crates/xpu/src/frame_input.rs:66-135

Parity status: the Rust crates cannot be built here (no cargo), so these layouts are pinned by the serde derive order
in the cited files and the bincode specification, not by bytes produced by the reference: "parity unpinned".
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional

import numpy as np

FRAME_MAGIC = b"Squishy Volumes Frame File Magic"
INPUT_MAGIC = b"Squishy Volumes Input File Magic"
VERSION = "0.3.4"   # rust/crates/file_util/Cargo.toml:3
IS_SOLID, IS_FLUID, USE_VISCOSITY, USE_SAND_ALPHA, HAS_GOAL, TOMBSTONED, FAILED = 1, 2, 4, 8, 16, 32, 64


def version_bytes(version: str = VERSION) -> bytes:
    b = version.encode()
    assert len(b) <= 64
    return b + b"\0" * (64 - len(b))


class Writer:
    def __init__(self):
        self.parts: List[bytes] = []
        self.size = 0

    def raw(self, b: bytes):
        self.parts.append(bytes(b))
        self.size += len(b)

    def u8(self, v): self.raw(struct.pack("<B", v))
    def u32(self, v): self.raw(struct.pack("<I", v))
    def i32(self, v): self.raw(struct.pack("<i", v))
    def u64(self, v): self.raw(struct.pack("<Q", v))
    def f32(self, v): self.raw(struct.pack("<f", v))
    def f64(self, v): self.raw(struct.pack("<d", v))

    def string(self, s: str):
        b = s.encode()
        self.u64(len(b))
        self.raw(b)

    def vec(self, a, dtype, tail=()):
        a = np.ascontiguousarray(a, dtype=dtype).reshape((-1,) + tuple(tail))
        self.u64(a.shape[0])
        self.raw(a.astype(np.dtype(dtype).newbyteorder("<"), copy=False).tobytes())

    def opt_vec(self, a, dtype, tail=()):
        if a is None:
            self.u8(0)
        else:
            self.u8(1)
            self.vec(a, dtype, tail)

    def bytes(self) -> bytes:
        return b"".join(self.parts)


class Reader:
    def __init__(self, b: bytes, pos: int = 0):
        self.b, self.pos = b, pos

    def _take(self, fmt):
        n = struct.calcsize(fmt)
        if self.pos + n > len(self.b):
            raise ValueError("unexpected end of input")
        v = struct.unpack_from(fmt, self.b, self.pos)[0]
        self.pos += n
        return v

    def u8(self): return self._take("<B")
    def u32(self): return self._take("<I")
    def i32(self): return self._take("<i")
    def u64(self): return self._take("<Q")
    def f32(self): return self._take("<f")
    def f64(self): return self._take("<d")

    def string(self) -> str:
        n = self.u64()
        s = self.b[self.pos:self.pos + n].decode()
        self.pos += n
        return s

    def vec(self, dtype, tail=()):
        n = self.u64()
        k = int(np.prod(tail)) if tail else 1
        nbytes = n * k * np.dtype(dtype).itemsize
        if self.pos + nbytes > len(self.b):
            raise ValueError("unexpected end of input")
        a = np.frombuffer(self.b, dtype=np.dtype(dtype).newbyteorder("<"), count=n * k, offset=self.pos).reshape((n,) + tuple(tail)).astype(dtype)
        self.pos += nbytes
        return a

    def opt_vec(self, dtype, tail=()):
        return self.vec(dtype, tail) if self.u8() else None


# ------------------------------------------------------------------------------------------------ frame files
def encode_io_state(time: float, p, grid=None, version: str = VERSION) -> bytes:
    """`p`: squishy_volumes_b200.types.Particles-like (flattened parameters); `grid`: GridNodes-like or None."""
    w = Writer()
    w.raw(FRAME_MAGIC)
    w.raw(version_bytes(version))
    w.f64(time)
    n = int(np.asarray(p.flags).shape[0])
    w.vec(p.flags, np.uint32)
    w.u64(n)
    for i in range(n):                                   # ParticleParameters, particles.rs:62-84
        fl = int(p.flags[i])
        w.f32(p.mass[i])
        w.f32(p.initial_volume[i])
        if fl & USE_VISCOSITY:
            w.u8(1); w.f32(p.viscosity_dynamic[i]); w.f32(p.viscosity_bulk[i])
        else:
            w.u8(0)
        if fl & IS_FLUID:                                # Fluid { exponent: i32, bulk_modulus: f32 }
            w.u32(1); w.i32(int(p.lambda_or_exponent[i])); w.f32(p.mu_or_bulk_modulus[i])
        else:                                            # Solid { mu, lambda, sand_alpha: Option<f32> }
            w.u32(0); w.f32(p.mu_or_bulk_modulus[i]); w.f32(p.lambda_or_exponent[i])
            if fl & USE_SAND_ALPHA:
                w.u8(1); w.f32(p.sand_alpha[i])
            else:
                w.u8(0)
    w.vec(p.elastic_energies, np.float32)
    w.vec(p.collider_bits, np.uint32)
    w.vec(p.positions, np.float32, (3,))
    w.vec(p.position_gradients, np.float32, (3, 3))
    w.vec(p.velocities, np.float32, (3,))
    w.vec(p.velocity_gradients, np.float32, (3, 3))
    w.vec(p.initial_positions, np.float32, (3,))
    if grid is None:
        w.u8(0)
    else:
        w.u8(1)
        w.vec(grid.node_ids, np.int32, (3,))
        w.vec(grid.collider_bits, np.uint32)
        w.vec(grid.masses, np.float32)
        w.vec(grid.velocities, np.float32, (3,))
    return w.bytes()


def decode_io_state(b: bytes, version: str = VERSION) -> dict:
    if b[:32] != FRAME_MAGIC:
        raise ValueError("magic mismatch")
    if b[32:96] != version_bytes(version):
        raise ValueError("version mismatch")
    r = Reader(b, 96)
    out = {"time": r.f64(), "flags": r.vec(np.uint32)}
    n = r.u64()
    cols = {k: np.zeros(n, np.float32) for k in ("mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha", "viscosity_dynamic", "viscosity_bulk")}
    for i in range(n):
        cols["mass"][i] = r.f32()
        cols["initial_volume"][i] = r.f32()
        if r.u8():
            cols["viscosity_dynamic"][i] = r.f32(); cols["viscosity_bulk"][i] = r.f32()
        variant = r.u32()
        if variant == 0:
            cols["mu_or_bulk_modulus"][i] = r.f32(); cols["lambda_or_exponent"][i] = r.f32()
            if r.u8():
                cols["sand_alpha"][i] = r.f32()
        else:
            cols["lambda_or_exponent"][i] = float(r.i32()); cols["mu_or_bulk_modulus"][i] = r.f32()
    out.update(cols)
    out["elastic_energies"] = r.vec(np.float32)
    out["collider_bits"] = r.vec(np.uint32)
    out["positions"] = r.vec(np.float32, (3,))
    out["position_gradients"] = r.vec(np.float32, (3, 3))
    out["velocities"] = r.vec(np.float32, (3,))
    out["velocity_gradients"] = r.vec(np.float32, (3, 3))
    out["initial_positions"] = r.vec(np.float32, (3,))
    if r.u8():
        out["grid"] = {"node_ids": r.vec(np.int32, (3,)), "collider_bits": r.vec(np.uint32), "masses": r.vec(np.float32), "velocities": r.vec(np.float32, (3,))}
    else:
        out["grid"] = None
    assert r.pos == len(b), "trailing bytes"
    return out


# ------------------------------------------------------------------------------------------------ input files
PARTICLE_ATTRS = [  # ParticlesInput after `flags`, in serde order (frame.rs:12-25): (name, dtype, tail)
    ("transforms", np.float32, (4, 4)), ("sizes", np.float32, ()), ("densities", np.float32, ()), ("youngs_moduluses", np.float32, ()),
    ("poissons_ratios", np.float32, ()), ("initial_positions", np.float32, (3,)), ("initial_velocities", np.float32, (3,)),
    ("viscosities_dynamic", np.float32, ()), ("viscosities_bulk", np.float32, ()), ("exponents", np.uint32, ()), ("bulk_moduluses", np.float32, ()),
    ("sand_alphas", np.float32, ()), ("goal_positions", np.float32, (3,))]


def encode_header(w: Writer, consts: dict, objects: Dict[str, tuple]):
    """objects: name -> ("particles", n) | ("collider", num_vertices, num_triangles); a BTreeMap: sorted by name."""
    w.f32(consts["grid_node_size"]); w.f32(consts["leaf_size"]); w.u32(consts["leaf_threshold"]); w.f32(consts["simulation_scale"])
    w.u32(consts["frames_per_second"])
    for v in consts["domain_min"]: w.f32(v)
    for v in consts["domain_max"]: w.f32(v)
    w.u64(len(objects))
    for name in sorted(objects, key=lambda s: s.encode()):
        o = objects[name]
        w.string(name)
        if o[0] == "particles":
            w.u32(0); w.u64(o[1])
        else:
            w.u32(1); w.u64(o[1]); w.u64(o[2])


def encode_frame(w: Writer, frame: dict):
    """frame: {"gravity": (3,), "particles": {name: {"flags": .., attr: array | None}}, "colliders": {name: {...}}}"""
    for v in frame["gravity"]: w.f32(v)
    ps = frame.get("particles", {})
    w.u64(len(ps))
    for name in sorted(ps, key=lambda s: s.encode()):
        d = ps[name]
        w.string(name)
        w.vec(d["flags"], np.uint32)
        for attr, dtype, tail in PARTICLE_ATTRS:
            w.opt_vec(d.get(attr), dtype, tail)
    cs = frame.get("colliders", {})
    w.u64(len(cs))
    for name in sorted(cs, key=lambda s: s.encode()):
        d = cs[name]
        w.string(name)
        w.vec(d["vertex_positions"], np.float32, (3,))
        w.vec(d["triangle_indices"], np.uint32, (3,))
        w.vec(d["triangle_frictions"], np.float32)
        w.vec(d["triangle_dampings"], np.float32)


def encode_input_file(consts: dict, objects: Dict[str, tuple], frames: List[dict], version: str = VERSION) -> bytes:
    w = Writer()
    w.raw(INPUT_MAGIC)
    w.raw(version_bytes(version))
    encode_header(w, consts, objects)
    offsets = []
    for fr in frames:
        offsets.append(w.size)
        encode_frame(w, fr)
    index_offset = w.size
    w.vec(np.asarray(offsets, dtype=np.uint64), np.uint64)
    w.raw(struct.pack("<Q", index_offset))
    return w.bytes()


def object_ranges(objects: Dict[str, tuple]) -> Dict[str, tuple]:
    """InputRanges::new (header.rs:118-160): name -> (kind, start, count[, start2, count2])."""
    out, tp, tv, tt = {}, 0, 0, 0
    for name in sorted(objects, key=lambda s: s.encode()):
        o = objects[name]
        if o[0] == "particles":
            out[name] = ("particles", tp, o[1]); tp += o[1]
        else:
            out[name] = ("collider", tv, o[1], tt, o[2]); tv += o[1]; tt += o[2]
    return out


def initialize_io_state_ref(consts: dict, objects: Dict[str, tuple], frame0: dict) -> dict:
    """core/src/initialization.rs:84-277 in numpy (f32 arithmetic in the reference's order)."""
    ranges = object_ranges(objects)
    n = sum(o[1] for o in objects.values() if o[0] == "particles")
    f32 = np.float32
    inv_scale = f32(1.0) / f32(consts["simulation_scale"])
    out = {k: np.zeros(n, f32) for k in ("mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha", "viscosity_dynamic", "viscosity_bulk", "elastic_energies")}
    out["flags"] = np.zeros(n, np.uint32)
    out["collider_bits"] = np.zeros(n, np.uint32)
    for k in ("positions", "velocities", "initial_positions"):
        out[k] = np.zeros((n, 3), f32)
    out["velocity_gradients"] = np.zeros((n, 3, 3), f32)
    out["position_gradients"] = np.tile(np.eye(3, dtype=f32), (n, 1, 1))
    for name, d in frame0.get("particles", {}).items():
        _, first, count = ranges[name]
        sl = slice(first, first + count)
        fl = np.asarray(d["flags"], np.uint32)
        out["flags"][sl] = fl
        s = inv_scale * np.asarray(d["sizes"], f32)
        vol = s * s * s
        out["initial_volume"][sl] = vol
        out["mass"][sl] = vol * np.asarray(d["densities"], f32)
        visc = (fl & USE_VISCOSITY) != 0
        if visc.any():
            out["viscosity_dynamic"][sl][visc] = np.asarray(d["viscosities_dynamic"], f32)[visc]
            out["viscosity_bulk"][sl][visc] = np.asarray(d["viscosities_bulk"], f32)[visc]
        solid = (fl & IS_SOLID) != 0
        if solid.any():
            E = np.asarray(d["youngs_moduluses"], f32)
            nu = np.asarray(d["poissons_ratios"], f32)
            mu = E / f32(2.0) / (f32(1.0) + nu)
            lam = E * nu / (f32(1.0) + nu) / (f32(1.0) - f32(2.0) * nu)
            out["mu_or_bulk_modulus"][sl][solid] = mu[solid]
            out["lambda_or_exponent"][sl][solid] = lam[solid]
            sand = solid & ((fl & USE_SAND_ALPHA) != 0)
            if sand.any():
                out["sand_alpha"][sl][sand] = np.asarray(d["sand_alphas"], f32)[sand]
        if (~solid).any():
            out["mu_or_bulk_modulus"][sl][~solid] = np.asarray(d["bulk_moduluses"], f32)[~solid]
            out["lambda_or_exponent"][sl][~solid] = np.asarray(d["exponents"], np.uint32).astype(np.int32).astype(f32)[~solid]
        t = np.asarray(d["transforms"], f32).reshape(-1, 4, 4)
        out["position_gradients"][sl] = t[:, :3, :3]
        out["positions"][sl] = inv_scale * t[:, 3, :3]
        if d.get("initial_velocities") is not None:
            out["velocities"][sl] = d["initial_velocities"]
        if d.get("initial_positions") is not None:
            out["initial_positions"][sl] = d["initial_positions"]
    return out


def keyframe_ref(consts: dict, objects: Dict[str, tuple], frame: dict) -> dict:
    """xpu/src/frame_input.rs:66-135."""
    ranges = object_ranges(objects)
    n = sum(o[1] for o in objects.values() if o[0] == "particles")
    scale = np.float32(consts["simulation_scale"])
    flags = np.zeros(n, np.uint32)
    goals = np.zeros((n, 3), np.float32)
    for name, d in frame.get("particles", {}).items():
        _, first, count = ranges[name]
        flags[first:first + count] = d["flags"]
        if d.get("goal_positions") is not None:
            goals[first:first + count] = d["goal_positions"]
    cs = frame.get("colliders", {})
    names = sorted(cs, key=lambda s: s.encode())
    verts = np.concatenate([np.asarray(cs[k]["vertex_positions"], np.float32).reshape(-1, 3) for k in names]) if names else np.zeros((0, 3), np.float32)
    fr = np.concatenate([np.asarray(cs[k]["triangle_frictions"], np.float32) for k in names]) if names else np.zeros(0, np.float32)
    da = np.concatenate([np.asarray(cs[k]["triangle_dampings"], np.float32) for k in names]) if names else np.zeros(0, np.float32)
    return {"gravity": np.asarray(frame["gravity"], np.float32), "particle_flags": flags, "particle_goal_positions": goals / scale, "vertex_positions": verts / scale,
            "triangle_frictions": fr, "triangle_dampings": da}
