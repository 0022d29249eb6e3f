"""ORACLE — test infrastructure only.  PARITY UNPINNED end-to-end (see oracle/README.md).

ctypes binding of the C++ restatement of the reference's CPU back end
(/root/reference/rust/crates/cpu/src/**).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the product
package (``squishy_volumes_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

from squishy_volumes_b200 import cstructs as cs
from squishy_volumes_b200.types import (FatalError, FrameInput, GridNodes, Harness, IoState, Particles,
                                        RunParameters, SimulationError)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/_build/liboracle.so with the committed Makefile (g++ -O3 -fopenmp)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("svo_capi.cpp", "svo_state.cpp", "svo_math.h", "svo_mesh.h", "svo_state.h", "Makefile")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.svo_create.restype = vp
        L.svo_create.argtypes = [C.POINTER(cs.SvbConsts), C.POINTER(cs.SvbParticles), C.c_double]
        L.svo_destroy.argtypes = [vp]
        L.svo_set_deterministic.argtypes = [vp, C.c_int]
        L.svo_set_threads.argtypes = [C.c_int]
        L.svo_max_threads.restype = C.c_int
        L.svo_set_topology.argtypes = [vp, C.c_uint32, cs.c_u32p, cs.c_u32p, cs.c_u32p]
        L.svo_set_keyframes.argtypes = [vp, C.c_uint64, C.POINTER(cs.SvbKeyframe), C.POINTER(cs.SvbKeyframe)]
        L.svo_advance.argtypes = [vp, C.c_double, C.c_float, C.c_int, vp]
        L.svo_time.restype = C.c_double
        L.svo_time.argtypes = [vp]
        L.svo_substeps.restype = C.c_uint64
        L.svo_substeps.argtypes = [vp]
        L.svo_allowed_time_step.restype = C.c_float
        L.svo_allowed_time_step.argtypes = [vp]
        L.svo_download.argtypes = [vp, C.POINTER(cs.SvbParticles)]
        L.svo_sort_map.argtypes = [vp, cs.c_u32p]
        L.svo_grid_count.restype = C.c_uint64
        L.svo_grid_count.argtypes = [vp]
        L.svo_download_grid.argtypes = [vp, C.POINTER(cs.SvbGrid)]
        for name in ("svo_kernel_linear", "svo_kernel_quadratic", "svo_kernel_cubic"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        L.svo_shift_quadratic.argtypes = [C.c_uint64, cs.c_f32p, C.c_float, cs.c_i32p]
        L.svo_bits_get.argtypes = [C.c_uint32, C.c_uint32]
        L.svo_bits_set.restype = C.c_uint32
        L.svo_bits_set.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.svo_bits_compatible.argtypes = [C.c_uint32, C.c_uint32]
        dp = C.POINTER(C.c_double)
        L.svo_lame_mu.restype = L.svo_lame_lambda.restype = C.c_double
        L.svo_lame_mu.argtypes = L.svo_lame_lambda.argtypes = [C.c_double, C.c_double]
        L.svo_energy_neo_hookean.restype = C.c_double
        L.svo_energy_neo_hookean.argtypes = [C.c_double, C.c_double, dp]
        L.svo_stress_neo_hookean.argtypes = [C.c_double, C.c_double, dp, dp]
        L.svo_stress_neo_hookean_svd_diag.argtypes = [C.c_double, C.c_double, dp, dp]
        L.svo_second_neo_hookean_svd_diag.argtypes = [C.c_double, C.c_double, dp, dp]
        L.svo_energy_inviscid.restype = C.c_double
        L.svo_energy_inviscid.argtypes = [C.c_double, C.c_int, dp]
        L.svo_stress_inviscid.argtypes = [C.c_double, C.c_int, dp, dp]
        L.svo_stress_inviscid_svd_diag.argtypes = [C.c_double, C.c_int, dp, dp]
        L.svo_second_inviscid_svd_diag.argtypes = [C.c_double, C.c_int, dp, dp]
        L.svo_viscosity_stress.argtypes = [C.c_double, C.c_double, dp, dp]
        L.svo_det3.restype = C.c_double
        L.svo_det3.argtypes = [dp]
        L.svo_svd3.argtypes = [dp, dp, dp, dp]
        L.svo_distance_to_triangle.restype = C.c_float
        L.svo_distance_to_triangle.argtypes = [cs.c_f32p] * 5
        L.svo_bvh_build.restype = vp
        L.svo_bvh_build.argtypes = [C.c_uint64, cs.c_f32p, C.c_float, C.c_float, C.c_uint32]
        L.svo_bvh_destroy.argtypes = [vp]
        L.svo_bvh_level.restype = C.c_uint32
        L.svo_bvh_level.argtypes = [vp]
        L.svo_bvh_num_nodes.restype = C.c_uint64
        L.svo_bvh_num_nodes.argtypes = [vp]
        L.svo_bvh_query.restype = C.c_uint64
        L.svo_bvh_query.argtypes = [vp, cs.c_i32p, cs.c_u32p, C.c_uint64]
        L.svo_handle_bvh_query.restype = C.c_uint64
        L.svo_handle_bvh_query.argtypes = [vp, cs.c_i32p, cs.c_u32p, C.c_uint64]
        L.svo_topology_counts.restype = C.c_uint64
        L.svo_topology_counts.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.svo_topology_get.argtypes = [vp, cs.c_u32p, cs.c_u32p, cs.c_u32p, cs.c_u32p]
        L.svo_interpolated_mesh.argtypes = [vp, cs.c_f32p, cs.c_f32p, cs.c_f32p]
        _lib = L
    return _lib


_STATUS_TEXT = {
    8: "Failed to compute the elastic energy of a particle (EnergyError::PositionGradientNonPositive)",
    -1: "The computation was canceled",
    -2: "The time step ended up being 0",
    -3: "Something went wrong accessing frame input (WrongFrameLoaded)",
    -4: "At this point, interpolated input should be ready",
}


class OracleState:
    """Mirror of ``CpuState`` (cpu/src/cpu_state.rs:15-197) over the C++ restatement."""

    def __init__(self, handle, n: int, frame_input: FrameInput):
        self._h = handle
        self.n = n
        self._nv = frame_input.num_vertices()
        self._nt = frame_input.num_triangles()
        self._loaded: Optional[Tuple[int, int]] = None

    @classmethod
    def from_io_state(cls, io_state: IoState, frame_input: FrameInput) -> "OracleState":
        L = lib()
        p = io_state.particles.normalized()
        ps = cs.particles_struct(p)
        consts = cs.consts_struct(frame_input.consts)
        h = L.svo_create(C.byref(consts), C.byref(ps), C.c_double(io_state.time))
        self = cls(h, p.n, frame_input)
        self._h_size = frame_input.consts.scaled_grid_node_size()
        self._params = p  # material parameters never change on the path; echoed back by to_io_state
        nv, nt, flat = cs.topology_arrays(frame_input)
        rc = L.svo_set_topology(h, len(frame_input.colliders), cs.uptr(nv), cs.uptr(nt), cs.uptr(flat))
        if rc != 0:
            raise FatalError(rc, "Something is wrong with the mesh inputs")
        return self

    def __del__(self):
        if getattr(self, "_h", None):
            lib().svo_destroy(self._h)
            self._h = None

    def set_deterministic(self, on: bool) -> None:
        lib().svo_set_deterministic(self._h, int(on))

    def _sync_keyframes(self, fi: FrameInput) -> None:
        key = (fi.frame, fi.a_index())
        if self._loaded == key:
            return
        a, keep_a = cs.keyframe_struct(fi.a(), self.n, self._nv, self._nt)
        b = fi.b()
        if b is not None:
            bs, keep_b = cs.keyframe_struct(b, self.n, self._nv, self._nt)
            lib().svo_set_keyframes(self._h, fi.frame, C.byref(a), C.byref(bs))
        else:
            lib().svo_set_keyframes(self._h, fi.frame, C.byref(a), None)
        self._loaded = key

    def produce_next_state(self, harness: Optional[Harness], frame_input: FrameInput, params: RunParameters):
        """-> (IoState, SimulationError | None); raises FatalError (outer Err of the reference)."""
        self._sync_keyframes(frame_input)
        cancel = harness.cancel_pointer() if harness is not None else None
        rc = lib().svo_advance(self._h, C.c_double(params.target_time), C.c_float(params.max_time_step),
                               int(params.adaptive_time_steps), cancel)
        if rc < 0:
            raise FatalError(rc, _STATUS_TEXT.get(rc, f"oracle status {rc}"))
        state = self.to_io_state(params.store_grid)
        return state, (SimulationError(rc, _STATUS_TEXT.get(rc, f"oracle status {rc}")) if rc > 0 else None)

    def to_io_state(self, store_grid: bool = False) -> IoState:
        out = Particles.empty(self.n)
        for name in ("mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha",
                     "viscosity_dynamic", "viscosity_bulk"):
            setattr(out, name, getattr(self._params, name).copy())
        s = cs.particles_struct(out)
        lib().svo_download(self._h, C.byref(s))
        grid = self.grid() if store_grid else None
        return IoState(time=self.time, particles=out, grid_nodes=grid)

    def grid(self) -> GridNodes:
        n = int(lib().svo_grid_count(self._h))
        g, s = cs.alloc_grid(n, with_counts=True)
        lib().svo_download_grid(self._h, C.byref(s))
        return g

    def binning(self):
        """(sort_map, cells) like B200State.binning: current order -> original index, base node per row
        (cells recomputed from the CURRENT positions with the restatement of kernels.rs:46-49)."""
        sm = self.sort_map()
        pos = self.to_io_state().particles.positions  # original order
        cells_orig = shift_quadratic(pos, float(self._h_size))
        return sm, cells_orig[sm]

    def sort_map(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint32)
        lib().svo_sort_map(self._h, cs.uptr(out))
        return out

    def interpolated_mesh(self):
        vp = np.zeros((self._nv, 3), np.float32)
        vn = np.zeros((self._nv, 3), np.float32)
        tn = np.zeros((self._nt, 3), np.float32)
        rc = lib().svo_interpolated_mesh(self._h, cs.fptr(vp), cs.fptr(vn), cs.fptr(tn))
        return (vp, vn, tn) if rc == 0 else None

    @property
    def time(self) -> float:
        return float(lib().svo_time(self._h))

    @property
    def substeps(self) -> int:
        return int(lib().svo_substeps(self._h))

    @property
    def allowed_time_step(self) -> float:
        return float(lib().svo_allowed_time_step(self._h))


def shift_quadratic(positions: np.ndarray, h: float) -> np.ndarray:
    p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(p.shape, dtype=np.int32)
    lib().svo_shift_quadratic(p.shape[0], cs.fptr(p), C.c_float(h), out.ctypes.data_as(cs.c_i32p))
    return out
