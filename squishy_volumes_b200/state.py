"""`B200State` — host-side mirror of the reference's back-end state for the MPM substep path.

Same surface as `CpuState` / `GpuState` (rust/crates/cpu/src/cpu_state.rs:26,71,146-156,
rust/crates/gpu/src/gpu_state.rs:38-45,331-345):

    state = B200State.from_io_state(io_state, frame_input, device=0)
    io_state, sim_error = state.produce_next_state(harness, frame_input, RunParameters(...))

* outer errors (cancelled, wrong frame loaded, zero time step, CUDA failure) raise `FatalError`;
* the inner, simulation-level error (`EnergyError`) is RETURNED together with a valid state, which
  the caller still stores (core/src/compute_thread.rs:165-169);
* the returned `IoState` is in original particle order.

Everything numerical happens in `lib/libsvb200.so` (CUDA, sm_100a) through the C ABI of
include/svb200.h; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import abi
from . import cstructs as cs
from .types import FatalError, FrameInput, GridNodes, Harness, IoState, Particles, RunParameters, SimulationError

_PROGRESS = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t)


def available_gpus():
    """`Context::available_gpus` (core/src/api_impl/context.rs:10-12)."""
    L = abi.load()
    buf = C.create_string_buffer(4096)
    n = L.svb_available_devices(buf, len(buf))
    if n < 0:
        return []
    return [line for line in buf.value.decode().splitlines() if line]


class B200State:
    def __init__(self, handle, n: int, frame_input: FrameInput):
        self._h = handle
        self.n = n
        self._nv = frame_input.num_vertices()
        self._nt = frame_input.num_triangles()
        self._loaded: Optional[Tuple[int, int]] = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_io_state(cls, io_state: IoState, frame_input: FrameInput, device: int = 0) -> "B200State":
        L = abi.load()
        p = io_state.particles.normalized()
        ps = cs.particles_struct(p)
        consts = cs.consts_struct(frame_input.consts)
        h = C.c_void_p()
        rc = L.svb_create(C.byref(consts), C.byref(ps), C.c_double(io_state.time), int(device), C.byref(h))
        if rc != 0:
            msg = L.svb_last_error(h).decode() if h else "svb_create failed (no CUDA device?)"
            if h:
                L.svb_destroy(h)
            raise FatalError(rc, msg)
        self = cls(h, p.n, frame_input)
        nv, nt, flat = cs.topology_arrays(frame_input)
        rc = L.svb_set_topology(h, len(frame_input.colliders), cs.uptr(nv), cs.uptr(nt), cs.uptr(flat))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(h).decode())
        return self

    @classmethod
    def from_io_state_multi(cls, io_state: IoState, frame_input: FrameInput, devices) -> "B200State":
        """One state over several GPUs of the box, driven from this single thread like the reference's compute thread drives a
        back end (core/src/compute_thread.rs:100-163): `svb_create_multi` cuts the particles into slabs along x inside the
        library, one slab rank per device; every other call of this class works unchanged (the introspection calls —
        `grid`, `binning`, `active_blocks`, `snapshot` — are single-device only)."""
        L = abi.load()
        p = io_state.particles.normalized()
        ps = cs.particles_struct(p)
        consts = cs.consts_struct(frame_input.consts)
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = L.svb_create_multi(C.byref(consts), C.byref(ps), C.c_double(io_state.time), devs, len(devices), C.byref(h))
        if rc != 0:
            msg = L.svb_last_error(h).decode() if h else "svb_create_multi failed (no CUDA device?)"
            if h:
                L.svb_destroy(h)
            raise FatalError(rc, msg)
        self = cls(h, p.n, frame_input)
        nv, nt, flat = cs.topology_arrays(frame_input)
        rc = L.svb_set_topology(h, len(frame_input.colliders), cs.uptr(nv), cs.uptr(nt), cs.uptr(flat))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(h).decode())
        return self

    def upload(self, io_state: IoState) -> None:
        """`from_io_state` into this (session-lived) handle: H2D of a new state, clock and step history restart;
        constants, colliders and device allocations are kept (include/svb200.h: svb_upload)."""
        L = abi.load()
        p = io_state.particles.normalized()
        ps = cs.particles_struct(p)
        rc = L.svb_upload(self._h, C.byref(ps), C.c_double(io_state.time))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self._h).decode())
        self.n = p.n
        self._loaded = None   # keyframe particle arrays are sized by the particle count

    def close(self) -> None:
        if getattr(self, "_h", None):
            abi.load().svb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ stepping
    def _sync_keyframes(self, fi: FrameInput) -> None:
        key = (fi.frame, fi.a_index())
        if self._loaded == key:
            return
        L = abi.load()
        a, keep_a = cs.keyframe_struct(fi.a(), self.n, self._nv, self._nt)
        b = fi.b()
        if b is not None:
            bs, keep_b = cs.keyframe_struct(b, self.n, self._nv, self._nt)
            rc = L.svb_set_keyframes(self._h, fi.frame, C.byref(a), C.byref(bs))
        else:
            rc = L.svb_set_keyframes(self._h, fi.frame, C.byref(a), None)
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self._h).decode())
        self._loaded = key

    def advance(self, harness: Optional[Harness], frame_input: FrameInput, params: RunParameters) -> Optional[SimulationError]:
        """The substep loop only (state stays on the device)."""
        L = abi.load()
        self._sync_keyframes(frame_input)
        L.svb_set_option(self._h, b"store_grid", 1.0 if params.store_grid else 0.0)
        cancel = harness.cancel_pointer() if harness is not None else None
        cb = None
        if harness is not None and harness.progress is not None:
            prog = harness.progress
            cb = _PROGRESS(lambda _user, ms: prog(int(ms)))
        rc = L.svb_advance(self._h, C.c_double(params.target_time), C.c_float(params.max_time_step),
                           int(params.adaptive_time_steps), cancel, C.cast(cb, C.c_void_p) if cb else None, None)
        if rc < 0:
            raise FatalError(rc, L.svb_last_error(self._h).decode())
        return SimulationError(rc, L.svb_last_error(self._h).decode()) if rc > 0 else None

    def produce_next_state(self, harness: Optional[Harness], frame_input: FrameInput, params: RunParameters, out: Optional[Particles] = None):
        """-> (IoState, SimulationError | None); raises FatalError for the reference's outer Err.
        `out` (optional) are caller-owned result arrays, e.g. page-locked ones, reused across frames."""
        err = self.advance(harness, frame_input, params)
        return self.to_io_state(params.store_grid, out=out), err

    # ------------------------------------------------------------------ readback
    def to_io_state(self, store_grid: bool = False, out: Optional[Particles] = None) -> IoState:
        L = abi.load()
        if out is None:
            out = Particles.empty(self.n)
        s = cs.particles_struct(out)
        rc = L.svb_download(self._h, C.byref(s))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self._h).decode())
        return IoState(time=self.time, particles=out, grid_nodes=self.grid() if store_grid else None)

    def grid(self) -> GridNodes:
        L = abi.load()
        n = int(L.svb_grid_count(self._h))
        if n < 0:
            raise FatalError(n, L.svb_last_error(self._h).decode())
        g, s = cs.alloc_grid(n, with_counts=True)
        rc = L.svb_download_grid(self._h, C.byref(s))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self._h).decode())
        return g

    def binning(self):
        """(sort_map, cells): current order -> original index, and base node (i,j,k) per current row."""
        sm = np.zeros(self.n, dtype=np.uint32)
        cells = np.zeros((self.n, 3), dtype=np.int32)
        rc = abi.load().svb_binning(self._h, cs.uptr(sm), cells.ctypes.data_as(cs.c_i32p))
        if rc != 0:
            raise FatalError(rc, abi.load().svb_last_error(self._h).decode())
        return sm, cells

    def active_blocks(self):
        """(block_ids (a,3) int32, collider_bits (a,)) of the last substep's active grid tiles."""
        L = abi.load()
        n = int(L.svb_active_block_count(self._h))
        ids = np.zeros((max(n, 0), 3), dtype=np.int32)
        bits = np.zeros(max(n, 0), dtype=np.uint32)
        if n > 0:
            rc = L.svb_active_blocks(self._h, ids.ctypes.data_as(cs.c_i32p), cs.uptr(bits))
            if rc != 0:
                raise FatalError(rc, L.svb_last_error(self._h).decode())
        return ids, bits

    def enable_stage_timing(self, on: bool = True) -> None:
        abi.load().svb_enable_stage_timing(self._h, int(on))

    def stage_times(self) -> Dict[str, float]:
        L = abi.load()
        names = (C.c_char_p * 32)()
        ms = np.zeros(32, dtype=np.float32)
        k = L.svb_stage_times(self._h, names, cs.fptr(ms), 32)
        return {names[i].decode(): float(ms[i]) for i in range(k)}

    EXCHANGE_WAITS = ("halo_send_for_p2g_boundary", "halo_recv_for_neighbours", "migrate_send_for_g2p_boundary", "migrate_recv_for_neighbours",
                      "migrate_recv_for_error_words", "time_step_reductions", "particle_tiles", "grid_tiles")

    def exchange_waits(self, reset: bool = False) -> Dict[str, float]:
        """Slab ranks: milliseconds this rank's exchange kernels have spent waiting since the last reset (device-side clocks)."""
        L = abi.load()
        ms = (C.c_double * 8)()
        k = L.svb_exchange_waits(self._h, ms, 8, int(reset))
        if k < 0:
            raise FatalError(k, L.svb_last_error(self._h).decode())
        return {self.EXCHANGE_WAITS[i]: float(ms[i]) for i in range(k)}

    def snapshot(self) -> None:
        rc = abi.load().svb_snapshot(self._h)
        if rc != 0:
            raise FatalError(rc, abi.load().svb_last_error(self._h).decode())

    def restore(self) -> None:
        rc = abi.load().svb_restore(self._h)
        if rc != 0:
            raise FatalError(rc, abi.load().svb_last_error(self._h).decode())

    @property
    def time(self) -> float:
        return float(abi.load().svb_time(self._h))

    @property
    def substeps(self) -> int:
        return int(abi.load().svb_substeps(self._h))

    @property
    def allowed_time_step(self) -> float:
        return float(abi.load().svb_allowed_time_step(self._h))

    @property
    def last_advance_ms(self) -> float:
        return float(abi.load().svb_last_advance_ms(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(abi.load().svb_kernel_launches(self._h))
