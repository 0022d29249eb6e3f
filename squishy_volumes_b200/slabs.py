"""Slab decomposition of the MPM domain across the GPUs of one box (the reference has
no multi-GPU path, its parity oracle is the single-domain result).

The domain is cut along x on grid-block planes (a block = 4 grid nodes, the tile size of the device
grid).  Rank r owns the particles whose base node `floor(x/h - 1/2)` lies in block columns
[lo_r, hi_r).  This module holds the host-side logic — planning the cut, splitting an `IoState`,
re-assembling results by original particle index — and `SlabState`, which drives one rank's
`libsvb200.so` handle (halo exchange and particle migration happen inside `svb_advance` over NCCL).
Pure numpy except for `SlabState`; `torch.distributed` (gloo or nccl) is only used to hand the NCCL
unique id around and to gather results.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import cstructs as cs
from .types import FatalError, FrameInput, IoState, Particles, RunParameters, SimulationError

BLOCK = 4                    # grid nodes per block edge (svb_device.cuh)
FAR = 1 << 15                # "infinity" in block units, inside the +-2^16 key range


def block_x(positions: np.ndarray, h: float) -> np.ndarray:
    """Block column of every particle: floor(floor(x/h - 1/2) / 4), in f32 like the device."""
    x = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)[:, 0]
    cell = np.floor(x / np.float32(h) - np.float32(0.5)).astype(np.int64)
    return np.floor_divide(cell, BLOCK)


def plan_slabs(positions: np.ndarray, h: float, n_ranks: int) -> List[Tuple[int, int]]:
    """Cut points on block planes that balance the particle count; the outer slabs extend to +-FAR.
    Every slab is at least one block wide; with fewer occupied columns than ranks the tail ranks
    get empty slabs beyond the material."""
    bx = block_x(positions, h)
    if bx.size == 0:
        lo0 = 0
        cuts = [lo0 + r for r in range(1, n_ranks)]
    else:
        cols, counts = np.unique(bx, return_counts=True)
        cum = np.cumsum(counts)
        total = int(cum[-1])
        cuts = []
        prev = int(cols[0])
        for r in range(1, n_ranks):
            target = total * r / n_ranks
            k = int(np.searchsorted(cum, target, side="left"))
            cut = int(cols[min(k, len(cols) - 1)]) + 1      # cut after column k
            cut = max(cut, prev + 1)
            cuts.append(cut)
            prev = cut
    bounds = [-FAR] + cuts + [FAR]
    return [(bounds[r], bounds[r + 1]) for r in range(n_ranks)]


def replan_slabs(plan: Sequence[Tuple[int, int]], first_col: int, hist: np.ndarray, max_move: int) -> List[Tuple[int, int]]:
    """New cuts that balance the particle count, given the GLOBAL column histogram of a running simulation
    (`hist[0]` = columns below `first_col`, `hist[1 + k]` = column `first_col + k`, `hist[-1]` = columns above) under the
    constraints of `svb_slab_rebalance`: the outer ends stay, every cut stays strictly inside the two old slabs it separates
    (rows only ever move to an adjacent rank), cuts stay inside the histogram window, and no cut shift hands over more than
    `max_move` particles at once (mailbox capacity).  Pure integer planning: every rank computes the same answer."""
    n = len(plan)
    if n < 2:
        return [tuple(p) for p in plan]
    hist = np.asarray(hist, dtype=np.int64)
    cols = hist.shape[0] - 2
    below = np.cumsum(hist)[:cols + 1]           # below[j] = particles in columns < first_col + j
    total = int(hist.sum())
    old = [int(p[0]) for p in plan[1:]]
    new: List[int] = []
    for k, cut in enumerate(old, start=1):
        if not (first_col <= cut <= first_col + cols):
            raise ValueError(f"cut {cut} lies outside the histogram window [{first_col}, {first_col + cols}]")
        target = total * k / n
        ideal = first_col + int(np.searchsorted(below, target, side="left"))
        lo_bound = (old[k - 2] if k >= 2 else int(plan[0][0])) + 1
        hi_bound = (old[k] if k < n - 1 else int(plan[-1][1])) - 1
        if new:
            lo_bound = max(lo_bound, new[-1] + 1)
        c = min(max(ideal, lo_bound, first_col), hi_bound, first_col + cols)
        step = 1 if c < cut else -1
        while c != cut and abs(int(below[c - first_col]) - int(below[cut - first_col])) > max_move:
            c += step                              # walk back towards the old cut until the hand-over fits
        new.append(c)
    bounds = [int(plan[0][0])] + new + [int(plan[-1][1])]
    return [(bounds[r], bounds[r + 1]) for r in range(n)]


def slab_of(positions: np.ndarray, h: float, plan: Sequence[Tuple[int, int]]) -> np.ndarray:
    bx = block_x(positions, h)
    cuts = np.array([hi for _, hi in plan[:-1]], dtype=np.int64)
    return np.searchsorted(cuts, bx, side="right").astype(np.int32)


def split_state(io_state: IoState, h: float, plan: Sequence[Tuple[int, int]], rank: int) -> Tuple[IoState, np.ndarray]:
    """-> (local IoState, original indices of its rows)."""
    owner = slab_of(io_state.particles.positions, h, plan)
    idx = np.nonzero(owner == rank)[0]
    return IoState(io_state.time, io_state.particles.select(idx)), idx.astype(np.int64)


def assemble(n: int, parts: Sequence[Tuple[np.ndarray, Particles]], template: Particles) -> Particles:
    """Scatter per-rank results (original index, rows) back into original particle order."""
    out = template.copy()
    seen = np.zeros(n, dtype=np.int32)
    for idx, p in parts:
        idx = np.asarray(idx, dtype=np.int64)
        np.add.at(seen, idx, 1)
        for f in dataclasses.fields(Particles):
            if f.name == "initial_positions":
                continue
            getattr(out, f.name)[idx] = getattr(p, f.name)
    if not np.all(seen == 1):
        raise ValueError(f"slab results do not partition the particles: {int((seen == 0).sum())} missing, {int((seen > 1).sum())} duplicated")
    return out


class SlabState:
    """One rank of a slab-decomposed run.  Same surface as `B200State`, plus `resident()`."""

    def __init__(self, inner, plan, rank: int, world: int, n_global: int):
        self.inner = inner
        self.plan = list(plan)
        self.rank = rank
        self.world = world
        self.n_global = n_global

    @classmethod
    def from_io_state(cls, io_state: IoState, frame_input: FrameInput, rank: int, world: int, device: int, unique_id: bytes,
                      plan: Optional[Sequence[Tuple[int, int]]] = None) -> "SlabState":
        from . import abi
        from .state import B200State
        h = frame_input.consts.scaled_grid_node_size()
        if plan is None:
            plan = plan_slabs(io_state.particles.positions, h, world)
        local, idx = split_state(io_state, h, plan, rank)
        inner = B200State.from_io_state(local, frame_input, device=device)
        inner.n = io_state.particles.n          # keyframe particle arrays stay in global original order
        L = abi.load()
        L.svb_set_option(inner._h, b"global_particles", float(io_state.particles.n))
        # the rows uploaded here are a subset of the global particle order: tell the device their global indices
        idx32 = np.ascontiguousarray(idx, dtype=np.uint32)
        if idx32.size:
            rc = L.svb_set_original_indices(inner._h, cs.uptr(idx32), idx32.size)
            if rc != 0:
                raise FatalError(rc, L.svb_last_error(inner._h).decode())
        uid = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        lo, hi = plan[rank]
        rc = L.svb_comm_init(inner._h, uid, rank, world, int(lo), int(hi), 0)
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(inner._h).decode())
        return cls(inner, plan, rank, world, io_state.particles.n)

    def upload(self, local: IoState, idx: np.ndarray) -> None:
        """Replace this rank's rows by `local` (rows `idx` of the global particle order, e.g. from `split_state`
        with this state's plan); communicator, mailboxes and allocations are kept.  Collective: every rank calls it."""
        from . import abi
        L = abi.load()
        n_global = self.n_global
        self.inner.upload(local)
        self.inner.n = n_global
        idx32 = np.ascontiguousarray(idx, dtype=np.uint32)
        if idx32.size:
            rc = L.svb_set_original_indices(self.inner._h, cs.uptr(idx32), idx32.size)
            if rc != 0:
                raise FatalError(rc, L.svb_last_error(self.inner._h).decode())

    def column_histogram(self, first_col: int, n_cols: int) -> np.ndarray:
        """Live particles per block column on this rank (n_cols + 2 entries, see include/svb200.h: svb_slab_histogram)."""
        from . import abi
        L = abi.load()
        out = np.zeros(n_cols + 2, dtype=np.uint64)
        rc = L.svb_slab_histogram(self.inner._h, int(first_col), int(n_cols), out.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self.inner._h).decode())
        return out

    def rebalance(self, dist, margin: int = 64, max_move: Optional[int] = None) -> bool:
        """Collective, between two `advance` calls: move the cuts so that the ranks hold equal particle counts again
        (SURVEY.md §8e).  `dist` is an initialised torch.distributed module (gloo or nccl) used to add up the column
        histograms.  Returns True when the plan changed."""
        from . import abi
        cuts = [p[0] for p in self.plan[1:]]
        if not cuts:
            return False
        first = min(cuts) - margin
        n_cols = max(cuts) - min(cuts) + 2 * margin
        local = self.column_histogram(first, n_cols).astype(np.int64)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, local)
        hist = np.sum(gathered, axis=0)
        if max_move is None:
            max_move = int(0.8 * (self.n_global // self.world // 4 + 65536))   # 80 % of a mailbox (svb_comm_init: n_max / 4 + 65536 rows)
        new_plan = replan_slabs(self.plan, first, hist, max_move)
        if [tuple(p) for p in new_plan] == [tuple(p) for p in self.plan]:
            return False
        L = abi.load()
        lo, hi = new_plan[self.rank]
        rc = L.svb_slab_rebalance(self.inner._h, int(lo), int(hi))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self.inner._h).decode())
        self.plan = list(new_plan)
        return True

    def advance(self, harness, frame_input: FrameInput, params: RunParameters) -> Optional[SimulationError]:
        return self.inner.advance(harness, frame_input, params)

    def resident(self, out: Optional[Particles] = None) -> Tuple[np.ndarray, Particles]:
        """(global original indices, rows) of the particles currently on this rank.  `out` (optional) are
        caller-owned result arrays (e.g. page-locked, reused across frames) with room for every resident row;
        the returned rows are then views into them."""
        from . import abi
        L = abi.load()
        cap = int(L.svb_particle_count(self.inner._h))
        if out is None:
            out = Particles.empty(max(cap, 1))
        elif out.n < cap:
            raise ValueError(f"resident(): out holds {out.n} rows, {cap} are resident")
        s = cs.particles_struct(out)
        orig = np.zeros(max(cap, 1), dtype=np.uint64)
        rc = L.svb_download_resident(self.inner._h, C.byref(s), orig.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc != 0:
            raise FatalError(rc, L.svb_last_error(self.inner._h).decode())
        m = int(s.n)
        rows = Particles(**{f.name: getattr(out, f.name)[:m] for f in dataclasses.fields(Particles)})
        return orig[:m].astype(np.int64), rows

    @property
    def time(self) -> float:
        return self.inner.time

    @property
    def substeps(self) -> int:
        return self.inner.substeps

    def close(self) -> None:
        self.inner.close()
