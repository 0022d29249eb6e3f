"""torchrun worker of the multi-GPU parity test: every rank runs its slab, rank 0 assembles the result and
compares it with the single-GPU run of the same scene (and the oracle).  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/slab_worker.py <scene> <substeps>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from squishy_volumes_b200 import abi, scenes, slabs  # noqa: E402
from squishy_volumes_b200.state import B200State  # noqa: E402
from squishy_volumes_b200.types import ParticleFlags, RunParameters  # noqa: E402
from tests import golden_scenes, parity  # noqa: E402


def build_scene(name):
    if name in ("jelly", "jelly_rebalance"):
        sc = scenes.jelly_collision(side=16)
        sc.io_state.particles.velocities[:, 0] *= 6.0          # fast approach: many particles cross the cut
        return sc
    if name == "jelly_shear":
        sc = scenes.jelly_collision(side=16)
        p = sc.io_state.particles
        p.velocities[:, 0] = np.where(p.positions[:, 1] > 0, 3.0, -3.0)   # both directions across every cut
        return sc
    if name == "split_layers":
        sc, _, _ = golden_scenes.GOLDEN["split_layers"]()
        return sc
    if name == "sand":
        sc, _, _ = golden_scenes.GOLDEN["sand_torus"]()
        return sc
    if name == "jelly_adaptive":
        # adaptive time steps on slab ranks: max_time_step 4x the fixed step, so the four limits decide; the global minima / maxima
        # travel through the mailbox headers (k_dt_*: dt_exchange) and every rank must take the very same steps
        sc = scenes.jelly_collision(side=16)
        sc.io_state.particles.velocities[:, 0] *= 6.0
        sc.time_step = 4e-3
        return sc
    if name == "energy_error":
        # a FAILED particle on one rank stops every rank; the state of the failing substep comes back with the simulation-level error
        sc = scenes.jelly_collision(side=16)
        sc.io_state.particles.velocities[:, 0] *= 6.0
        sc.io_state.particles.position_gradients[7] = -np.eye(3, dtype=np.float32)
        return sc
    raise SystemExit(name)


def main():
    name, steps = sys.argv[1], int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    L = abi.load()
    sc = build_scene(name)
    sc.frame_input.consts.frames_per_second = 1
    box = [None]
    if rank == 0:
        import ctypes as C
        buf = (C.c_uint8 * 128)()
        assert L.svb_comm_unique_id(buf) == 0
        box[0] = bytes(buf)
    dist.broadcast_object_list(box, src=0)
    plan = None
    if name == "jelly_rebalance":
        # start from cuts that are off balance by a block column each, run half the substeps, rebalance, run the rest
        h0 = sc.frame_input.consts.scaled_grid_node_size()
        good = slabs.plan_slabs(sc.io_state.particles.positions, h0, world)
        cuts = [p[0] + (1 if k % 2 == 0 else -1) for k, p in enumerate(good[1:])]
        for k in range(1, len(cuts)):
            cuts[k] = max(cuts[k], cuts[k - 1] + 1)
        bounds = [good[0][0]] + cuts + [good[-1][1]]
        plan = [(bounds[r], bounds[r + 1]) for r in range(world)]
    st = slabs.SlabState.from_io_state(sc.io_state, sc.frame_input, rank, world, local, box[0], plan=plan)
    params = RunParameters(target_time=(steps - 0.5) * sc.time_step, max_time_step=sc.time_step)
    adaptive = name == "jelly_adaptive"
    if adaptive:
        params = RunParameters(target_time=steps * 1e-3, max_time_step=sc.time_step, adaptive_time_steps=True)
    if name == "energy_error":
        err = st.advance(None, sc.frame_input, params)
        idx, rows = st.resident()
        gathered = [None] * world
        dist.all_gather_object(gathered, (idx, rows, st.substeps, st.time, None if err is None else err.status))
        if rank == 0:
            single = B200State.from_io_state(sc.io_state, sc.frame_input, device=local)
            ref, err1 = single.produce_next_state(None, sc.frame_input, params)
            assert err1 is not None and err1.status & 8
            assert all(g[4] is not None and g[4] & 8 for g in gathered), [g[4] for g in gathered]          # every rank reports it
            assert all(g[2] == single.substeps and g[3] == single.time for g in gathered), ([(g[2], g[3]) for g in gathered], single.substeps, single.time)
            got = slabs.assemble(sc.n, [(g[0], g[1]) for g in gathered], sc.io_state.particles)
            assert np.array_equal(got.flags, ref.particles.flags) and got.flags[7] & ParticleFlags.FAILED
            print(f"[{name}] world={world}: every rank stopped after the failing substep ({single.substeps} counted), flags equal the single-GPU run; within tolerance", flush=True)
        st.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    if name == "jelly_rebalance":
        half = RunParameters(target_time=(steps // 2 - 0.5) * sc.time_step, max_time_step=sc.time_step)
        err = st.advance(None, sc.frame_input, half)
        assert err is None and st.substeps == steps // 2
        before = [tuple(p) for p in st.plan]
        n_before = len(st.resident()[0])
        changed = st.rebalance(dist, margin=8)
        counts = [None] * world
        dist.all_gather_object(counts, (n_before, len(st.resident()[0])))
        if rank == 0:
            print(f"[{name}] plan {before} -> {[tuple(p) for p in st.plan]} changed={changed}; resident before/after {counts}", flush=True)
            assert changed and sum(c[1] for c in counts) == sc.n
            assert max(c[1] for c in counts) - min(c[1] for c in counts) <= max(c[0] for c in counts) - min(c[0] for c in counts)
    err = st.advance(None, sc.frame_input, params)
    idx, rows = st.resident()
    gathered = [None] * world
    dist.all_gather_object(gathered, (idx, rows, st.substeps, None if err is None else err.status, st.time, st.inner.allowed_time_step))
    ok = True
    if rank == 0:
        print('resident per rank', [len(g[0]) for g in gathered], 'sum', sum(len(g[0]) for g in gathered), 'of', sc.n, 'unique', len(np.unique(np.concatenate([g[0] for g in gathered]))), flush=True)
        got = slabs.assemble(sc.n, [(g[0], g[1]) for g in gathered], sc.io_state.particles)
        single = B200State.from_io_state(sc.io_state, sc.frame_input, device=local)
        ref, _ = single.produce_next_state(None, sc.frame_input, params)
        if adaptive:
            # every rank took the very same steps (the clocks agree bit for bit); against the single-GPU run the limits differ in the
            # last bits (float atomics order the grid sums differently), so its clock is only close
            assert all(g[2] == gathered[0][2] and g[4] == gathered[0][4] and g[5] == gathered[0][5] for g in gathered), [(g[2], g[4], g[5]) for g in gathered]
            assert gathered[0][2] == single.substeps and single.substeps < steps, (gathered[0][2], single.substeps)
            assert abs(gathered[0][4] - single.time) <= 1e-5 * single.time and abs(gathered[0][5] - single.allowed_time_step) <= 1e-4 * single.allowed_time_step
        else:
            assert all(g[2] == steps for g in gathered), [g[2] for g in gathered]
        from squishy_volumes_b200.types import IoState
        h = sc.frame_input.consts.scaled_grid_node_size()
        # slab ownership after the run: every particle sits on the rank that owns its block column
        owner = slabs.slab_of(got.positions, h, st.plan)
        live = (got.flags & ParticleFlags.TOMBSTONED) == 0
        for r, g in enumerate(gathered):
            mine = np.zeros(sc.n, bool)
            mine[g[0]] = True
            assert np.all(owner[mine & live] == r), f"rank {r} holds particles of another slab"
        rep = parity.compare_states(IoState(0.0, got), ref, rtol=parity.RTOL_RUN, h=h)
        moved = int(np.count_nonzero(slabs.slab_of(sc.io_state.particles.positions, h, st.plan) != owner))
        print(f"[{name}] world={world} n={sc.n} substeps={steps} migrated={moved} per-rank={[len(g[0]) for g in gathered]} max err "
              + ", ".join(f"{k}={v[0]:.2e}/{v[1]:.2e}" for k, v in rep.items() if isinstance(v, tuple)), flush=True)
        import oracle.oracle as orc
        o = orc.OracleState.from_io_state(sc.io_state, sc.frame_input)
        oref, _ = o.produce_next_state(None, sc.frame_input, params)
        parity.compare_states(IoState(0.0, got), oref, rtol=parity.RTOL_RUN, h=h)
        print(f"[{name}] slab result == single-GPU result == oracle within tolerance; integer fields exact", flush=True)
    if name == "jelly":
        # the same handles, communicator and mailboxes take a new state (svb_upload): the rerun reproduces the first run
        h = sc.frame_input.consts.scaled_grid_node_size()
        local, lidx = slabs.split_state(sc.io_state, h, st.plan, rank)
        st.upload(local, lidx)
        assert st.substeps == 0
        err2 = st.advance(None, sc.frame_input, params)
        idx2, rows2 = st.resident()
        again = [None] * world
        dist.all_gather_object(again, (idx2, rows2, st.substeps, None if err2 is None else err2.status))
        if rank == 0:
            assert all(g[2] == steps and g[3] is None for g in again), [(g[2], g[3]) for g in again]
            from squishy_volumes_b200.types import IoState as _Io
            first_run = slabs.assemble(sc.n, [(g[0], g[1]) for g in gathered], sc.io_state.particles)
            second_run = slabs.assemble(sc.n, [(g[0], g[1]) for g in again], sc.io_state.particles)
            assert np.array_equal(first_run.flags, second_run.flags)
            parity.compare_states(_Io(0.0, second_run), _Io(0.0, first_run), rtol=parity.RTOL_RUN, h=h)
            print(f"[{name}] upload + rerun on the same handles reproduces the first run", flush=True)
    st.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        print(f"[rank {os.environ.get('RANK')}] FAILED:\n" + traceback.format_exc(), flush=True)   # on stdout: torchrun's own summary drowns stderr
        os._exit(1)
