"""Host-side logic of the multi-GPU path on CPU: slab planning, splitting and re-assembly, exercised
with a 2-process gloo group (no GPU, no compute: the device exchange itself is covered by
tests/test_gpu_slabs.py on the B200 box)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from squishy_volumes_b200 import scenes, slabs


def test_plan_is_contiguous_and_balanced():
    sc = scenes.jelly_collision(side=16)
    h = sc.frame_input.consts.scaled_grid_node_size()
    for world in (1, 2, 3, 4, 8):
        plan = slabs.plan_slabs(sc.io_state.particles.positions, h, world)
        assert len(plan) == world and plan[0][0] == -slabs.FAR and plan[-1][1] == slabs.FAR
        assert all(plan[r][1] == plan[r + 1][0] for r in range(world - 1))
        assert all(lo < hi for lo, hi in plan)
        owner = slabs.slab_of(sc.io_state.particles.positions, h, plan)
        counts = np.bincount(owner, minlength=world)
        assert counts.sum() == sc.n
        if world <= 4:
            assert counts.max() <= 1.6 * sc.n / world + 4 * 512   # cut points sit on block planes
        bx = slabs.block_x(sc.io_state.particles.positions, h)
        for r, (lo, hi) in enumerate(plan):
            assert np.all((bx[owner == r] >= lo) & (bx[owner == r] < hi))


def test_block_x_matches_device_formula():
    import oracle.oracle as orc
    rng = np.random.Generator(np.random.Philox(4))
    pos = (rng.random((5000, 3), dtype=np.float32) * 8 - 4).astype(np.float32)
    h = 0.037
    cells = orc.shift_quadratic(pos, h)
    assert np.array_equal(slabs.block_x(pos, h), cells[:, 0] >> 2)


def test_split_and_assemble_roundtrip():
    sc = scenes.jelly_collision(side=16)
    h = sc.frame_input.consts.scaled_grid_node_size()
    plan = slabs.plan_slabs(sc.io_state.particles.positions, h, 3)
    parts = []
    for r in range(3):
        local, idx = slabs.split_state(sc.io_state, h, plan, r)
        assert local.particles.n == idx.size
        parts.append((idx, local.particles))
    out = slabs.assemble(sc.n, parts, sc.io_state.particles)
    for f in ("positions", "velocities", "flags", "mass"):
        assert np.array_equal(getattr(out, f), getattr(sc.io_state.particles, f))
    with pytest.raises(ValueError):
        slabs.assemble(sc.n, parts[1:], sc.io_state.particles)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = scenes.jelly_collision(side=12)
        h = sc.frame_input.consts.scaled_grid_node_size()
        plan = slabs.plan_slabs(sc.io_state.particles.positions, h, world)
        local, idx = slabs.split_state(sc.io_state, h, plan, rank)
        # what rank 0 does with the per-rank results: gather (index, rows) and assemble
        gathered = [None] * world
        dist.all_gather_object(gathered, (idx, local.particles))
        # what the launcher does with the NCCL id: rank 0 creates 128 bytes, everyone receives them
        box = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ok = box[0] == bytes(range(128))
        if rank == 0:
            out = slabs.assemble(sc.n, gathered, sc.io_state.particles)
            ok = ok and np.array_equal(out.positions, sc.io_state.particles.positions) and sum(g[0].size for g in gathered) == sc.n
        q.put((rank, bool(ok), int(idx.size)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_roundtrip():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    assert sum(n for _, _, n in results) == scenes.jelly_collision(side=12).n


def _hist(bx, first, n_cols):
    k = bx - first
    bins = np.where(k < 0, 0, np.where(k >= n_cols, n_cols + 1, k + 1))
    return np.bincount(bins, minlength=n_cols + 2)


def test_replan_moves_cuts_towards_balance_within_its_constraints():
    """replan_slabs: the planning half of svb_slab_rebalance (rebalanced by particle count)."""
    rng = np.random.Generator(np.random.Philox(9))
    bx = np.concatenate([rng.integers(0, 40, 30000), rng.integers(40, 48, 30000)])      # dense material on the right
    plan = [(-slabs.FAR, 10), (10, 20), (20, 30), (30, slabs.FAR)]
    first, n_cols = -10, 80
    hist = _hist(bx, first, n_cols)
    new = slabs.replan_slabs(plan, first, hist, max_move=10 ** 9)
    assert new[0][0] == -slabs.FAR and new[-1][1] == slabs.FAR
    assert all(new[r][1] == new[r + 1][0] and new[r][0] < new[r][1] for r in range(3))
    for k in range(1, 4):                                   # every cut strictly inside the two old slabs it separates
        assert plan[k - 1][0] < new[k][0] < plan[k][1]

    def spread(p):
        cuts = np.array([q[0] for q in p[1:]])
        c = np.bincount(np.searchsorted(cuts, bx, side="right"), minlength=4)
        return c.max() - c.min()
    assert spread(new) < spread(plan)
    again = new
    for _ in range(8):                                      # repeated rebalancing converges to near-equal counts
        again = slabs.replan_slabs(again, first, hist, max_move=10 ** 9)
    assert spread(again) <= 2 * hist.max()
    assert slabs.replan_slabs(again, first, hist, max_move=10 ** 9) == again
    # a tight mailbox limits how many particles one step hands over
    limited = slabs.replan_slabs(plan, first, hist, max_move=2000)
    below = np.cumsum(hist)
    for k in range(1, 4):
        assert abs(int(below[limited[k][0] - first]) - int(below[plan[k][0] - first])) <= 2000
    assert slabs.replan_slabs(plan, first, hist, max_move=0) == plan
    with pytest.raises(ValueError):
        slabs.replan_slabs([(-slabs.FAR, 500), (500, slabs.FAR)], first, hist, max_move=10)
    assert slabs.replan_slabs([(-slabs.FAR, slabs.FAR)], first, hist, 10) == [(-slabs.FAR, slabs.FAR)]
