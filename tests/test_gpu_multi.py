"""`svb_create_multi`: ONE handle over several GPUs, driven from a single thread like the reference's compute thread
(core/src/compute_thread.rs:100-163) — against the single-GPU result and the oracle.  Needs >= 2 GPUs (`gpurun --gpus 2`)."""
import numpy as np
import pytest

from squishy_volumes_b200 import scenes
from squishy_volumes_b200.types import ParticleFlags, RunParameters
from tests import golden_scenes, parity

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _scene(name):
    if name == "jelly":
        sc = scenes.jelly_collision(side=16)
        sc.io_state.particles.velocities[:, 0] *= 6.0          # fast approach: many particles cross the cuts
        return sc, 30, False
    if name == "jelly_adaptive":
        sc = scenes.jelly_collision(side=16)
        sc.io_state.particles.velocities[:, 0] *= 6.0
        sc.time_step = 4e-3
        return sc, 40, True
    if name == "sand":
        sc, _, _ = golden_scenes.GOLDEN["sand_torus"]()
        return sc, 40, False
    sc, _, _ = golden_scenes.GOLDEN["split_layers"]()
    return sc, 25, False


@pytest.mark.parametrize("name", ["jelly", "jelly_adaptive", "sand", "split_layers"])
def test_one_handle_over_all_gpus_matches_single_gpu_and_oracle(name):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    import oracle.oracle as orc
    from squishy_volumes_b200.state import B200State
    sc, steps, adaptive = _scene(name)
    sc.frame_input.consts.frames_per_second = 1
    dt = sc.time_step
    params = RunParameters(target_time=steps * 1e-3, max_time_step=dt, adaptive_time_steps=True) if adaptive else RunParameters(target_time=(steps - 0.5) * dt, max_time_step=dt)
    h = sc.frame_input.consts.scaled_grid_node_size()
    multi = B200State.from_io_state_multi(sc.io_state, sc.frame_input, devices=list(range(min(n, 4))))
    got, err = multi.produce_next_state(None, sc.frame_input, params)
    single = B200State.from_io_state(sc.io_state, sc.frame_input, device=0)
    want, err1 = single.produce_next_state(None, sc.frame_input, params)
    assert err is None and err1 is None
    assert multi.substeps == single.substeps and (adaptive or multi.substeps == steps)
    assert multi.time == pytest.approx(single.time, rel=1e-5)
    assert np.array_equal(got.particles.flags, want.particles.flags)
    live = (want.particles.flags & ParticleFlags.TOMBSTONED) == 0
    mism = int(np.count_nonzero(got.particles.collider_bits[live] != want.particles.collider_bits[live]))
    assert mism <= 2e-3 * sc.n, mism
    for f in ("mass", "initial_volume", "mu_or_bulk_modulus", "lambda_or_exponent", "sand_alpha", "initial_positions"):
        assert np.array_equal(getattr(got.particles, f), getattr(sc.io_state.particles, f)), f       # carried through split / migrate / assemble untouched
    rep = parity.assert_percentiles(got.particles, want.particles, h, parity.PCT_RUN, live, label=f"multi vs single {name}")
    o = orc.OracleState.from_io_state(sc.io_state, sc.frame_input)
    ro, _ = o.produce_next_state(None, sc.frame_input, params)
    parity.assert_percentiles(got.particles, ro.particles, h, parity.PCT_RUN, live, label=f"multi vs oracle {name}")
    # a second frame on the same handle, then a new state into it (svb_upload re-plans the slabs)
    params2 = RunParameters(target_time=params.target_time + 5 * (1e-3 if adaptive else dt), max_time_step=dt, adaptive_time_steps=adaptive)
    got2, err = multi.produce_next_state(None, sc.frame_input, params2)
    want2, _ = single.produce_next_state(None, sc.frame_input, params2)
    assert err is None and multi.substeps == single.substeps
    parity.assert_percentiles(got2.particles, want2.particles, h, parity.PCT_RUN, live, label=f"multi vs single, second frame {name}")
    multi.upload(sc.io_state)
    again, err = multi.produce_next_state(None, sc.frame_input, params)
    assert err is None and np.array_equal(again.particles.flags, want.particles.flags)
    parity.assert_percentiles(again.particles, want.particles, h, parity.PCT_RUN, live, label=f"multi after upload {name}")
    print(name, "multi-device handle over", min(n, 4), "GPUs: substeps", multi.substeps, "bit mismatches", mism, "percentiles vs single GPU", rep)
    multi.close()
    single.close()
