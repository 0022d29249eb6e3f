"""Shared helpers of the parity tests: run the oracle and the CUDA path on the same scene and
compare.  Tolerances are stated here once:

* integer stages (cell keys, per-cell membership, active block set, flags, collider bits): exact;
* positions / velocities / F / C / energy: |a-b| <= ATOL + RTOL * scale, where `scale` is the field's
  max magnitude over the scene (float atomics reorder sums, FMA contraction differs from rustc).
  The reference's own GPU-vs-CPU bar is rel 1e-2 / abs 1e-6 (gpu/src/test_util.rs:2731-2736); ours
  is two orders tighter per substep.
"""
from __future__ import annotations

import numpy as np

from squishy_volumes_b200.types import ParticleFlags, RunParameters

RTOL_STEP = 2e-4   # one substep
RTOL_RUN = 2e-3    # tens of substeps
ATOL = 1e-6

FIELDS = ("positions", "velocities", "position_gradients", "velocity_gradients")


def field_error(a: np.ndarray, b: np.ndarray, mask=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if mask is not None:
        a = a[mask]
        b = b[mask]
    if a.size == 0:
        return 0.0, 1.0
    scale = max(float(np.max(np.abs(b))), 1e-12)
    return float(np.max(np.abs(a - b))), scale


def compare_states(got, ref, rtol, atol=ATOL, check_energy=True, h=None):
    """got / ref: IoState.  Returns dict of (err, scale); raises AssertionError on violation."""
    gp, rp = got.particles, ref.particles
    report = {}
    assert np.array_equal(gp.flags, rp.flags), f"flags differ at {np.nonzero(gp.flags != rp.flags)[0][:10]}"
    live = (rp.flags & ParticleFlags.TOMBSTONED) == 0
    assert np.array_equal(gp.collider_bits[live], rp.collider_bits[live]), "collider bits differ"
    vmax = max(float(np.max(np.abs(rp.velocities[live]))) if live.any() else 0.0, 1e-12)
    for name in FIELDS:
        err, scale = field_error(getattr(gp, name), getattr(rp, name), live)
        if name == "velocity_gradients" and h is not None:
            scale = max(scale, vmax / h)  # C = 4/h^2 sum w v (x_n - x)^T: cancellation noise scales with |v|/h
        report[name] = (err, scale)
        assert err <= atol + rtol * scale, f"{name}: max abs err {err:.3e} vs scale {scale:.3e} (rtol {rtol})"
    if check_energy:
        # energies are differences of O(modulus) terms: absolute floor = f32 eps * modulus
        modulus = float(np.max(np.abs(rp.mu_or_bulk_modulus) + np.where((rp.flags & ParticleFlags.IS_FLUID) != 0, 0.0, np.abs(rp.lambda_or_exponent))))
        ok = live & ((rp.flags & ParticleFlags.FAILED) == 0)
        err, scale = field_error(gp.elastic_energies, rp.elastic_energies, ok)
        report["elastic_energies"] = (err, scale)
        assert err <= 4e-6 * modulus + rtol * scale, f"elastic_energies: max abs err {err:.3e} vs scale {scale:.3e}, modulus {modulus:.3e}"
    return report


def cells_by_original(state):
    """(n,3) int32 base nodes indexed by ORIGINAL particle index, from a B200State or OracleState."""
    sm, cells = state.binning()
    out = np.zeros_like(cells)
    out[sm] = cells
    return out


def node_blocks(grid, min_count=1):
    """Set of (bx,by,bz,bits) derived from a GridNodes with contributor counts."""
    keep = grid.contributor_counts >= min_count if grid.contributor_counts is not None else np.ones(len(grid.masses), bool)
    ids = grid.node_ids[keep] >> 2
    bits = grid.collider_bits[keep]
    return set(map(tuple, np.concatenate([ids, bits[:, None].astype(np.int64)], axis=1).tolist()))


def run_both(scene, n_substeps, adaptive=False, store_grid=False, device=0):
    import oracle.oracle as orc
    from squishy_volumes_b200.state import B200State
    dt = scene.time_step
    target = scene.io_state.time + dt * (n_substeps - 0.5) if not adaptive else scene.io_state.time + dt * n_substeps
    params = RunParameters(target_time=target, max_time_step=dt, adaptive_time_steps=adaptive, store_grid=store_grid)
    o = orc.OracleState.from_io_state(scene.io_state, scene.frame_input)
    g = B200State.from_io_state(scene.io_state, scene.frame_input, device=device)
    ro, eo = o.produce_next_state(None, scene.frame_input, params)
    rg, eg = g.produce_next_state(None, scene.frame_input, params)
    return (o, ro, eo), (g, rg, eg)
