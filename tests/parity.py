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

# ---- per-particle relative errors (so that a particle with a small v or C cannot hide behind the field's maximum):
#   e_i = ||got_i - ref_i||_2 / max(||ref_i||_2, floor(field))
# with the absolute floor of each field stated here (below it a value is summation noise, not signal):
#   positions            floor = h            (the error is measured in grid cells: the origin is arbitrary)
#   velocities           floor = 1e-3 * max_i ||v_i||
#   position_gradients   floor = 1            (F is O(1) by construction, I at rest)
#   velocity_gradients   floor = 1e-3 * max_i ||v_i|| / h     (C = 4/h^2 sum w v (x_n - x)^T enters the path as C*h next to v — scatter_momentum.rs:60,
#                        advance_particles.rs:44 — so the floor is the velocity floor divided by h; in rigid motion C is pure summation noise)
# Bounds asserted on the percentiles of e over the live particles (P50, P99, P99.9, max):
PCT = (50.0, 99.0, 99.9, 100.0)
# (measured on B200, profiles/r2_parity_percentiles.txt: one substep x/v/F max 1e-7 / 6e-7 / 1.2e-7, C max 1.8e-3; 12-40 substeps with
#  colliders, sand and fluid: x/v/F max 4e-6 / 7e-4 / 4e-6, C max 1e-2 — the bounds leave about one order of magnitude)
PCT_STEP = {"positions": (2e-7, 1e-6, 1e-6, 2e-6), "velocities": (1e-6, 3e-6, 5e-6, 1e-5), "position_gradients": (5e-7, 1e-6, 2e-6, 5e-6),
            "velocity_gradients": (1e-3, 3e-3, 4e-3, 5e-3)}            # one substep from identical state
# (240 substeps of config 1 / 100 substeps of config 2 in contact: x P50 1.6e-5 max 1.6e-4, v P50 3e-6 P99.9 4e-3 max 2.4e-2, F max 3e-5)
PCT_RUN = {"positions": (1e-4, 5e-4, 1e-3, 5e-3), "velocities": (1e-4, 2e-3, 1e-2, 5e-2), "position_gradients": (1e-4, 5e-4, 1e-3, 5e-3),
           "velocity_gradients": (1e-2, 3e-2, 5e-2, 1e-1)}             # tens to hundreds of substeps (errors compound through contact)


def particle_relative_errors(got, ref, name, h, mask=None):
    """Per-particle relative error of one field (see the floors above) -> 1-d float64 array over the masked particles."""
    def get(o, f):   # Particles, or a mapping of arrays (a loaded golden .npz)
        return getattr(o, f) if hasattr(o, f) else o[f]
    a = np.asarray(get(got, name), dtype=np.float64)
    b = np.asarray(get(ref, name), dtype=np.float64)
    v = np.asarray(get(ref, "velocities"), dtype=np.float64)
    n = a.shape[0]
    a = a.reshape(n, -1)
    b = b.reshape(n, -1)
    v = v.reshape(n, -1)
    if mask is not None:
        a, b, v = a[mask], b[mask], v[mask]
    if a.shape[0] == 0:
        return np.zeros(0)
    mag = np.linalg.norm(b, axis=1)
    vmax = float(np.linalg.norm(v, axis=1).max())
    if name == "positions":
        floor = float(h)
        mag = np.zeros_like(mag)
    elif name == "velocities":
        floor = 1e-3 * vmax
    elif name == "position_gradients":
        floor = 1.0
    else:
        floor = 1e-3 * vmax / float(h)
    floor = max(floor, 1e-30)
    return np.linalg.norm(a - b, axis=1) / np.maximum(mag, floor)


def error_percentiles(got, ref, h, mask=None):
    """{field: (P50, P99, P99.9, max)} of the per-particle relative errors."""
    out = {}
    for name in FIELDS:
        e = particle_relative_errors(got, ref, name, h, mask)
        out[name] = tuple(float(np.percentile(e, q)) for q in PCT) if e.size else (0.0,) * len(PCT)
    return out


def assert_percentiles(got, ref, h, bounds, mask=None, label=""):
    rep = error_percentiles(got, ref, h, mask)
    for name, vals in rep.items():
        for q, v, bound in zip(PCT, vals, bounds[name]):
            assert v <= bound, f"{label} {name}: P{q:g} of the per-particle relative error = {v:.3e} > {bound:.1e}   (all: {vals})"
    return rep


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


def compare_states(got, ref, rtol, atol=ATOL, check_energy=True, h=None, pct=None):
    """got / ref: IoState.  Returns dict of (err, scale); raises AssertionError on violation.
    With `h` given, the per-particle relative error percentiles are bounded as well (`pct`: PCT_STEP below rtol 1e-3, else PCT_RUN)."""
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
    if h is not None:
        report["percentiles"] = assert_percentiles(gp, rp, h, pct if pct is not None else (PCT_STEP if rtol < 1e-3 else PCT_RUN), live)
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
