"""Frame files, input files, scene set-up and keyframes (include/svb_files.h, csrc/svb_files.cpp) against the independent
format oracle (oracle/bincode_ref.py), byte for byte, and the reference's own test cases for the input file
(rust/crates/file_input/src/tests.rs:70-290, rust/crates/file_util/src/lib.rs:103-145).  CPU only."""
import os
import struct

import numpy as np
import pytest

from oracle import bincode_ref as ref
from squishy_volumes_b200 import files
from squishy_volumes_b200.types import GridNodes, InputConsts, IoState, Particles

CONSTS = dict(grid_node_size=0.5, leaf_size=1.0, leaf_threshold=16, simulation_scale=1.0, frames_per_second=24,
              domain_min=(-100.0,) * 3, domain_max=(100.0,) * 3)          # InputConsts::test_input, header.rs:22-33


def consts_obj(d=CONSTS):
    return InputConsts(**d)


def random_particles(n, seed=7):
    rng = np.random.default_rng(seed)
    p = Particles.empty(n)
    kinds = np.array([1, 1 | 4, 1 | 8, 1 | 4 | 8, 2, 2 | 4, 1 | 16, 2 | 32, 1 | 64], np.uint32)     # every Option / variant combination
    p.flags[:] = kinds[rng.integers(0, len(kinds), n)]
    for name in ("mass", "initial_volume", "mu_or_bulk_modulus", "sand_alpha", "viscosity_dynamic", "viscosity_bulk", "elastic_energies"):
        getattr(p, name)[:] = rng.random(n, dtype=np.float32) + 0.1
    fluid = (p.flags & 2) != 0
    p.lambda_or_exponent[:] = np.where(fluid, rng.integers(2, 9, n).astype(np.float32), rng.random(n, dtype=np.float32) * 100)
    # fields a flag switches off are not stored: keep them zero so that a round trip is the identity
    p.sand_alpha[(p.flags & 8) == 0] = 0
    p.sand_alpha[fluid] = 0
    p.viscosity_dynamic[(p.flags & 4) == 0] = 0
    p.viscosity_bulk[(p.flags & 4) == 0] = 0
    p.collider_bits[:] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    for name in ("positions", "velocities", "initial_positions"):
        getattr(p, name)[:] = rng.standard_normal((n, 3)).astype(np.float32)
    p.position_gradients[:] = rng.standard_normal((n, 3, 3)).astype(np.float32)
    p.velocity_gradients[:] = rng.standard_normal((n, 3, 3)).astype(np.float32)
    return p


def random_grid(g, seed=11):
    rng = np.random.default_rng(seed)
    return GridNodes(rng.integers(-50, 50, (g, 3)).astype(np.int32), rng.integers(0, 2 ** 32, g, dtype=np.uint64).astype(np.uint32),
                     rng.random(g, dtype=np.float32), rng.standard_normal((g, 3)).astype(np.float32))


def same_particles(a: Particles, b: Particles):
    import dataclasses
    for f in dataclasses.fields(Particles):
        assert np.array_equal(getattr(a, f.name), getattr(b, f.name)), f.name


# ------------------------------------------------------------------------------------------------ frame files
@pytest.mark.parametrize("n,g", [(0, None), (1, 0), (257, None), (1000, 333)])
def test_frame_bytes_equal_the_format_oracle(tmp_path, n, g):
    p = random_particles(n)
    grid = None if g is None else random_grid(g)
    path = tmp_path / "frame_00003.bin"
    written = files.write_frame(str(path), IoState(0.125, p, grid))
    want = ref.encode_io_state(0.125, p, grid)
    got = path.read_bytes()
    assert written == len(got) == len(want)
    assert got == want
    assert not (tmp_path / "temp.bin").exists()                 # io_state.rs:33-60: written as temp.bin, then renamed
    back = files.read_frame(str(path))
    assert back.time == 0.125
    same_particles(back.particles, p)
    if grid is None:
        assert back.grid_nodes is None
    else:
        for k in ("node_ids", "collider_bits", "masses", "velocities"):
            assert np.array_equal(getattr(back.grid_nodes, k), getattr(grid, k))
    d = ref.decode_io_state(got)                                # and the oracle reads what the library wrote
    assert np.array_equal(d["lambda_or_exponent"], p.lambda_or_exponent) and np.array_equal(d["sand_alpha"], p.sand_alpha)


def test_frame_reader_accepts_oracle_bytes_and_rejects_damage(tmp_path):
    p = random_particles(64, seed=3)
    good = ref.encode_io_state(2.5, p, random_grid(5))
    path = tmp_path / "f.bin"
    path.write_bytes(good)
    same_particles(files.read_frame(str(path)).particles, p)
    bad = bytearray(good); bad[5:9] = b"\x01\x02\x03\x04"
    path.write_bytes(bad)
    with pytest.raises(files.FileError) as e:
        files.read_frame(str(path))
    assert e.value.status == -21                                 # MagicMismatch
    bad = bytearray(good); bad[32 + 2] = 0                        # file_util tests: wrong_version
    path.write_bytes(bad)
    with pytest.raises(files.FileError) as e:
        files.read_frame(str(path))
    assert e.value.status == -22
    with pytest.raises(files.FileError) as e:                    # a reader built for another version refuses the file
        files.read_frame(str(path), version="9.9.9")
    assert e.value.status == -22
    for cut in (40, 100, 104, 200, len(good) // 2, len(good) - 1):
        path.write_bytes(good[:cut])
        with pytest.raises(files.FileError) as e:
            files.read_frame(str(path))
        assert e.value.status in (-20, -23), cut
    bad = bytearray(good); struct.pack_into("<Q", bad, 96 + 8, 2 ** 60)   # an absurd Vec length must not allocate
    path.write_bytes(bad)
    with pytest.raises(files.FileError) as e:
        files.read_frame(str(path))
    assert e.value.status == -23
    with pytest.raises(files.FileError) as e:
        files.read_frame(str(tmp_path / "missing.bin"))
    assert e.value.status == -20


def test_frame_path_and_version():
    assert files.frame_path("/cache", 7) == "/cache/frame_00007.bin"        # cache/src/util.rs:9-11
    assert files.frame_path("/cache", 123456) == "/cache/frame_123456.bin"
    assert files.default_version() == ref.VERSION


# ------------------------------------------------------------------------------------------------ input files
def particles_input(n, rng, with_everything=True):
    """ParticlesInput::random (frame.rs:29-48) with parameters kept inside their bounds."""
    fl = np.where(rng.random(n) < 0.5, np.uint32(1 | 4 | 8), np.uint32(2 | 4)).astype(np.uint32)
    d = {"flags": fl,
         "transforms": rng.standard_normal((n, 4, 4)).astype(np.float32), "sizes": (rng.random(n) + 0.5).astype(np.float32),
         "densities": (rng.random(n) * 1000 + 1).astype(np.float32), "youngs_moduluses": (rng.random(n) * 1e5).astype(np.float32),
         "poissons_ratios": (rng.random(n) * 0.49).astype(np.float32), "initial_positions": rng.standard_normal((n, 3)).astype(np.float32),
         "initial_velocities": rng.standard_normal((n, 3)).astype(np.float32), "viscosities_dynamic": rng.random(n).astype(np.float32),
         "viscosities_bulk": rng.random(n).astype(np.float32), "exponents": rng.integers(2, 9, n).astype(np.uint32),
         "bulk_moduluses": (rng.random(n) * 1e3).astype(np.float32), "sand_alphas": rng.random(n).astype(np.float32),
         "goal_positions": rng.standard_normal((n, 3)).astype(np.float32)}
    if not with_everything:
        for k in ("initial_positions", "goal_positions", "sand_alphas", "exponents", "bulk_moduluses"):
            d[k] = None
        d["flags"] = np.full(n, 1 | 4, np.uint32)
    return d


def collider_input(nv, nt, rng):
    return {"vertex_positions": rng.standard_normal((nv, 3)).astype(np.float32), "triangle_indices": rng.integers(0, max(nv, 1), (nt, 3)).astype(np.uint32),
            "triangle_frictions": rng.random(nt).astype(np.float32), "triangle_dampings": rng.random(nt).astype(np.float32)}


def header_objects(n=10, nv=9, nt=3):
    return {"foo": ("particles", n), "bar": ("particles", n), "car": ("collider", nv, nt)}      # tests.rs:25-43


def make_frames(n=10, nv=9, nt=3, seed=42):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(3):
        out.append({"gravity": (0.0, 0.0, -float(k)), "particles": {"foo": particles_input(n, rng), "bar": particles_input(n, rng, with_everything=(k != 1))},
                    "colliders": {"car": collider_input(nv, nt, rng)}})
    return out


def write_full(path, n=10, nv=9, nt=3):
    frames = make_frames(n, nv, nt)
    w = files.InputWriter(str(path), consts_obj(), header_objects(n, nv, nt))
    for fr in frames:
        w.record_frame(fr["gravity"], fr["particles"], fr["colliders"])
    w.finish()
    return frames


def test_input_writer_bytes_equal_the_format_oracle(tmp_path):
    path = tmp_path / "simulation_input.bin"
    frames = write_full(path)
    assert path.read_bytes() == ref.encode_input_file(CONSTS, header_objects(), frames)


def test_input_write_start_and_partial(tmp_path):          # tests.rs:70-83
    files.InputWriter(str(tmp_path / "a.bin"), consts_obj(), header_objects(100, 99, 33))
    w = files.InputWriter(str(tmp_path / "b.bin"), consts_obj(), header_objects(100, 99, 33))
    for fr in make_frames(100, 99, 33)[:2]:
        w.record_frame(fr["gravity"], fr["particles"], fr["colliders"])


def test_input_read_header_and_random_frames(tmp_path):    # tests.rs:109-141
    path = tmp_path / "simulation_input.bin"
    path.write_bytes(ref.encode_input_file(CONSTS, header_objects(), make_frames()))      # written by the oracle, read by the library
    f = files.InputFile(str(path))
    assert f.n_frames == 3 and f.size == path.stat().st_size
    assert f.consts == consts_obj()
    assert [(o.name, o.kind, o.count, o.count2, o.start, o.start2) for o in f.objects] == [("bar", 0, 10, 0, 0, 0), ("car", 1, 9, 3, 0, 0), ("foo", 0, 10, 0, 10, 0)]
    assert (f.total_particles, f.total_vertices, f.total_triangles) == (20, 9, 3)
    frames = make_frames()
    for idx in np.random.default_rng(42).permutation(3):
        want = ref.keyframe_ref(CONSTS, header_objects(), frames[idx])
        k = f.keyframe(int(idx))
        assert tuple(k.gravity) == tuple(np.float32(want["gravity"]))
        for name in ("particle_flags", "particle_goal_positions", "vertex_positions", "triangle_frictions", "triangle_dampings"):
            assert np.array_equal(getattr(k, name), want[name]), name
    topo = f.topology()
    assert len(topo) == 1 and topo[0].num_vertices == 9 and np.array_equal(topo[0].triangles, frames[0]["colliders"]["car"]["triangle_indices"])
    fi = f.frame_input()
    assert len(fi.keyframes) == 3 and fi.num_vertices() == 9 and fi.num_triangles() == 3


def test_input_wrong_magic_version_and_index(tmp_path):    # tests.rs:143-222
    path = tmp_path / "simulation_input.bin"
    write_full(path)
    good = path.read_bytes()

    def status_of(data):
        path.write_bytes(data)
        with pytest.raises(files.FileError) as e:
            files.InputFile(str(path))
        return e.value.status
    bad = bytearray(good); bad[5:9] = bytes([1, 2, 3, 4])
    assert status_of(bad) == -21
    bad = bytearray(good); bad[32 + 5:32 + 9] = bytes([1, 2, 3, 4])
    assert status_of(bad) == -22
    bad = bytearray(good); bad[-8:] = b"\xff" * 8          # index offset no seek can reach: IoError in the reference
    assert status_of(bad) == -20
    bad = bytearray(good); bad[-8:] = b"\x00" * 8          # index "at" the magic bytes: BincodeError in the reference
    assert status_of(bad) == -23
    assert status_of(good[:50]) in (-20, -23)
    path.write_bytes(good)
    f = files.InputFile(str(path))
    with pytest.raises(files.FileError) as e:               # tests.rs:179-188
        f.keyframe(3)
    assert e.value.status == -24


def test_input_frame_verification(tmp_path):                # tests.rs:224-290
    rng = np.random.default_rng(1)
    w = files.InputWriter(str(tmp_path / "v.bin"), consts_obj(), header_objects(10, 9, 3))
    with pytest.raises(files.FileError) as e:               # test_length_mismatch
        w.record_frame((0, 0, 0), {"foo": particles_input(1, rng), "bar": particles_input(1, rng)}, {"car": collider_input(2, 3, rng)})
    assert e.value.status == -28
    with pytest.raises(files.FileError) as e:               # test_collider_missing
        w.record_frame((0, 0, 0), {"foo": particles_input(10, rng)}, {})
    assert e.value.status == -29
    with pytest.raises(files.FileError) as e:               # test_object_changed_type
        w.record_frame((0, 0, 0), {}, {"car": collider_input(9, 3, rng), "foo": collider_input(0, 0, rng)})
    assert e.value.status == -25 and "changed type" in e.value.message
    with pytest.raises(files.FileError) as e:               # test_object_not_in_header
        w.record_frame((0, 0, 0), {}, {"car": collider_input(9, 3, rng), "newfoo": collider_input(0, 0, rng)})
    assert e.value.status == -25 and "not in header" in e.value.message
    w.record_frame((0, 0, 0), {"foo": particles_input(10, rng)}, {"car": collider_input(9, 3, rng)})   # a frame may leave particle objects out
    w.finish()
    assert files.InputFile(str(tmp_path / "v.bin")).n_frames == 1


# ------------------------------------------------------------------------------------------------ scene set-up
def test_initialize_io_state_matches_the_restatement(tmp_path):
    consts = dict(CONSTS, simulation_scale=2.5)
    objects = {"jelly": ("particles", 37), "water": ("particles", 12), "floor": ("collider", 4, 2)}
    rng = np.random.default_rng(5)
    frames = [{"gravity": (0, 0, -9.8), "particles": {"jelly": particles_input(37, rng), "water": particles_input(12, rng)},
               "colliders": {"floor": collider_input(4, 2, rng)}}]
    path = tmp_path / "simulation_input.bin"
    path.write_bytes(ref.encode_input_file(consts, objects, frames))
    f = files.InputFile(str(path))
    st = f.initialize_io_state()
    want = ref.initialize_io_state_ref(consts, objects, frames[0])
    assert st.time == 0.0 and st.grid_nodes is not None and st.grid_nodes.masses.shape[0] == 0        # initialization.rs:274: grid_nodes = Some(default)
    for k, v in want.items():
        assert np.array_equal(getattr(st.particles, k), v), k
    kf = f.keyframe(0)
    want_k = ref.keyframe_ref(consts, objects, frames[0])
    assert np.array_equal(kf.particle_goal_positions, want_k["particle_goal_positions"]) and np.array_equal(kf.vertex_positions, want_k["vertex_positions"])


@pytest.mark.parametrize("damage,status", [
    (lambda d: d.update(transforms=None), -27),                                         # MissingInput (object level)
    (lambda d: d.update(flags=np.full(8, 3, np.uint32)), -26),                          # SolidXorFluid
    (lambda d: d.update(flags=np.full(8, 0, np.uint32)), -26),
    (lambda d: d.update(flags=np.full(8, 1 | 128, np.uint32)), -26),                    # UnknownFlagsSet
    (lambda d: d.update(flags=np.full(8, 1, np.uint32), youngs_moduluses=None), -27),   # ParticleInvalid::MissingInput
    (lambda d: d.update(flags=np.full(8, 1, np.uint32), poissons_ratios=np.full(8, 0.5, np.float32)), -26),   # EnergyError bounds
    (lambda d: d.update(flags=np.full(8, 1, np.uint32), youngs_moduluses=np.full(8, -1.0, np.float32)), -26),
    (lambda d: d.update(flags=np.full(8, 2, np.uint32), exponents=np.full(8, 1, np.uint32)), -26),
    (lambda d: d.update(flags=np.full(8, 2, np.uint32), bulk_moduluses=np.full(8, -2.0, np.float32)), -26),
    (lambda d: d.update(flags=np.full(8, 1 | 4, np.uint32), viscosities_bulk=None), -27),
    (lambda d: d.update(flags=np.full(8, 1 | 8, np.uint32), sand_alphas=None), -27),
])
def test_initialize_io_state_rejects_invalid_particles(tmp_path, damage, status):
    rng = np.random.default_rng(9)
    d = particles_input(8, rng)
    damage(d)
    path = tmp_path / "simulation_input.bin"
    path.write_bytes(ref.encode_input_file(CONSTS, {"p": ("particles", 8)}, [{"gravity": (0, 0, 0), "particles": {"p": d}, "colliders": {}}]))
    with pytest.raises(files.FileError) as e:
        files.InputFile(str(path)).initialize_io_state()
    assert e.value.status == status, e.value.message


def test_recorded_scene_round_trips_through_the_files(tmp_path):
    """A synthetic scene recorded as an input file and read back gives the scene's IoState and keyframes (the bench scenes
    use unit transforms: F = I, x = translation), and its result frame reads back unchanged."""
    from squishy_volumes_b200 import scenes
    sc = scenes.elastic_cube(side=6, h=0.1, n_keyframes=3)
    p = sc.io_state.particles
    n = p.n
    t = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    t[:, 3, :3] = p.positions
    size = np.cbrt(p.initial_volume.astype(np.float64)).astype(np.float32)
    topo = sc.frame_input.colliders
    assert len(topo) == 1
    w = files.InputWriter(str(tmp_path / "simulation_input.bin"), sc.frame_input.consts,
                          {"cube": ("particles", n), "ground": ("collider", topo[0].num_vertices, topo[0].triangles.shape[0])})
    E, nu = 1e4, 0.3
    for k in sc.frame_input.keyframes:
        w.record_frame(k.gravity, {"cube": {"flags": p.flags, "transforms": t, "sizes": size, "densities": p.mass / (size * size * size),
                                            "youngs_moduluses": np.full(n, E, np.float32), "poissons_ratios": np.full(n, nu, np.float32),
                                            "initial_velocities": p.velocities, "initial_positions": p.initial_positions}},
                       {"ground": {"vertex_positions": k.vertex_positions, "triangle_indices": topo[0].triangles, "triangle_frictions": k.triangle_frictions,
                                   "triangle_dampings": k.triangle_dampings}})
    w.finish()
    f = files.InputFile(str(tmp_path / "simulation_input.bin"))
    st = f.initialize_io_state()
    assert np.array_equal(st.particles.positions, p.positions) and np.array_equal(st.particles.position_gradients, p.position_gradients)
    assert np.array_equal(st.particles.velocities, p.velocities) and np.array_equal(st.particles.flags, p.flags)
    assert np.allclose(st.particles.initial_volume, p.initial_volume, rtol=1e-6) and np.allclose(st.particles.mass, p.mass, rtol=1e-6)
    assert np.allclose(st.particles.mu_or_bulk_modulus, p.mu_or_bulk_modulus, rtol=1e-6) and np.allclose(st.particles.lambda_or_exponent, p.lambda_or_exponent, rtol=1e-6)
    fi = f.frame_input()
    assert len(fi.keyframes) == 3 and np.array_equal(fi.keyframes[1].vertex_positions, sc.frame_input.keyframes[1].vertex_positions)
    assert np.array_equal(fi.colliders[0].triangles, topo[0].triangles)
    out = tmp_path / files.frame_path(".", 0)
    files.write_frame(str(out), st)
    same_particles(files.read_frame(str(out)).particles, st.particles)


# ------------------------------------------------------------------------------------------------ committed fixtures
def test_golden_files_are_reproduced_and_read(tmp_path):
    """tests/golden/{frame,input}_small.bin (written by `python -m tests.golden_files` with the format oracle) are reproduced
    byte for byte by the library's writers and read back by its readers: drift in either implementation shows up here."""
    from tests import golden_files as gf
    golden_frame = open(os.path.join(gf.HERE, "frame_small.bin"), "rb").read()
    golden_input = open(os.path.join(gf.HERE, "input_small.bin"), "rb").read()
    t, p, g = gf.frame_inputs()
    assert ref.encode_io_state(t, p, g) == golden_frame
    out = tmp_path / "frame_00001.bin"
    files.write_frame(str(out), IoState(t, p, g))
    assert out.read_bytes() == golden_frame
    st = files.read_frame(os.path.join(gf.HERE, "frame_small.bin"))
    assert st.time == t and st.grid_nodes.masses.shape[0] == 11
    same_particles(st.particles, p)
    frames = gf.input_frames()
    assert ref.encode_input_file(gf.CONSTS, gf.OBJECTS, frames) == golden_input
    w = files.InputWriter(str(tmp_path / "in.bin"), consts_obj(gf.CONSTS), gf.OBJECTS)
    for fr in frames:
        w.record_frame(fr["gravity"], fr["particles"], fr["colliders"])
    w.finish()
    assert (tmp_path / "in.bin").read_bytes() == golden_input
    f = files.InputFile(os.path.join(gf.HERE, "input_small.bin"))
    assert [o.name for o in f.objects] == ["ball", "floor", "jelly", "water"] and f.n_frames == 2
    assert (f.total_particles, f.total_vertices, f.total_triangles) == (17, 10, 10)
    want = ref.initialize_io_state_ref(gf.CONSTS, gf.OBJECTS, frames[0])
    got = f.initialize_io_state().particles
    for k, v in want.items():
        assert np.array_equal(getattr(got, k), v), k
    topo = f.topology()
    assert [t_.num_vertices for t_ in topo] == [6, 4]          # colliders in name order: ball, floor
