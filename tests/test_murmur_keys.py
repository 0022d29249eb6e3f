"""the reference's node key — murmur3_x86_32 of the ordered-u32 node id, seed = collider bits
(gpu/src/util.rs:71-100; crate murmur3 0.5.2 is not vendored, so the published MurmurHash3_x86_32 algorithm is restated here in
plain Python and first checked against its published verification vectors) — against the product's implementation: on the host
through the math shim, on the GPU through `svb_node_ids_to_murmur` (the reference's node_ids_to_murmur stage, test.rs:15-96), and as
the tile table's hash function (option "murmur_table_hash": results must not depend on it)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
M32 = 0xffffffff


def murmur3_32(data: bytes, seed: int) -> int:
    """MurmurHash3_x86_32 (Austin Appleby, public domain), restated."""
    rotl = lambda x, r: ((x << r) | (x >> (32 - r))) & M32
    h = seed & M32
    n = len(data) // 4
    for i in range(n):
        k = int.from_bytes(data[4 * i:4 * i + 4], "little")
        k = (k * 0xcc9e2d51) & M32
        k = rotl(k, 15)
        k = (k * 0x1b873593) & M32
        h ^= k
        h = rotl(h, 13)
        h = (h * 5 + 0xe6546b64) & M32
    tail = data[4 * n:]
    k = 0
    if len(tail) >= 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * 0xcc9e2d51) & M32
        k = rotl(k, 15)
        k = (k * 0x1b873593) & M32
        h ^= k
    h ^= len(data)
    h ^= h >> 16
    h = (h * 0x85ebca6b) & M32
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & M32
    h ^= h >> 16
    return h


def reference_key(node_id, seed):
    """gpu/src/util.rs:79-100: x ^ 0x8000_0000 per coordinate, little-endian bytes, murmur3_32 with the seed."""
    b = b"".join(((int(c) & M32) ^ 0x80000000).to_bytes(4, "little") for c in node_id)
    return murmur3_32(b, int(seed))


SIMPLE_IDS = np.array([[-5, -5, -5], [-5, -5, 5], [-5, 5, -5], [-5, 5, 5], [5, -5, -5], [5, -5, 5], [5, 5, -5], [5, 5, 5]], np.int32)   # node_ids_to_murmur/test.rs:42-52
SIMPLE_BITS = np.array([0x00000, 0x10000, 0x20000, 0x30000, 0x40000, 0x50000, 0x60000, 0x70000], np.uint32)                          # :54-56


def cases():
    rng = np.random.default_rng(42)   # (the reference draws 1000 random i32 triples and u32 bits from ChaCha8(42); any seeded stream serves)
    ids = np.concatenate([SIMPLE_IDS, rng.integers(-2**31, 2**31, (1000, 3), dtype=np.int64).astype(np.int32),
                          np.array([[0, 0, 0], [-1, -1, -1], [2**31 - 1, -2**31, 0]], np.int32)])
    bits = np.concatenate([SIMPLE_BITS, rng.integers(0, 2**32, 1000, dtype=np.int64).astype(np.uint32), np.array([0, 0xffffffff, 0x00010001], np.uint32)])
    return ids, bits


def test_python_murmur3_against_published_vectors():
    for data, seed, want in ((b"", 0, 0), (b"", 1, 0x514e28b7), (b"", 0xffffffff, 0x81f16f39), (b"test", 0, 0xba6bd213), (b"Hello, world!", 0, 0xc0363e43),
                             (b"The quick brown fox jumps over the lazy dog", 0, 0x2e4ff723), (b"\xff\xff\xff\xff", 0, 0x76293b50), (b"\x21\x43\x65\x87", 0, 0xf55b516b),
                             (b"\x21\x43\x65\x87", 0x5082edee, 0x2362f9de), (b"\x21\x43\x65", 0, 0x7e4a8634), (b"\x21\x43", 0, 0xa0f7b07a), (b"\x21", 0, 0x72661cf4),
                             (b"\x00\x00\x00\x00", 0, 0x2362f9de), (b"\x00\x00\x00", 0, 0x85f0b427), (b"\x00\x00", 0, 0x30f4c306), (b"\x00", 0, 0x514e28b7)):
        assert murmur3_32(data, seed) == want, (data, seed)


def test_product_node_key_on_the_host(tmp_path):
    out = str(tmp_path / "math_shim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, os.path.join(HERE, "native", "math_shim.cpp")], check=True)
    L = C.CDLL(out)
    L.shim_node_id_to_murmur.restype = C.c_uint32
    L.shim_node_id_to_murmur.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint32]
    ids, bits = cases()
    for node, b in zip(ids.tolist(), bits.tolist()):
        assert L.shim_node_id_to_murmur(*node, 0) == reference_key(node, 0)
        assert L.shim_node_id_to_murmur(*node, b) == reference_key(node, b)


@pytest.mark.gpu
def test_node_ids_to_murmur_stage_on_the_gpu():
    from squishy_volumes_b200 import abi, cstructs as cs
    ids, bits = cases()
    plain = np.zeros(len(ids), np.uint32)
    seeded = np.zeros(len(ids), np.uint32)
    rc = abi.load().svb_node_ids_to_murmur(0, np.ascontiguousarray(ids).ctypes.data_as(cs.c_i32p), cs.uptr(bits), len(ids), cs.uptr(plain), cs.uptr(seeded))
    assert rc == 0
    assert plain.tolist() == [reference_key(n, 0) for n in ids.tolist()]
    assert seeded.tolist() == [reference_key(n, b) for n, b in zip(ids.tolist(), bits.tolist())]


@pytest.mark.gpu
def test_murmur_as_the_tile_table_hash_changes_nothing():
    """The tile table hashed with the reference's node key: same active tile set, same bins, same particles (integers exactly)."""
    from squishy_volumes_b200 import abi, scenes
    from squishy_volumes_b200.state import B200State
    from squishy_volumes_b200.types import RunParameters
    from tests import golden_scenes, parity
    for sc in (scenes.jelly_collision(side=12), golden_scenes.GOLDEN["split_layers"]()[0]):
        sc.frame_input.consts.frames_per_second = 1
        params = RunParameters(9.5 * sc.time_step, sc.time_step)
        a = B200State.from_io_state(sc.io_state, sc.frame_input)
        b = B200State.from_io_state(sc.io_state, sc.frame_input)
        abi.load().svb_set_option(b._h, b"murmur_table_hash", 1.0)
        ra, _ = a.produce_next_state(None, sc.frame_input, params)
        rb, _ = b.produce_next_state(None, sc.frame_input, params)
        assert np.array_equal(ra.particles.flags, rb.particles.flags) and np.array_equal(ra.particles.collider_bits, rb.particles.collider_bits)
        assert np.array_equal(parity.cells_by_original(a), parity.cells_by_original(b))
        ia, ba = a.active_blocks()
        ib, bb = b.active_blocks()
        assert set(map(tuple, np.column_stack([ia, ba]).tolist())) == set(map(tuple, np.column_stack([ib, bb]).tolist()))
        parity.compare_states(rb, ra, rtol=parity.RTOL_STEP)   # (float atomics: sums in another order, values within the per-substep bound)
