"""CPU check of the PRODUCT's particle layout (squishy_volumes_b200/csrc/svb_device.cuh: Field, ParticleBuf) compiled for the host:
the 34 state words of a particle live in nine 16-byte quads; every (word, particle) pair must own its own 4 bytes inside the
buffer, the word and quad accessors must agree, and the groups the kernels index as `P?? + k` must be consecutive words."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) or shutil.which("g++") is None:
        pytest.skip("needs the CUDA headers and g++")
    out = str(tmp_path_factory.mktemp("layout") / "layout_shim.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + CUDA_INC, "-o", out, os.path.join(HERE, "native", "layout_shim.cpp")], check=True)
    S = C.CDLL(out)
    S.shim_word_offset.restype = C.c_longlong
    S.shim_word_offset.argtypes = [C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_int]
    S.shim_quad_offset.restype = C.c_longlong
    S.shim_quad_offset.argtypes = [C.c_int, C.c_ulonglong, C.c_ulonglong]
    return S


def counts(S):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    S.shim_layout_counts(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def test_counts(shim):
    nfields, nquads, nwords = counts(shim)
    assert nfields == 34            # file_frame/src/particles.rs:93-109 (24 vector words + 7 scalars + flags + bits) + the original index
    assert nquads * 4 == nwords >= nfields and nwords - nfields < 4


@pytest.mark.parametrize("cap", [64, 1000, 1 << 20])
def test_every_word_owns_its_bytes(shim, cap):
    nfields, nquads, nwords = counts(shim)
    rows = np.unique(np.array([0, 1, 2, 31, 32, 33, cap // 2, cap - 2, cap - 1]))
    seen = set()
    for f in range(nfields):
        for i in rows:
            o = shim.shim_word_offset(f, cap, int(i), 0)
            assert o == shim.shim_word_offset(f, cap, int(i), 1)          # float and u32 views agree
            assert 0 <= o < nwords * cap                                   # inside the allocation (NWORDS * cap words)
            assert o not in seen
            seen.add(o)
            q = shim.shim_quad_offset(f >> 2, cap, int(i))
            assert q % 4 == 0 and q == (o // 4) * 4 and o - q == (f & 3)   # the word sits in quad field >> 2 of the same particle
    # consecutive particles of a quad are 16 bytes apart, quads of one particle `cap` elements apart
    assert shim.shim_quad_offset(3, cap, 8) - shim.shim_quad_offset(3, cap, 7) == 4
    assert shim.shim_quad_offset(4, cap, 7) - shim.shim_quad_offset(3, cap, 7) == 4 * cap


def test_groups_the_kernels_index(shim):
    ids = (C.c_int * 14)()
    assert shim.shim_field_ids(ids) == 14
    PX, PFLAGS, PF, PMASS, PVOL, PP0, PP1, PALPHA, PVD, PVB, PBITS, PORIG, PV, PC = list(ids)
    nfields, _, _ = counts(shim)
    used = []
    for first, n in ((PX, 3), (PFLAGS, 1), (PF, 9), (PMASS, 1), (PVOL, 1), (PP0, 1), (PP1, 1), (PALPHA, 1), (PVD, 1), (PVB, 1), (PBITS, 1), (PORIG, 1), (PV, 3), (PC, 9)):
        used += list(range(first, first + n))
    assert sorted(used) == list(range(nfields))                      # the groups tile the 34 words exactly
    assert [PVOL, PP0, PP1, PALPHA, PVD, PVB] == [PMASS + k for k in range(1, 7)]   # G2P carries PMASS + 0..6 as one array
    # what the kernels rely on when they unpack whole quads (svb_kernels.cuh: k_p2g, k_g2p, k_advance)
    assert (PX, PFLAGS) == (0, 3) and PF == 4 and PMASS == 13 and PBITS == 20 and PORIG == 21 and PV == 22 and PC == 25
