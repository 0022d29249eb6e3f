"""The C-ABI library loads and exports every symbol include/svb200.h declares (no compute calls:
this runs without a GPU), and the product path fails loudly when no device / library is present."""
import ctypes as C
import os
import subprocess

import pytest

from squishy_volumes_b200 import abi, cstructs as cs


@pytest.fixture(scope="module")
def lib_path():
    return abi.build()


def test_exports_every_declared_symbol(lib_path):
    declared = abi.declared_symbols()
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    L = abi.load()
    for s in declared:
        assert hasattr(L, s)


def test_struct_layouts_match_header():
    # plain-C layouts (x86-64): these sizes are what a cgo / Rust #[repr(C)] binding would see
    assert C.sizeof(cs.SvbConsts) == 44
    assert C.sizeof(cs.SvbParticles) == 8 + 15 * 8
    assert C.sizeof(cs.SvbKeyframe) == 16 + 5 * 8 or C.sizeof(cs.SvbKeyframe) == 12 + 4 + 5 * 8
    assert C.sizeof(cs.SvbGrid) == 8 + 5 * 8


def test_built_for_sm_100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_oracle_in_product():
    """The shipped package never imports or links the oracle."""
    root = os.path.dirname(abi.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "svo_" not in text and "liboracle" not in text, f


def test_fails_loudly_without_device(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from squishy_volumes_b200 import scenes
    from squishy_volumes_b200.state import B200State
    from squishy_volumes_b200.types import FatalError
    sc = scenes.elastic_cube(side=4, h=0.1)
    with pytest.raises(FatalError):
        B200State.from_io_state(sc.io_state, sc.frame_input)


def test_headers_are_plain_c(tmp_path):
    """include/*.h is the drop-in boundary: plain C (no C++ or torch types), usable from cgo / Rust bindgen / ctypes;
    the ctypes mirrors of the file-layer structs have the header's sizes."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text(f'#include "{root}/include/svb200.h"\n#include "{root}/include/svb_files.h"\n'
                   "#include <stdio.h>\nint main(void){printf(\"%zu %zu %zu\\n\", sizeof(SvbfObjectDesc), sizeof(SvbfParticlesInput), sizeof(SvbfColliderInput));return 0;}\n")
    exe = tmp_path / "hdr"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    from squishy_volumes_b200 import files
    assert sizes == [C.sizeof(files.SvbfObjectDesc), C.sizeof(files.SvbfParticlesInput), C.sizeof(files.SvbfColliderInput)]
