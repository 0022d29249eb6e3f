"""ctypes binding of the C ABI declared in include/svb200.h (the drop-in seam a Rust
`ComputeState::B200` arm would bind through `extern "C"`).

`load()` fails loudly when `lib/libsvb200.so` is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

from . import cstructs as cs

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVB200_LIB") or os.path.join(_HERE, "lib", "libsvb200.so")  # SVB200_LIB: developer override for A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "svb200.h")
FILES_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "svb_files.h")
CSRC = os.path.join(_HERE, "csrc")
REPLAY_PATH = os.path.join(_HERE, "lib", "svb_replay")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

_lib = None


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".cpp"))] + [HEADER_PATH, FILES_HEADER_PATH]


def build(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> squishy_volumes_b200/lib/libsvb200.so (in-tree)."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in sources())
    if stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "svb200.cu"), os.path.join(CSRC, "svb_files.cpp"), "-lnccl"]
        subprocess.run(cmd, check=True, cwd=CSRC)
        # the compute thread's frame loop over the two C ABIs (csrc/svb_replay.cpp), linked against the library next to it
        gxx = os.environ.get("CXX", "g++")
        subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", os.path.join(CSRC, "svb_replay.cpp"), "-o", REPLAY_PATH, "-L" + os.path.dirname(LIB_PATH), "-lsvb200",
                        "-Wl,-rpath,$ORIGIN"], check=True, cwd=CSRC)
    return LIB_PATH


def declared_symbols():
    """Every function name include/svb200.h and include/svb_files.h declare."""
    names = set()
    for path in (HEADER_PATH, FILES_HEADER_PATH):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(svbf?_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


class LibraryMissing(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback for the MPM substep)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.svb_available_devices.restype = C.c_int32
    L.svb_available_devices.argtypes = [C.c_char_p, C.c_size_t]
    L.svb_create.restype = C.c_int32
    L.svb_create.argtypes = [C.POINTER(cs.SvbConsts), C.POINTER(cs.SvbParticles), C.c_double, C.c_int32, C.POINTER(vp)]
    L.svb_create_multi.restype = C.c_int32
    L.svb_create_multi.argtypes = [C.POINTER(cs.SvbConsts), C.POINTER(cs.SvbParticles), C.c_double, C.POINTER(C.c_int32), C.c_int32, C.POINTER(vp)]
    L.svb_destroy.restype = None
    L.svb_destroy.argtypes = [vp]
    L.svb_upload.restype = C.c_int32
    L.svb_upload.argtypes = [vp, C.POINTER(cs.SvbParticles), C.c_double]
    L.svb_set_topology.restype = C.c_int32
    L.svb_set_topology.argtypes = [vp, C.c_uint32, cs.c_u32p, cs.c_u32p, cs.c_u32p]
    L.svb_set_keyframes.restype = C.c_int32
    L.svb_set_keyframes.argtypes = [vp, C.c_uint64, C.POINTER(cs.SvbKeyframe), C.POINTER(cs.SvbKeyframe)]
    L.svb_advance.restype = C.c_int32
    L.svb_advance.argtypes = [vp, C.c_double, C.c_float, C.c_int32, vp, vp, vp]
    L.svb_download.restype = C.c_int32
    L.svb_download.argtypes = [vp, C.POINTER(cs.SvbParticles)]
    L.svb_grid_count.restype = C.c_int64
    L.svb_grid_count.argtypes = [vp]
    L.svb_download_grid.restype = C.c_int32
    L.svb_download_grid.argtypes = [vp, C.POINTER(cs.SvbGrid)]
    L.svb_time.restype = C.c_double
    L.svb_time.argtypes = [vp]
    L.svb_substeps.restype = C.c_uint64
    L.svb_substeps.argtypes = [vp]
    L.svb_allowed_time_step.restype = C.c_float
    L.svb_allowed_time_step.argtypes = [vp]
    L.svb_status.restype = C.c_uint32
    L.svb_status.argtypes = [vp]
    L.svb_last_error.restype = C.c_char_p
    L.svb_last_error.argtypes = [vp]
    L.svb_kernel_launches.restype = C.c_uint64
    L.svb_kernel_launches.argtypes = [vp]
    L.svb_last_advance_ms.restype = C.c_float
    L.svb_last_advance_ms.argtypes = [vp]
    L.svb_binning.restype = C.c_int32
    L.svb_binning.argtypes = [vp, cs.c_u32p, cs.c_i32p]
    L.svb_active_block_count.restype = C.c_int64
    L.svb_active_block_count.argtypes = [vp]
    L.svb_active_blocks.restype = C.c_int32
    L.svb_active_blocks.argtypes = [vp, cs.c_i32p, cs.c_u32p]
    L.svb_stage_times.restype = C.c_int32
    L.svb_stage_times.argtypes = [vp, C.POINTER(C.c_char_p), cs.c_f32p, C.c_int32]
    L.svb_exchange_waits.restype = C.c_int32
    L.svb_exchange_waits.argtypes = [vp, C.POINTER(C.c_double), C.c_int32, C.c_int32]
    L.svb_enable_stage_timing.restype = None
    L.svb_enable_stage_timing.argtypes = [vp, C.c_int32]
    L.svb_set_option.restype = None
    L.svb_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.svb_node_ids_to_murmur.restype = C.c_int32
    L.svb_node_ids_to_murmur.argtypes = [C.c_int32, cs.c_i32p, cs.c_u32p, C.c_uint64, cs.c_u32p, cs.c_u32p]
    L.svb_snapshot.restype = C.c_int32
    L.svb_snapshot.argtypes = [vp]
    L.svb_restore.restype = C.c_int32
    L.svb_restore.argtypes = [vp]
    L.svb_particle_count.restype = C.c_uint64
    L.svb_particle_count.argtypes = [vp]
    L.svb_comm_unique_id.restype = C.c_int32
    L.svb_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.svb_comm_init.restype = C.c_int32
    L.svb_comm_init.argtypes = [vp, C.POINTER(C.c_uint8), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64]
    L.svb_set_original_indices.restype = C.c_int32
    L.svb_set_original_indices.argtypes = [vp, cs.c_u32p, C.c_uint64]
    L.svb_slab_histogram.restype = C.c_int32
    L.svb_slab_histogram.argtypes = [vp, C.c_int32, C.c_uint32, C.POINTER(C.c_uint64)]
    L.svb_slab_rebalance.restype = C.c_int32
    L.svb_slab_rebalance.argtypes = [vp, C.c_int32, C.c_int32]
    L.svb_download_resident.restype = C.c_int32
    L.svb_download_resident.argtypes = [vp, C.POINTER(cs.SvbParticles), C.POINTER(C.c_uint64)]
    _lib = L
    return L
