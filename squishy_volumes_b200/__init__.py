"""squishy_volumes_b200 — B200-native MPM substep behind the reference's back-end boundary.

Only the hot path of Algebraic-UG/squishy_volumes is here: the CUDA kernels + C-ABI
(`csrc/`, `include/svb200.h`) and the host-side mirror of `CpuState` / `GpuState`
(`state.B200State`).  Importing the package does not load the CUDA library; `abi.load()` does and
fails loudly when it is missing.
"""
from .types import (ColliderTopology, FatalError, FrameInput, GridNodes, Harness, InputConsts, IoState, Keyframe,
                    ParticleFlags, Particles, RunParameters, SimulationError)

__all__ = ["ColliderTopology", "FatalError", "FrameInput", "GridNodes", "Harness", "InputConsts", "IoState",
           "Keyframe", "ParticleFlags", "Particles", "RunParameters", "SimulationError"]
