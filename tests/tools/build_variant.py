"""Development: build squishy_volumes_b200/lib/variants/NAME.so with extra -D flags (A/B runs: tests/tools/ab.sh).
usage: python tests/tools/build_variant.py NAME -DSVB_P2G_CTAS_PER_SM=6 -DP2G_PREFETCH=0 ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from squishy_volumes_b200 import abi  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(abi.LIB_PATH), "variants", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
cmd = [nvcc] + abi.NVCC_FLAGS + flags + ["-o", out, os.path.join(abi.CSRC, "svb200.cu"), os.path.join(abi.CSRC, "svb_files.cpp"), "-lnccl"]
subprocess.run(cmd, check=True, cwd=abi.CSRC)
print(out)
