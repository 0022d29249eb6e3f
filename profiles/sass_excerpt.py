"""SASS evidence of the built library (no GPU needed): per kernel the instruction count and the mnemonics that matter on this path —
packed fp32 (FFMA2 / FMUL2 / FADD2), vector reductions to HBM (REDG.E.ADD.F32x4 / RED), shared-memory atomics (ATOMS), warp votes
(MATCH), 128-bit shared loads (LDS.128) — plus the first lines of P2G's walk loop.  No tensor-core or TMA mnemonics are expected
(UTCMMA / UTMALDG / LDTM): nothing on the path is a dense contraction.
    python profiles/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "squishy_volumes_b200/lib/libsvb200.so"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[name] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if m and name:
        kernels[name].append(m.group(1).strip())
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {LIB}: architectures {arch}")
keys = ("FFMA2", "FMUL2", "FADD2", "FFMA", "REDG", "RED.", "ATOMS", "ATOMG", "MATCH", "LDS.128", "LDG", "STG", "MUFU", "UTCMMA", "UTMALDG", "LDTM", "HMMA")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in keys))
for k, ins in kernels.items():
    if not ins:
        continue
    cnt = [sum(1 for i in ins if re.search(r"(^|\s)" + re.escape(key), i)) for key in keys]
    print(f"{k[-58:]:58s} {len(ins):6d} " + " ".join(f"{c:7d}" for c in cnt))
p2g = next((v for k, v in kernels.items() if "k_p2g<false>" in k or "k_p2g<(bool)0>" in k), None)
if p2g:
    first = next((i for i, s in enumerate(p2g) if "FFMA2" in s), 0)
    print("\n# k_p2g<false>: the walk (lane = stencil node), first packed-fp32 group")
    for s in p2g[max(0, first - 6):first + 30]:
        print("   ", s)
    red = next((i for i, s in enumerate(p2g) if "RED" in s), None)
    if red is not None:
        print("# ... tile -> HBM")
        for s in p2g[red - 2:red + 2]:
            print("   ", s)
