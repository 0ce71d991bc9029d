"""Static instruction mix of the fused operator's main loop (two x-planes per trip), from `cuobjdump -sass` of the built
library: how many issue slots are FP64 and how many are overhead.  No GPU needed.
    python scripts/sass_loop_mix.py [mangled-name-fragment]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "easysimp.jl_b200" / "libeasysimp_b200.so"
frag = sys.argv[1] if len(sys.argv) > 1 else "k_matvec_fusedILi16ELi16ELb0ELi2ELi0E"
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if "Function :" in l and frag in l)
end = next(i for i in range(start + 1, len(sass)) if "Function :" in sass[i] or i == len(sass) - 1)
ins = []
for l in sass[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
back = []
for a, t in ins:
    m = re.search(r"BRA\s+(?:.*\s)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        back.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
span, lo, hi = max(back)
loop = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for a, t in ins if lo <= a <= hi]
c = collections.Counter(loop)
fp64 = sum(v for k, v in c.items() if k.startswith(("DFMA", "DADD", "DMUL")))
mem = sum(v for k, v in c.items() if k.startswith(("LDS", "STS", "LDG", "STG", "LDC", "LDGSTS")))
print(f"kernel {frag}: {len(ins)} instructions, main loop {len(loop)} (0x{lo:x}..0x{hi:x})")
print(f"  FP64 {fp64} ({100 * fp64 / len(loop):.0f} %)  memory {mem}  other {len(loop) - fp64 - mem}")
print(f"  FP64 pipe cycles per trip per warp at 16 lanes/clk/SMSP: {2 * fp64}; issue slots: {len(loop)}")
try:
    for k, v in c.most_common():
        print(f"  {v:4d} {k}")
except BrokenPipeError:
    pass
