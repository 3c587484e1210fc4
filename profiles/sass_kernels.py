"""Static SASS opcode histogram per kernel of a built library (no GPU needed):
    python profiles/sass_kernels.py goi-hyperplane_b200/lib/libgoi_raster.so [regex] > profiles/r02_sass_hist.txt
Shows, per kernel, the instruction count and the opcodes that prove which hardware paths are used
(HMMA = warp-level mma.sync, UTCHMMA / LDTM / UTCBAR = tcgen05 + tensor memory, REDG = vector reductions,
LDGSTS = cp.async, FFMA2 = packed fp32)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(r"k_composite|k_mask|k_preprocess|k_emit|k_tile|k_pad|k_semloss")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
KEY = ("HMMA", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "REDG", "RED", "ATOMG", "LDGSTS", "FFMA2", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "VOTE")
for name, pretty in sorted(zip(hist, demangle), key=lambda kv: kv[1]):
    if not pat.search(pretty):
        continue
    h = hist[name]
    short = re.sub(r"\(.*", "", pretty.replace("(anonymous namespace)::", "")).replace("void ", "")
    print(f"{short:48s} {sum(h.values()):6d} instr | " + " ".join(f"{k}={h[k]}" for k in KEY if h[k]))
