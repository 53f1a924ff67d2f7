"""Per-kernel SASS opcode counts of libdagma_b200.so (cuobjdump -sass): the instructions that prove which hardware
paths a kernel uses -- DMMA (FP64 tensor pipe), UTMALDG (TMA loads), LDTM / STTM / UTCBAR-family (tensor memory),
STL / LDL (local-memory spills), plus DFMA, SHFL, BAR, SYNCS (mbarrier) for context.

    python scripts/sass_opcodes.py [lib.so] > profiles/sass_opcodes_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "midagma_b200", "lib", "libdagma_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
COLS = ["DMMA", "UTMALDG", "LDTM", "STTM", "UTC", "STL", "LDL", "DFMA", "DMUL", "DADD", "SHFL", "BAR", "SYNCS", "LDS", "STS",
        "LDG", "STG", "RED", "MUFU", "total"]
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for c in COLS[:-1]:
            if op == c or (c == "UTC" and op.startswith("UTC")) or (c == "UTMALDG" and op.startswith("UTMALDG")):
                counts[kern][c] += 1


def demangle(name):
    try:
        full = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        return name
    if not full.endswith(")"):
        return full
    depth = 0                                   # cut the trailing parameter list (template arguments may hold parentheses)
    for i in range(len(full) - 1, -1, -1):
        depth += full[i] == ")"
        depth -= full[i] == "("
        if depth == 0:
            return full[:i]
    return full


print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass; static counts, not executed counts)")
print("# STL / LDL = local-memory stores / loads (register spills or stack arrays); UTC = tcgen05 alloc / dealloc / etc.")
print(f"{'kernel':64s} " + " ".join(f"{c:>7s}" for c in COLS))
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["DMMA"]):
    name = demangle(k).replace("dagma::", "").replace("(bool)1", "true").replace("(bool)0", "false")
    print(f"{name[:64]:64s} " + " ".join(f"{c[col]:7d}" for col in COLS))
