"""SASS opcode summary of the product library, per kernel:  python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pl-nerf_b200", "libplnerf_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMAPF", "SYNCS", "REDG", "CCTL", "LDL", "STL", "F2FP", "FADD2", "FMNMX", "MUFU",
        "BAR", "DADD", "DMUL", "HMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m:
        op = m.group(1)
        funcs[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                funcs[cur][k] += 1
        if op.startswith("REDG") and "F32x4" in line:
            funcs[cur]["REDG.F32x4"] += 1
        if op.startswith("STG") and ".EF" in line:
            funcs[cur]["STG.EF(streaming)"] += 1
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode summary of the product library (cuobjdump -sass pl-nerf_b200/libplnerf_b200.so), per kernel; made by profiles/sass_summary.py")
print("# tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld/st = LDTM/STTM, cp.async.bulk (1-D TMA) = UBLKCP, mbarrier = SYNCS,")
print("# red.global = REDG (.F32x4: four columns per reduction), LDL/STL = local-memory (spill) accesses; no mma.sync (HMMA) anywhere\n")
for name, c in zip(names, funcs.values()):
    name = re.sub(r"plnerf::|\(anonymous namespace\)::", "", name)
    extra = "  ".join(f"{k}={c[k]}" for k in list(KEYS) + ["REDG.F32x4", "STG.EF(streaming)"] if c[k])
    print(f"{name[:72]:72s} total={c['total']:6d}  {extra}")
