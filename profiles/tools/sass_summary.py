"""Mnemonic histogram + evidence lines of the shipped g + jac_g kernel (cuobjdump -sass of the object that is linked into
libmpx.so).  Usage: python profiles/tools/sass_summary.py > profiles/r02/sass_gjac2_d15.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GEN = os.path.join(ROOT, "mpopt_b200", "csrc", "gen")
KEY = next(f[8:-3] for f in sorted(os.listdir(GEN)) if f.endswith(".cu")
           and open(os.path.join(GEN, f)).readline().strip() == "// problem: synthetic_6_3")
OBJ = os.path.join(ROOT, "mpopt_b200", "build", f"mpx_aot_{KEY}.o")
NAME = f"_Z16mpx_gjac2_kernelI24MpxPh_{KEY}_0Lb1ELi15EEv12MpxPhaseArgs"
txt = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
start = txt.index("Function : " + NAME)
nxt = txt.find("Function : ", start + 20)
body = txt[start: nxt if nxt > 0 else len(txt)]
lines = [l for l in body.splitlines() if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
ops = collections.Counter()
for l in lines:
    m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
    if m:
        ops[m.group(2).split(".")[0]] += 1
print(f"cuobjdump -sass mpopt_b200/build/mpx_aot_{KEY}.o   (the object linked into the shipped libmpx.so)")
print("kernel:", NAME, " = mpx_gjac2_kernel<synthetic_6_3, JAC = true, DEG = 15>")
print("instructions:", len(lines))
print("\nmnemonic histogram:")
for k, v in ops.most_common():
    print(f"  {k:12s} {v}")
print("\nevidence (bulk copy engine = TMA 1-D, mbarrier, programmatic dependent launch, fp64 FMA; no tensor-core op: "
      "there is no contraction on this path):")
for key in ("UBLKCP", "SYNCS", "ACQBULK", "DEPBAR", "DFMA", "DMUL", "DADD", "STS", "LDS", "LDG", "STG", "HMMA", "UTC", "UTMA"):
    hits = [l.strip() for l in lines if re.search(r"\b" + key, l)]
    print(f"  {key}: {len(hits)}")
    for h in hits[:2]:
        print("      " + re.sub(r"\s+", " ", h)[:140])
