"""SASS opcode census of libsfcvit.so: per kernel (template variants merged) the number of tcgen05 MMA (UTC*MMA), TMEM
load / store (LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), legacy tensor (HMMA) and MUFU instructions — the evidence
that the hot kernels run on the Blackwell tensor / TMA path. usage: python tools/sass_census.py [lib.so] > profiles/rN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "space-filling-curves-for-vision-transformers_b200", "lib", "libsfcvit.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU", "variants", "instructions"]
agg = collections.OrderedDict()
cur, k = None, -1
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        base = re.sub(r"<.*", "", re.sub(r"^.*::", "", re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", ""))))
        cur = agg.setdefault(base, collections.Counter())
        cur["variants"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(2)
    cur["instructions"] += 1
    if re.match(r"UTC[A-Z]*MMA", op): cur["UTC*MMA"] += 1
    for key in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU"):
        if op.startswith(key): cur[key] += 1
print(f"SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); template variants of a kernel are summed")
print(f"{'kernel':34s} " + " ".join(f"{c:>9s}" for c in cols))
for name, c in sorted(agg.items(), key=lambda kv: -kv[1]["UTC*MMA"]):
    print(f"{name[:34]:34s} " + " ".join(f"{c[x]:9d}" for x in cols))
tot = collections.Counter()
for c in agg.values():
    tot.update(c)
print(f"{'TOTAL':34s} " + " ".join(f"{tot[x]:9d}" for x in cols))
# mma.sync is allowed in ONE place: the small-shape attention for head dimensions other than 64 (csrc/attention_generic.cu,
# 64-token tiles too small for a tcgen05 plan — DESIGN.md §2 K4g). Every BASELINE.json configuration has head_dim 64 and
# never launches it.
legacy = {n: c["HMMA"] for n, c in agg.items() if c["HMMA"]}
print("kernels with legacy HMMA:", legacy)
assert all("attn_mma_" in n for n in legacy), "legacy mma.sync / wmma instructions outside attention_generic.cu"
