"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md "What proves a Blackwell-native kernel"):
    python profiles/sass_summary.py > profiles/r2_sass_summary.txt
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA unit, non-tensor bulk copy), UTCBAR = tcgen05.commit,
REDG = red.global, LDGMC = multimem.ld_reduce (NVLS in-switch reduction), STG...MC / multimem.st, USETMAXREG = setmaxnreg, SYNCS = mbarrier."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "nerf2mesh_b200", "libn2m_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UTCHMMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"), ("UTMALDG/STG", r"\bUTMA(LDG|STG)"),
                                ("UTCBAR", r"\bUTCBAR"), ("SYNCS(mbarrier)", r"\bSYNCS"), ("REDG.F32x4", r"\bRED\w*\.E\.ADD\.F32x4"), ("RED(other)", r"\bRED\w*\.E"),
                                ("ATOMG", r"\bATOMG"), ("LDGMC(multimem.ld_reduce)", r"\bLDGMC"), ("multimem.st", r"\bST\w*\.MC|\bSTGMC|\bREDGMC"),
                                ("USETMAXREG", r"\bUSETMAXREG"), ("HMMA(legacy)", r"\bHMMA")])
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = re.search(r"(k_[a-z0-9_]+)", m.group(1))
        cur = name.group(1) if name else m.group(1)[:60]
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            counts[cur][k] += 1
print("kernel".ljust(28) + "".join(k.rjust(12)[:12] + " " for k in pats))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(k.ljust(28) + "".join(str(c[p]).rjust(12) + " " for p in pats))
    tot.update(c)
print("TOTAL".ljust(28) + "".join(str(tot[p]).rjust(12) + " " for p in pats))
