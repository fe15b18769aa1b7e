"""Opcode evidence from the built library (no GPU needed): cuobjdump -sass cvmatrix_b200/libcvmx.so -> profiles/r02_sass_evidence.txt
(DMMA = FP64 tensor cores, UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = TMA bulk copy, SYNCS = mbarrier,
LDGSTS = cp.async, USETMAXREG = setmaxnreg)."""
import collections
import os
import re
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "cvmatrix_b200", "libcvmx.so")], capture_output=True, text=True).stdout
counts, per, cur = collections.Counter(), collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur:
        counts[m.group(1)] += 1
        per[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"] + list(per.keys()), capture_output=True, text=True).stdout.splitlines()
keys = ["DMMA.8x8x4", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "UBLKCP.S.G", "USETMAXREG.TRY_ALLOC.CTAPOOL", "USETMAXREG.DEALLOC.CTAPOOL"]
keys += sorted(k for k in counts if k.split(".")[0] in ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "LDGSTS", "SYNCS", "ARRIVES", "SHFL"))
with open(os.path.join(ROOT, "profiles", "r02_sass_evidence.txt"), "w") as f:
    f.write("# SASS evidence (cuobjdump -sass cvmatrix_b200/libcvmx.so, sm_100a only; regenerate: python tools/sass_evidence.py)\n\n")
    f.write("## opcode counts over the whole library\n")
    for k in keys:
        if counts[k]:
            f.write(f"{counts[k]:7d} {k}\n")
    f.write("\n## per kernel (float64 instantiations; LOO flag variants folded): DMMA (FP64 tensor) | UBLKCP (TMA bulk copy) | SYNCS (mbarrier) | "
            "LDGSTS (cp.async) | DADD | DMUL | DFMA | SHFL\n")
    seen = set()
    for name, c in zip(names, per.values()):
        short = re.sub(r"k_loo_folds<(\w+), \d+>", r"k_loo_folds<\1, *>", name)
        if ("float" in short and "k_gram_tc" not in short) or short in seen:
            continue
        seen.add(short)
        g = lambda p: sum(v for o, v in c.items() if o.startswith(p))  # noqa: E731
        tc = f" | UTCHMMA {g('UTCHMMA')} | LDTM {g('LDTM')} | UTCBAR {g('UTCBAR')}" if g("UTCHMMA") else ""
        f.write(f"{short[:90]}{tc} | DMMA {g('DMMA')} | UBLKCP {g('UBLKCP')} | SYNCS {g('SYNCS')} | LDGSTS {g('LDGSTS')} | DADD {c['DADD']} | "
                f"DMUL {c['DMUL']} | DFMA {c['DFMA']} | SHFL {g('SHFL')}\n")
print(open(os.path.join(ROOT, "profiles", "r02_sass_evidence.txt")).read())
