"""Per-kernel SASS evidence (no GPU needed: cuobjdump on the built library): instruction counts and the mnemonics that show
what the kernel is made of — UBLKCP (1-D bulk async copy = TMA engine), SYNCS (mbarrier), STG.E.EF.128 / STG.E.128
(128-bit streaming / vector stores), LDS / STS, SHFL, BAR.SYNC, F2I.* .RN-style conversions (cvt.rni).  Writes
profiles/<tag>_sass_excerpt.txt.

    python tools/sass_excerpt.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("STG.E.EF.128", r"STG\.E\.EF\.128"), ("STG.E.128", r"STG\.E\.128"),
       ("STG (all)", r"\bSTG"), ("LDG (all)", r"\bLDG"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("SHFL", r"\bSHFL"),
       ("BAR.SYNC", r"BAR\.SYNC"), ("WARPSYNC", r"\bWARPSYNC"), ("F2I", r"\bF2I"), ("DMUL/DFMA", r"\bD(MUL|FMA|ADD)"),
       ("FFMA", r"\bFFMA"), ("FMUL", r"\bFMUL"), ("FADD", r"\bFADD"), ("MUFU", r"\bMUFU")]


def main(tag):
    so = os.path.join(ROOT, "glava_b200", "libglava_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\n\s+REG:(\d+)", res))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels[cur].append(line)
    demangled = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    out = [f"# SASS excerpt of glava_b200/libglava_b200.so ({tag}): cuobjdump -sass, sm_100a.  One block per kernel.", ""]
    for (name, lines), dem in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", dem).replace("void glb::", "").replace("glb::", "")
        out.append(f"## {short}   [{len(lines)} instructions, {regs.get(name, '?')} registers]")
        counts = [(k, sum(1 for l in lines if re.search(p, l))) for k, p in PAT]
        out.append("   " + "  ".join(f"{k}={v}" for k, v in counts if v))
        shown = set()
        for key in ("UBLKCP", "SYNCS", "STG.E.EF.128", "STG.E.128", "SHFL"):
            for l in lines:
                if re.search(dict(PAT)[key], l) and key not in shown:
                    out.append("   e.g. " + re.sub(r"\s+", " ", l.split("*/", 1)[1].split("/*")[0]).strip()); shown.add(key)
        out.append("")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.txt")
    open(path, "w").write("\n".join(out))
    print("wrote", path, len(kernels), "kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
