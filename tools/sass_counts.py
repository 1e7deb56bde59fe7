"""profiles/rNN_sass_counts.md: SASS instruction counts per kernel of libmunk_b200.so
(cuobjdump -sass), the evidence that the FP64 path is DMMA + TMA + mbarriers.
usage: python tools/sass_counts.py profiles/r02_sass_counts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "munk_b200", "lib", "libmunk_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter(m.group(2) for m in re.finditer(
        r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", f, re.M))
    cnt = lambda *pre: sum(v for k, v in c.items() if k.startswith(pre))  # noqa: E731
    rows.append([name, sum(c.values()), cnt("DMMA"), cnt("UTMALDG"), cnt("SYNCS"), cnt("DFMA", "DMUL", "DADD"),
                 cnt("LDS"), cnt("LDG"), cnt("STG"), cnt("SHFL"), cnt("ATOM", "RED"),
                 " ".join(sorted(k for k in c if k.startswith(("DMMA", "UTMALDG", "SYNCS"))))])
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.split("\n")
out = ["# SASS instruction counts per kernel, libmunk_b200.so, round 2", "",
       "`cuobjdump -sass munk_b200/lib/libmunk_b200.so`, counted per function; cubin architectures: %s." % ", ".join(arch),
       "FP64 tensor work is `DMMA.8x8x4` (tcgen05 has no f64 kind, so no `UTCMMA`/`LDTM` can appear); the GEMM's",
       "operands are staged by TMA (`UTMALDG.2D`) and completed through mbarriers (`SYNCS.*`).", "",
       "| kernel | instr | DMMA | UTMALDG | SYNCS | DFMA/DMUL/DADD | LDS | LDG | STG | SHFL | ATOM/RED | tensor / TMA / mbarrier mnemonics |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r, n in sorted(zip(rows, names), key=lambda x: -x[0][2]):
    n = n.replace("(anonymous namespace)::", "").replace("munk::detail::", "").replace("munk::", "")
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    out.append("| `%s` | %s | %s |" % (n[:80], " | ".join(str(x) for x in r[1:11]), r[11]))
open(sys.argv[1] if len(sys.argv) > 1 else "/dev/stdout", "w").write("\n".join(out) + "\n")
