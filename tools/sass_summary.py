#!/usr/bin/env python
"""Per-kernel SASS instruction counts of the in-tree library (tcgen05 / TMEM / TMA evidence).

    python tools/sass_summary.py > profiles/sass_summary.txt

Counts, for every kernel in imagescry_b200/lib/obj/*.o, the mnemonics that prove which hardware path
the kernel takes (profiling recipe: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG /
UTMASTG = TMA load / store, UTMAPF = TMA prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier)."""
from __future__ import annotations

import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(REPO, "imagescry_b200", "lib", "obj")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "LDGSTS", "FFMA", "HFMA2", "IDP", "LDL", "STL"]


def main() -> None:
    print("# SASS summary of imagescry_b200/lib/obj/*.o (cuobjdump -sass), sm_100a")
    print("# columns: " + " ".join(KEYS) + "  (UTCHMMA counts exclude the .2CTA form)")
    total = {k: 0 for k in KEYS}
    for obj in sorted(os.listdir(OBJ)):
        if not obj.endswith(".o"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        kernels = re.split(r"\n\s*Function : ", sass)[1:]
        print(f"\n## {obj}")
        for body in kernels:
            name = body.split("\n", 1)[0].strip()
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
            dem = dem.replace("isx::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
            dem = re.sub(r"^void ", "", dem)
            m = re.match(r"([\w:]+(?:<.*?>)?)\(", dem)
            dem = m.group(1) if m else dem.split("(")[0]
            counts = {}
            for k in KEYS:
                if k == "UTCHMMA":
                    counts[k] = len(re.findall(r"\bUTCHMMA(?!\.2CTA)", body))
                else:
                    counts[k] = len(re.findall(r"\b" + re.escape(k), body))
                total[k] += counts[k]
            print(f"{dem:70s} " + " ".join(f"{k}={counts[k]}" for k in KEYS if counts[k]))
    print("\n## whole library")
    print(" ".join(f"{k}={total[k]}" for k in KEYS))


if __name__ == "__main__":
    main()
