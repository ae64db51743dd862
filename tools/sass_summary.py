#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell data path (B200_PROFILING.md):
UBLKCP (TMA bulk copy), UTCHMMA/UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), SHFL (warp-shuffle butterflies), LDG.E.128 / STG.E.128 (16-byte global accesses), FFMA.

    python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt

Runs on the build host (cuobjdump only, no GPU)."""
from __future__ import annotations

import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "spatial_audio_framework_b200" / "libsafconv_b200.so"
MNEMONICS = ["UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "SYNCS", "SHFL", "LDG.E.128", "STG.E.128",
             "LDG.E.64", "LDS.128", "LDS.64", "FFMA", "HFMA2", "BAR.SYNC"]


def main() -> int:
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                           capture_output=True, text=True, check=True).stdout.splitlines()
    per: "OrderedDict[str, dict]" = OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        if "Function :" in line:
            cur = next(it)
            cur = re.sub(r"^void ", "", cur)
            cur = re.sub(r"\((?:[A-Za-z]+Args|.*)\)$", "", cur)
            per[cur] = {m: 0 for m in MNEMONICS}
            per[cur]["instr"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["instr"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                per[cur][mn] += 1
    cols = ["instr"] + MNEMONICS
    w = max(len(k) for k in per) + 1
    print(f"# {lib.name}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); static instruction counts")
    print("kernel".ljust(w) + " ".join(c.rjust(max(6, len(c))) for c in cols))
    for k in sorted(per):
        print(k.ljust(w) + " ".join(str(per[k][c]).rjust(max(6, len(c))) for c in cols))
    tot = {c: sum(v[c] for v in per.values()) for c in cols}
    print("TOTAL".ljust(w) + " ".join(str(tot[c]).rjust(max(6, len(c))) for c in cols))
    return 0


if __name__ == "__main__":
    sys.exit(main())
