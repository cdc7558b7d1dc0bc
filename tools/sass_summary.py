#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-specific instructions in libfava_b200.so (cuobjdump -sass): TMA bulk
copies (UBLKCP), TMA tensor loads (UTMALDG), mbarrier transactions (SYNCS), cp.async (LDGSTS), fp64 math (DFMA/DADD/DMUL),
spills (LDL/STL).  Writes a markdown table (default profiles/r02_sass_summary.md)."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PATTERNS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDL", "STL", "BAR", "UTC"]


def main():
    lib = ROOT / "fava_b200" / "lib" / "libfava_b200.so"
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r02_sass_summary.md"
    text = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    rows = []
    name = None
    counts = Counter()
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, counts))
            name, counts = m.group(1), Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for p in PATTERNS:
                if op.startswith(p):
                    counts[p] += 1
            counts["total"] += 1
    if name:
        rows.append((name, counts))
    demangled = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS instruction counts per kernel of libfava_b200.so (sm_100a), `tools/sass_summary.py`", "",
             "UBLKCP = cp.async.bulk (TMA bulk copy), UTMALDG = cp.async.bulk.tensor (TMA tensor tile load), SYNCS = mbarrier "
             "arrive/expect_tx/try_wait, LDGSTS = cp.async; no UTC*MMA anywhere: nothing on this path is a dense contraction.", "",
             "| kernel | total | " + " | ".join(PATTERNS) + " |", "|---|---|" + "---|" * len(PATTERNS)]
    for (mangled, c), dm in sorted(zip(rows, demangled), key=lambda t: t[1]):
        short = re.sub(r"\(.*", "", dm).replace("void fava::", "")
        lines.append(f"| `{short}` | {c['total']} | " + " | ".join(str(c[p]) if c[p] else "" for p in PATTERNS) + " |")
    out.write_text("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main()
