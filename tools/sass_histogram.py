#!/usr/bin/env python
"""Opcode histogram of the built libv3d.so, per kernel (the .so itself is git-ignored, so this listing is the
committed evidence of what the binary contains).

    python tools/sass_histogram.py > profiles/<round>_sass_opcodes.txt

Lists, for every kernel, the instruction count and the Blackwell-specific / hot-loop opcodes the design relies on:
UBLKCP (cp.async.bulk, TMA bulk copies), SYNCS (mbarrier), STAS (st.async into distributed shared memory),
UCGABAR / CGA barriers (thread-block clusters), VIMNMX / VIMNMX3 / VIADDMNMX (DPX packed min/add), CREDUX (warp
reduction), RED (L2 reductions), LDGSTS (cp.async), SHFL, PRMT, LDS/STS.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "video-3d-pipeline_b200" / "video_3d_pipeline" / "libv3d.so"
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "STAS", "UCGABAR", "CGAERRBAR", "VIMNMX3", "VIMNMX", "VIADDMNMX", "VIADD", "CREDUX",
         "REDUX", "RED", "ATOM", "LDGSTS", "SHFL", "PRMT", "LDS", "STS", "LDG", "STG", "BAR", "FFMA", "FADD", "MUFU", "IMAD", "LOP3",
         "SEL", "UTCMMA", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(SO)], check=True, capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"^void ", "", name).split("(")[0]
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("VIMNMX", "VIMNMX3", "VIADDMNMX", "RED", "SYNCS", "UBLKCP", "STAS", "CREDUX"):
                cur["~" + full] += 1
    print(f"# cuobjdump -sass {SO.relative_to(ROOT)}  (sm_100a); opcode counts per kernel (static instructions)")
    tot = collections.Counter()
    for name, c in kernels.items():
        tot.update(c)
        watched = "  ".join(f"{k}={c[k]}" for k in WATCH if c[k])
        print(f"\n{name}\n  instructions={c['_total']}  {watched}")
        variants = "  ".join(f"{k[1:]}={v}" for k, v in sorted(c.items()) if k.startswith("~"))
        if variants:
            print(f"  variants: {variants}")
    print("\n# whole library")
    print("  " + "  ".join(f"{k}={tot[k]}" for k in WATCH if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
