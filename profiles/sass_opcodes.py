#!/usr/bin/env python
"""Opcode histogram of every kernel in libruf_b200.so (cuobjdump -sass), with the Blackwell / Hopper-era mnemonics that
prove the async-copy path called out first (B200_PROFILING.md "What proves a Blackwell-native kernel").

    python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "realtime_urdf_filter_b200", "libruf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kern, ops = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        ops[kern][m.group(1) + m.group(2)] += 1
print(f"# {os.path.relpath(lib, ROOT)}: cubin architectures {arch}")
MARK = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "MATCH", "ATOMS", "ATOMG", "REDG", "RED", "PRMT", "VIMNMX3", "BAR", "SHFL",
        "UCGABAR_ARV", "UCGABAR_WAIT", "CGAERRBAR")      # UCGABAR_*: barrier.cluster of the cluster-split raster variant
for k, c in sorted(ops.items(), key=lambda kv: -sum(kv[1].values())):
    tot = sum(c.values())
    print(f"\n== {k}: {tot} SASS instructions")
    marks = {o: n for o, n in c.items() if o.split(".")[0] in MARK}
    print("   async-copy / barrier / atomic / match mnemonics: " + ", ".join(f"{o} x{n}" for o, n in sorted(marks.items())))
    base = collections.Counter()
    for o, n in c.items():
        base[o.split(".")[0]] += n
    print("   " + " ".join(f"{o}:{n}" for o, n in base.most_common(28)))
