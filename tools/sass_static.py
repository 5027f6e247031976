#!/usr/bin/env python
"""Static SASS opcode histogram of every kernel in the shipped library (cuobjdump -sass), with the Blackwell-specific mnemonics
checked: this path is integer-multiply bound, so IMAD.WIDE dominates and no tensor-core / TMA mnemonic (UTC*MMA, LDTM, UTMA*)
appears.    python tools/sass_static.py [kernel substring] > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "aeonflux_b200", "csrc", "libaeonflux_b200.so")
want = sys.argv[1] if len(sys.argv) > 1 else ""
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        per[cur][m.group(1).rstrip(";")] += 1
print("library:", os.path.relpath(so, ROOT), " arch:", ", ".join(sorted(set(re.findall(r"arch = (\S+)", txt)))))
for name, ops in per.items():
    if want not in name:
        continue
    tot = sum(ops.values())
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    special = {k: v for k, v in ops.items() if k.startswith(("UTC", "LDTM", "STTM", "UTMA", "HMMA", "IMMA", "QMMA", "WGMMA", "UBLKCP"))}
    print("\n== %s: %d instructions (%.0f KB), local-memory LDL %d / STL %d, tensor-core or TMA mnemonics: %s"
          % (short, tot, tot * 16 / 1024, sum(v for k, v in ops.items() if k.startswith("LDL")), sum(v for k, v in ops.items() if k.startswith("STL")), special or "none"))
    for op, n in ops.most_common(16 if not want else 40):
        print("   %-22s %6d  %5.1f %%" % (op, n, 100.0 * n / tot))
