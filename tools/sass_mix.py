#!/usr/bin/env python
"""Executed-instruction mix of the kernels in an `ncu --page source --csv --print-source sass` dump: per kernel, executed warp
instructions by opcode (top 25) and the share of local-memory instructions (LDL/STL), with the hottest of those by address -- the
evidence for "no spill traffic inside the window body".  Usage: python tools/sass_mix.py src.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    kernel, hdr = None, None
    per = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if len(r) == 1 or (r[0].startswith("Kernel") and len(r) <= 3):
            kernel = r[-1]; hdr = None
            continue
        if hdr is None and ("Source" in r or "# Source" in r or any(c.strip() in ("Source", "SASS") for c in r)):
            hdr = [c.strip() for c in r]
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        src = d.get("Source") or d.get("SASS") or ""
        ex = d.get("Warp Instructions Executed") or d.get("Instructions Executed") or d.get("# Instructions Executed") or "0"
        try:
            n = int(float(ex.replace(",", "")))
        except ValueError:
            continue
        toks = src.replace("@P", " @P").split()
        toks = [t for t in toks if not t.startswith("@") and not t.startswith("/*")]
        if not toks:
            continue
        op = toks[0].rstrip(";")
        k = per.setdefault(kernel or "?", {"ops": collections.Counter(), "local": []})
        k["ops"][op] += n
        if op.startswith(("LDL", "STL")):
            k["local"].append((n, d.get("Address", ""), src.strip()))
    for name, k in per.items():
        tot = sum(k["ops"].values())
        if not tot:
            continue
        print("== %s: %d executed warp instructions" % (name, tot))
        for op, n in k["ops"].most_common(25):
            print("  %-22s %14d  %5.2f %%" % (op, n, 100.0 * n / tot))
        loc = sum(n for n, _, _ in k["local"])
        print("  local-memory (LDL/STL) executed: %d = %.3f %% of all" % (loc, 100.0 * loc / tot))
        for n, a, srcl in sorted(k["local"], reverse=True)[:8]:
            print("    %12d  %s  %s" % (n, a, srcl))


if __name__ == "__main__":
    main()
