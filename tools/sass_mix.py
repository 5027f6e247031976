#!/usr/bin/env python
"""Executed-instruction report of ONE kernel from an `ncu -i X.ncu-rep --page source --csv --print-source sass` dump (optionally .gz):
executed warp instructions by opcode, and every local-memory instruction (LDL/STL = register spills) with how often it executes per
warp -- the evidence for "no spill traffic inside the multiply / window bodies".
    python tools/sass_mix.py gpurun_out/src_k_ladders.csv.gz > profiles/rNN_sass_k_ladders.txt"""
import collections
import csv
import gzip
import io
import sys


def opcode(src):
    toks = [t for t in src.split() if not t.startswith("@")]
    return toks[0].rstrip(";") if toks else "?"


def main():
    path = sys.argv[1]
    f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path, newline="")
    rows = list(csv.reader(f))
    name = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
    hdr = rows[1]
    ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    ins = [(int(r[ia], 16), r[isrc].strip(), int(r[iex])) for r in rows[2:] if len(r) > iex and r[ia].startswith("0x")]
    warps = ins[0][2]                      # the first instruction runs once per warp
    tot = sum(n for _, _, n in ins)
    print("kernel: %s" % name)
    print("static instructions: %d   warps launched: %d   executed warp instructions: %d (%.0f per warp on average)" % (len(ins), warps, tot, tot / warps))
    ops, static = collections.Counter(), collections.Counter()
    for _, s, n in ins:
        ops[opcode(s)] += n
        static[opcode(s)] += 1
    print("\nopcode                 executed warp instr     share   static count")
    for op, n in ops.most_common(30):
        print("  %-20s %18d  %6.2f %%  %8d" % (op, n, 100.0 * n / tot, static[op]))
    loc = [(a, s, n) for a, s, n in ins if opcode(s).startswith(("LDL", "STL"))]
    le = sum(n for _, _, n in loc)
    print("\nlocal-memory instructions (LDL/STL): %d static, %d executed = %.3f %% of all executed instructions" % (len(loc), le, 100.0 * le / tot))
    hist = collections.Counter()
    for _, _, n in loc:
        x = n / warps
        hist["never" if n == 0 else "< 1" if x < 1 else "1 .. 8" if x < 8 else "8 .. 64" if x < 64 else ">= 64"] += 1
    print("executions per launched warp -> number of LDL/STL instructions:", dict(hist))
    hottest = max(n for _, _, n in ins)
    print("for scale: the hottest instruction of the kernel executes %.0f times per launched warp; IMAD.WIDE instructions average %.0f"
          % (hottest / warps, sum(n for _, s, n in ins if opcode(s).startswith("IMAD.WIDE")) / max(1, sum(1 for _, s, _ in ins if opcode(s).startswith("IMAD.WIDE"))) / warps))
    print("the local-memory instructions that execute most often:")
    for a, s, n in sorted(loc, key=lambda t: -t[2])[:12]:
        print("  %8.1f per warp   0x%x   %s" % (n / warps, a, s))


if __name__ == "__main__":
    main()
