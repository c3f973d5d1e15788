#!/usr/bin/env python
"""Instruction mix and stall hot spots of one kernel from an .ncu-rep source page.
Usage: tools/ncu_source.py rep kernel_regex [top_n]"""
import collections
import csv
import subprocess
import sys


def main(rep, pattern, top=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = []
    for i, r in enumerate(rows[h + 1:]):
        if len(r) <= max(isamp, iex, ia):
            continue
        try:
            data.append((int(r[isamp] or 0), int(r[iex] or 0), i, r[ia]))
        except ValueError:
            pass
    tot = sum(d[0] for d in data)
    print("kernel", rows[0][1][:80] if rows[0] else "", "| total samples", tot, "| SASS instructions", len(data))
    mix, ex = collections.Counter(), collections.Counter()
    for s, e, i, src in data:
        parts = src.split()
        op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
        op = op.split(".")[0]
        mix[op] += 1
        ex[op] += e
    tote = sum(ex.values())
    print("executed warp-instr by opcode:", [(k, "%.1f%%" % (100.0 * v / tote)) for k, v in ex.most_common(16)])
    for s, e, i, src in sorted(data, reverse=True)[:top]:
        print("%6d %5.1f%%  #%d  %s" % (s, 100.0 * s / max(tot, 1), i, src.strip()))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
