#!/usr/bin/env python
"""Warp-stall samples and executed instructions per CUDA source line of one kernel (needs -lineinfo).
Usage: tools/ncu_lines.py rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main(rep, pattern, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + pattern, "--launch-count", "1"], capture_output=True, text=True).stdout
    cur, hdr, rows = None, None, []
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            isamp, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            try:
                rows.append((int(r[isamp] or 0), int(r[iex] or 0), cur, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot, tote = sum(x[0] for x in rows), sum(x[1] for x in rows)
    print("total samples", tot, "| warp instructions", tote)
    byfile = {}
    for s, e, f, ln, src in rows:
        a = byfile.setdefault(f, [0, 0])
        a[0] += s
        a[1] += e
    for f, (s, e) in sorted(byfile.items(), key=lambda x: -x[1][0]):
        print("  %-20s samples %5.1f%%  instr %5.1f%%" % (f, 100.0 * s / max(tot, 1), 100.0 * e / max(tote, 1)))
    for s, e, f, ln, src in sorted(rows, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / max(tot, 1), 100.0 * e / max(tote, 1), f, ln, src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
