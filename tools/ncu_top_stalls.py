#!/usr/bin/env python
"""Top stall sites of the kernels in an .ncu-rep (needs -lineinfo + --import-source on): per kernel the N SASS lines
with the most warp-stall samples and their share.  Usage: tools/ncu_top_stalls.py x.ncu-rep out.txt [N]"""
import csv
import subprocess
import sys


def main(rep, out, n=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    with open(out, "w") as f:
        for b in blocks:
            if len(b["rows"]) < 2:
                continue
            hdr, data = b["rows"][0], b["rows"][1:]
            ia, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            tot = sum(int(r[ia]) for r in data if r[ia].isdigit()) or 1
            f.write("== %s  (%d SASS lines, %d samples)\n" % (b["name"][:110], len(data), tot))
            for r in sorted(data, key=lambda r: -int(r[ia]) if r[ia].isdigit() else 0)[:n]:
                f.write("%6.2f%% %9s  %s\n" % (100.0 * int(r[ia]) / tot, r[ie], r[1][:100]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
